#!/bin/bash
# round 2, GPU call F (2 GPUs): native band driver over real NCCL ranks - bit identity, then bench --gpus 2 (streams + bands records)
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02f
nvidia-smi -L
timeout -k 10 600 python -m pytest tests/test_band_driver.py -m gpu -x -q -p no:cacheprovider > $out/${tag}_pytest.txt 2>&1
echo "rc=$?" >> $out/${tag}_pytest.txt
tail -25 $out/${tag}_pytest.txt | cut -c1-300
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 64 --warmup 8 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
echo "bench rc=$?"; tail -c 1500 $out/${tag}_bench_n2.err
python -c "
import json
d=json.load(open('$out/${tag}_bench_n2.json'))
for k in ('value','ms_per_step','stage_ms_per_frame','streams_1080p','bands'): print(k, d.get(k))
print('e2e', {k:v for k,v in d['e2e'].items() if k!='api'})
"

#!/bin/bash
# round 2, GPU call L (8 GPUs): the driver's N = 8 launch of bench.py (streams + 64 x 1080p + 8K bands with bit-identity check)
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02l
nvidia-smi -L | wc -l
timeout -k 10 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 64 --warmup 8 > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err
echo "bench rc=$?"; grep -v Warning $out/${tag}_bench_n8.err | tail -5 | cut -c1-300
python -c "
import json
d=json.loads(open('$out/${tag}_bench_n8.json').read().splitlines()[-1])
for k in ('value','ms_per_step','stage_ms_per_frame','streams_1080p','bands'): print(k, d.get(k))
print('e2e', {k:v for k,v in d['e2e'].items() if k!='api'})
"

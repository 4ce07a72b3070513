#!/bin/bash
# round 2, GPU call N (2 GPUs): band driver after the two-range / merged-exchange changes - bit identity + timing
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02n
timeout -k 10 600 python -m pytest tests/test_band_driver.py -m gpu -x -q -p no:cacheprovider > $out/${tag}_pytest.txt 2>&1
echo "rc=$?" >> $out/${tag}_pytest.txt
tail -6 $out/${tag}_pytest.txt | cut -c1-300
run() { timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 295$3 bench.py --gpus 2 --mode bands --workload 8k --steps 32 --warmup 6 $2 > $out/${tag}_$1.json 2> $out/${tag}_$1.err; python -c "
import json
d=json.loads(open('$out/${tag}_$1.json').read().splitlines()[-1])
print('$1', d['ms_per_step'], d['config']['ms_per_step_by_rank'], d['config']['local_rows_by_rank'], d['config']['bit_identical_to_one_gpu'], d['config'].get('calibration_ms_by_rank'))
"; }
run bal2 "--band-check-frames 3" 61
run bal1 "--band-check-frames 0 --band-balance 1" 62

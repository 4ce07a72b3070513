#!/bin/bash
# round 2, GPU call U (2 GPUs): band transports over real ranks - bit identity (pytest) and 8K timing: ipc / nccl / no exchange
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02u
mkdir -p $out
N=${1:-2}
timeout -k 10 600 python -m pytest tests/test_band_driver.py -m gpu -k "real_ranks" -x -q -p no:cacheprovider > $out/${tag}_pytest.txt 2>&1
echo "rc=$?" >> $out/${tag}_pytest.txt; tail -4 $out/${tag}_pytest.txt | cut -c1-300
run() { timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$3 bench.py --gpus $N --mode bands --workload 8k --steps 32 --warmup 6 $2 > $out/${tag}_$1.json 2> $out/${tag}_$1.err; python -c "
import json
d=json.loads(open('$out/${tag}_$1.json').read().splitlines()[-1])
print('$1', d['ms_per_step'], d['config']['ms_per_step_by_rank'], d['config']['local_rows_by_rank'], d['config']['bit_identical_to_one_gpu'], d['config']['transport'], d['gpu_launches'])
" || tail -5 $out/${tag}_$1.err; }
run ipc "--band-transport ipc --band-check-frames 3" 81
run nccl "--band-transport nccl --band-check-frames 0" 82
run dry "--band-transport ipc --band-check-frames 0 --flags 512" 83

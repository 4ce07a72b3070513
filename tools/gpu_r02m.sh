#!/bin/bash
# round 2, GPU call M (8 GPUs): 8K bands - with exchanges vs compute only (dry run), balanced vs equal heights
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02m
run() { timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 295$3 bench.py --gpus 8 --mode bands --workload 8k --steps 32 --warmup 6 $2 > $out/${tag}_$1.json 2> $out/${tag}_$1.err; python -c "
import json
d=json.loads(open('$out/${tag}_$1.json').read().splitlines()[-1])
print('$1', d['ms_per_step'], d['config']['ms_per_step_by_rank'], d['config']['local_rows_by_rank'], d['config']['bit_identical_to_one_gpu'])
"; }
run bal "--band-check-frames 1" 51
run bal_dry "--band-check-frames 0 --flags 512" 52
run eq "--band-check-frames 0 --band-balance 0" 53
run eq_dry "--band-check-frames 0 --band-balance 0 --flags 512" 54

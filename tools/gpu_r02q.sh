#!/bin/bash
# round 2, GPU call Q (8 GPUs): the driver's N=8 line (bands + streams records), then the 8K band frame under two NCCL channel limits
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02q
mkdir -p $out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -k 10 500 $T --master-port 29561 bench.py --gpus 8 > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err; echo "rc=$?"
python -c "
import json
d=json.loads(open('$out/${tag}_bench_n8.json').read().splitlines()[-1])
print('n8', d['value'], d['ms_per_step'], 'e2e', d['e2e'].get('value'))
print('bands', {k:v for k,v in (d.get('bands') or {}).items() if k!='plan'})
print('streams', d.get('streams_1080p'))
"
run() { timeout -k 10 300 $T --master-port 295$3 bench.py --gpus 8 --mode bands --workload 8k --steps 32 --warmup 6 $2 > $out/${tag}_$1.json 2> $out/${tag}_$1.err; python -c "
import json
d=json.loads(open('$out/${tag}_$1.json').read().splitlines()[-1])
print('$1', d['ms_per_step'], d['config']['ms_per_step_by_rank'], d['config']['local_rows_by_rank'], d['config']['bit_identical_to_one_gpu'])
"; }
NCCL_MAX_P2P_NCHANNELS=4 run ch4 "--band-check-frames 2" 62
NCCL_MAX_P2P_NCHANNELS=8 run ch8 "--band-check-frames 0" 63

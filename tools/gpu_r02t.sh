#!/bin/bash
# round 2, GPU call T (1 GPU): peer-memory band transport with N processes time-slicing one GPU (real IPC, real flags)
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02t
mkdir -p $out
run() { timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 295$2 tools/band_check.py --transport ipc --same-gpu $3 > $out/${tag}_$2.json 2> $out/${tag}_$2.err; echo "rc=$? n=$1 $3"; grep "^{" $out/${tag}_$2.json | tail -1 | cut -c1-400; grep -iE "error|trap|Traceback" $out/${tag}_$2.err | head -5; }
run 2 71 "--width 1280 --height 720 --frames 4"
run 3 72 "--width 1920 --height 1080 --frames 5"
run 4 73 "--width 1030 --height 420 --frames 5 --storage f32"
run 3 74 "--width 1280 --height 720 --frames 4 --levels 3"
run 8 75 "--width 1920 --height 1080 --frames 4"
nvidia-smi --query-gpu=name,memory.used --format=csv

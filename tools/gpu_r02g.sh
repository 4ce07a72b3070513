#!/bin/bash
# round 2, GPU call G (2 GPUs): 8K bands timing, balanced vs equal heights, after the priority fix
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02g
for bal in 1 0; do
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --mode bands --workload 8k --steps 24 --warmup 4 --band-balance $bal --band-check-frames 2 > $out/${tag}_bands_n2_bal$bal.json 2> $out/${tag}_bands_n2_bal$bal.err
echo "rc=$?"; tail -n 1 $out/${tag}_bands_n2_bal$bal.json | cut -c1-1500
done
timeout -k 10 300 python bench.py --workload 8k --steps 24 --warmup 4 --skip-extras --no-cpu-baseline > $out/${tag}_8k_n1.json 2> $out/${tag}_8k_n1.err; tail -n 1 $out/${tag}_8k_n1.json | cut -c1-400

#!/bin/bash
# round 2, GPU call A: first hardware run of the staged (TMA) a-trous path
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
timeout -k 10 900 python -m pytest tests/test_staged_levels.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02a_pytest_staged.txt 2>&1
echo "staged rc=$?" >> gpurun_out/r02a_pytest_staged.txt
tail -5 gpurun_out/r02a_pytest_staged.txt
for fl in 0 32 256 8 40; do
  timeout -k 10 300 python bench.py --steps 64 --warmup 8 --flags $fl --no-cpu-baseline > gpurun_out/r02a_bench_flags$fl.json 2> gpurun_out/r02a_bench_flags$fl.err
  tail -c 1500 gpurun_out/r02a_bench_flags$fl.json
done
timeout -k 10 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_staged_levels.py > gpurun_out/r02a_pytest_rest.txt 2>&1
echo "rest rc=$?" >> gpurun_out/r02a_pytest_rest.txt
tail -5 gpurun_out/r02a_pytest_rest.txt

#!/bin/bash
# round 2, GPU call D (1 GPU): whole -m gpu suite on the current tree + bench A/B
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02d
mkdir -p $out/profiles
timeout -k 10 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $out/${tag}_pytest.txt 2>&1
echo "rc=$?" >> $out/${tag}_pytest.txt
tail -15 $out/${tag}_pytest.txt
for fl in 0 32 8; do
  timeout -k 10 300 python bench.py --steps 64 --warmup 8 --flags $fl --no-cpu-baseline > $out/${tag}_bench_flags$fl.json 2> $out/${tag}_bench_flags$fl.err
  python - <<PY
import json
d=json.load(open("$out/${tag}_bench_flags$fl.json"))
print("flags $fl", d["ms_per_step"], d["stage_ms_per_frame"], d["e2e"]["ms_per_step"])
PY
done

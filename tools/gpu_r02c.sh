#!/bin/bash
# round 2, GPU call C: persistent lattice kernel (correctness + time), TAA tests
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02c
mkdir -p $out/profiles
timeout -k 10 900 python -m pytest tests/test_staged_levels.py tests/test_taa.py tests/test_host_path.py -m gpu -x -q -p no:cacheprovider > $out/${tag}_pytest.txt 2>&1
echo "rc=$?" >> $out/${tag}_pytest.txt
tail -15 $out/${tag}_pytest.txt
for fl in 0 256 8; do
  timeout -k 10 300 python bench.py --steps 64 --warmup 8 --flags $fl --no-cpu-baseline > $out/${tag}_bench_flags$fl.json 2> $out/${tag}_bench_flags$fl.err
  python - <<PY
import json
d=json.load(open("$out/${tag}_bench_flags$fl.json"))
print("flags $fl", d["ms_per_step"], d["stage_ms_per_frame"], d["e2e"]["ms_per_step"])
PY
done
B="python bench.py --steps 6 --warmup 3 --ring 9 --no-cpu-baseline --e2e-steps 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv $B > $out/launches_$tag.log 2>&1
python tools/ncu_summary.py --launches $out/launches_$tag.csv $out/profiles/launches_$tag.md
cat $out/profiles/launches_$tag.md
timeout 900 ncu --set full --clock-control none --import-source on -k regex:atrous_lattice -s 20 -c 2 -o $out/atrous_$tag -f $B > $out/atrous_$tag.log 2>&1
python tools/ncu_summary.py $out/atrous_$tag.ncu-rep $out/profiles/atrous_$tag

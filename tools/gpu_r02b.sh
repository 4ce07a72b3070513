#!/bin/bash
# round 2, GPU call B: where does the staged a-trous frame spend its time (launch list + full ncu set of one frame's levels)
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02b
mkdir -p $out/profiles
B="python bench.py --steps 6 --warmup 3 --ring 9 --no-cpu-baseline --e2e-steps 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv $B > $out/launches_$tag.log 2>&1
python tools/ncu_summary.py --launches $out/launches_$tag.csv $out/profiles/launches_$tag.md
cat $out/profiles/launches_$tag.md
timeout 900 ncu --set full --clock-control none --import-source on -k regex:atrous_ -s 25 -c 5 -o $out/atrous_$tag -f $B > $out/atrous_$tag.log 2>&1
python tools/ncu_summary.py $out/atrous_$tag.ncu-rep $out/profiles/atrous_$tag
ls -la $out/*.ncu-rep
# general case (no uniform tiles)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:atrous_lattice -s 16 -c 4 -o $out/atrous_general_$tag -f $B --flags 8 > $out/atrous_general_$tag.log 2>&1
python tools/ncu_summary.py $out/atrous_general_$tag.ncu-rep $out/profiles/atrous_general_$tag
rm -f $out/atrous_general_$tag.ncu-rep

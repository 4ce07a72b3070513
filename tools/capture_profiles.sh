#!/bin/bash
# Runs on the GPU box (under gpurun, ONE GPU): the ncu evidence kept under profiles/.
#   tools/capture_profiles.sh <tag>     -> gpurun_out/profiles/{launches,atrous,atrous_general,temporal,taa}_<tag>.*
# (summaries written by tools/ncu_summary.py on the box; the .ncu-rep files are 25-50 MB each and are dropped).
# Numbers printed by bench.py under ncu are never bench values.
set -u
tag=${1:-r02}
out=gpurun_out
mkdir -p $out/profiles
B="python bench.py --steps 6 --warmup 3 --ring 9 --no-cpu-baseline --skip-extras --e2e-steps 0"
# per-launch durations of the whole command (the kernel's SHARE of the step is what must agree with bench.py)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv $B > $out/launches_$tag.log 2>&1
python tools/ncu_summary.py --launches $out/launches_$tag.csv $out/profiles/launches_$tag.md
cp $out/launches_$tag.csv $out/profiles/launches_$tag.csv
# full-set captures: the five a-trous levels of a steady-state frame (packed level 0 + four lattice levels), the same
# with the uniform-tile shortcut off, one temporal launch, one sparse-variance launch, one TAA launch
timeout 900 ncu --set full --clock-control none --import-source on -k regex:atrous_ -s 25 -c 5 -o $out/atrous_$tag -f $B > $out/atrous_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:atrous_ -s 25 -c 5 -o $out/atrous_general_$tag -f $B --flags 8 > $out/atrous_general_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:temporal_kernel -s 5 -c 1 -o $out/temporal_$tag -f $B > $out/temporal_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:variance_sparse -s 6 -c 1 -o $out/variance_$tag -f $B > $out/variance_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:taa_kernel -s 2 -c 1 -o $out/taa_$tag -f $B > $out/taa_$tag.log 2>&1
for k in atrous atrous_general temporal variance taa; do
  [ -f $out/${k}_$tag.ncu-rep ] && python tools/ncu_summary.py $out/${k}_$tag.ncu-rep $out/profiles/${k}_$tag && rm -f $out/${k}_$tag.ncu-rep
done
ls -la $out/profiles

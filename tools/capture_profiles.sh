#!/bin/bash
# Runs on the GPU box (under gpurun, ONE GPU): the ncu evidence kept under profiles/.
#   tools/capture_profiles.sh <tag>     -> gpurun_out/launches_<tag>.csv and gpurun_out/profiles/{launches,atrous,temporal,fused}_<tag>.*
# (summaries written by tools/ncu_summary.py on the box).  Numbers printed by bench.py under ncu are never bench values.
set -u
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 6 --warmup 3 --ring 9 --no-cpu-baseline --e2e-steps 0"
# per-launch durations of the whole command (the kernel's SHARE of the step is what must agree with bench.py)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv $B > $out/launches_$tag.log 2>&1
# full-set captures: five consecutive a-trous levels of a steady-state frame, one temporal launch, one fused 0+1 launch
timeout 900 ncu --set full --clock-control none --import-source on -k regex:atrous_packed -s 25 -c 5 -o $out/atrous_$tag -f $B > $out/atrous_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:temporal_kernel -s 5 -c 1 -o $out/temporal_$tag -f $B > $out/temporal_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:atrous_fused01 -s 5 -c 1 -o $out/fused_$tag -f $B --flags 16 > $out/fused_$tag.log 2>&1
# summarise on the box (the reports are 40-50 MB each; gpurun_out/ travels back only below 64 MiB) and drop the reports
mkdir -p $out/profiles
python tools/ncu_summary.py --launches $out/launches_$tag.csv $out/profiles/launches_$tag.md
for k in atrous temporal fused; do
  [ -f $out/${k}_$tag.ncu-rep ] && python tools/ncu_summary.py $out/${k}_$tag.ncu-rep $out/profiles/${k}_$tag && rm -f $out/${k}_$tag.ncu-rep
done
ls -la $out/profiles

#!/bin/bash
# round 2, GPU call J (1 GPU): occupancy probe - forced-uniform lattice kernel at 3 CTAs/SM (exp3) vs 2 CTAs/SM (exp4); TIMING ONLY
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02j
run() { timeout -k 10 300 python bench.py --steps 64 --warmup 8 --skip-extras --no-cpu-baseline --e2e-steps 0 $2 > $out/${tag}_bench_$1.json 2> $out/${tag}_bench_$1.err; python -c "
import json
d=json.loads(open('$out/${tag}_bench_$1.json').read().splitlines()[-1])
print('$1', d['ms_per_step'], d['stage_ms_per_frame'])
"; }
cp svgf_b200/libsvgf_b200.so /tmp/libsvgf_b200.keep
for e in 3 4; do
cp svgf_b200/libsvgf_b200_exp$e.so svgf_b200/libsvgf_b200.so
run exp$e ""
done
cp /tmp/libsvgf_b200.keep svgf_b200/libsvgf_b200.so
run default ""
B="python bench.py --steps 6 --warmup 3 --ring 9 --no-cpu-baseline --skip-extras --e2e-steps 0"
cp svgf_b200/libsvgf_b200_exp3.so svgf_b200/libsvgf_b200.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:atrous_lattice -s 16 -c 2 -o $out/atrous_exp3 -f $B > $out/atrous_exp3.log 2>&1
python tools/ncu_summary.py $out/atrous_exp3.ncu-rep $out/profiles/atrous_exp3_$tag; rm -f $out/atrous_exp3.ncu-rep
cp /tmp/libsvgf_b200.keep svgf_b200/libsvgf_b200.so

#!/bin/bash
# round 2, GPU call H (1 GPU): tests on the current tree; A/B of the 6-row lattice tile (4 CTAs/SM); TAA timing
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02h
timeout -k 10 1200 python -m pytest tests/test_taa.py tests/test_paper_extensions.py tests/test_staged_levels.py tests/test_texture_gbuffer.py -m gpu -q -p no:cacheprovider > $out/${tag}_pytest.txt 2>&1
echo "rc=$?" >> $out/${tag}_pytest.txt
tail -12 $out/${tag}_pytest.txt | cut -c1-300
run() { timeout -k 10 300 python bench.py --steps 64 --warmup 8 --skip-extras --no-cpu-baseline $2 > $out/${tag}_bench_$1.json 2> $out/${tag}_bench_$1.err; python -c "
import json
d=json.loads(open('$out/${tag}_bench_$1.json').read().splitlines()[-1])
print('$1', d['ms_per_step'], d['stage_ms_per_frame'])
"; }
run default ""
run default_general "--flags 8"
cp svgf_b200/libsvgf_b200.so /tmp/libsvgf_b200.keep
cp svgf_b200/libsvgf_b200_exp2.so svgf_b200/libsvgf_b200.so
timeout -k 10 600 python -m pytest tests/test_staged_levels.py -m gpu -q -x -p no:cacheprovider -k "teacher_forced or level_by_level" > $out/${tag}_pytest_exp2.txt 2>&1; tail -3 $out/${tag}_pytest_exp2.txt
run exp2 ""
run exp2_general "--flags 8"
run exp2_1080p "--workload 1080p"
cp /tmp/libsvgf_b200.keep svgf_b200/libsvgf_b200.so
run default_1080p "--workload 1080p"

#!/bin/bash
# round 2, GPU call O (1 GPU): full -m gpu suite, smoke, bench (both arms), ncu evidence, carve-out experiment on the 64-stream record
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02o
mkdir -p $out/profiles
timeout -k 10 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/${tag}_pytest.txt 2>&1
echo "rc=$?" >> $out/${tag}_pytest.txt
tail -6 $out/${tag}_pytest.txt | cut -c1-300
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.txt 2>&1; tail -2 $out/${tag}_smoke.txt
timeout -k 10 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$out/${tag}_bench.json').read().splitlines()[-1])
for k in ('value','ms_per_step','stage_ms_per_frame','general_case','parity','taa','streams_1080p'): print(k, d.get(k))
print('e2e', {k:v for k,v in d['e2e'].items() if k!='api'})
print('roofline', d['roofline'])
"
timeout -k 10 600 python bench.py --impl reference > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $out/${tag}_bench_ref.json | tail -1
bash tools/capture_profiles.sh r02 > $out/${tag}_capture.log 2>&1; tail -3 $out/${tag}_capture.log
# carve-out experiment: 64 x 1080p streams record with every kernel on the max-shared carve-out
cp svgf_b200/libsvgf_b200.so /tmp/libsvgf_b200.keep
cp svgf_b200/libsvgf_b200_exp5.so svgf_b200/libsvgf_b200.so
timeout -k 10 400 python bench.py --steps 32 --warmup 8 --no-cpu-baseline > $out/${tag}_bench_exp5.json 2> $out/${tag}_bench_exp5.err
python -c "
import json
d=json.loads(open('$out/${tag}_bench_exp5.json').read().splitlines()[-1])
print('exp5', d['ms_per_step'], d['stage_ms_per_frame'], d['streams_1080p']['value'], d['streams_1080p']['ms_per_frame_per_stream_slot'])
"
cp /tmp/libsvgf_b200.keep svgf_b200/libsvgf_b200.so

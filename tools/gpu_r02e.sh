#!/bin/bash
# round 2, GPU call E (1 GPU): TAA + texture tests, new bench.py (both arms), whole-frame parity report
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02e
mkdir -p $out/profiles
timeout -k 10 900 python -m pytest tests/test_taa.py tests/test_texture_gbuffer.py tests/test_band_driver.py -m gpu -q -p no:cacheprovider > $out/${tag}_pytest.txt 2>&1
echo "rc=$?" >> $out/${tag}_pytest.txt
tail -8 $out/${tag}_pytest.txt
timeout -k 10 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; tail -c 600 $out/${tag}_bench.err
python -c "
import json
d=json.load(open('$out/${tag}_bench.json'))
for k in ('value','ms_per_step','stage_ms_per_frame','general_case','parity','cpu_baseline','taa','streams_1080p'): print(k, d.get(k))
print('e2e', {k:v for k,v in d['e2e'].items() if k!='api'})
print('cfg', d['config']['timed_region'])
"
timeout -k 10 600 python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"; tail -c 300 $out/${tag}_bench_ref.err; cut -c1-600 $out/${tag}_bench_ref.json
timeout -k 10 900 python tools/parity_report.py --out $out/parity_r02.json > $out/${tag}_parity.log 2>&1; echo "parity rc=$?"; tail -20 $out/${tag}_parity.log | cut -c1-400

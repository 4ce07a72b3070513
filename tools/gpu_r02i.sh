#!/bin/bash
# round 2, GPU call I (1 GPU): per-warp uniform decision - tests + bench; TAA pin; FP64 cascade truth merged into the parity report
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02i
timeout -k 10 1200 python -m pytest tests/test_taa.py tests/test_staged_levels.py tests/test_uniform_tiles.py tests/test_parity_sequence.py -m gpu -q -p no:cacheprovider > $out/${tag}_pytest.txt 2>&1
echo "rc=$?" >> $out/${tag}_pytest.txt
tail -12 $out/${tag}_pytest.txt | cut -c1-400
run() { timeout -k 10 300 python bench.py --steps 64 --warmup 8 --skip-extras --no-cpu-baseline $2 > $out/${tag}_bench_$1.json 2> $out/${tag}_bench_$1.err; python -c "
import json
d=json.loads(open('$out/${tag}_bench_$1.json').read().splitlines()[-1])
print('$1', d['ms_per_step'], d['stage_ms_per_frame'])
"; }
run default ""
run general "--flags 8"
cp profiles/parity_r02.json $out/parity_r02.json
timeout -k 10 600 python tools/parity_report.py --truth-only --out $out/parity_r02.json > $out/${tag}_truth.log 2>&1; tail -3 $out/${tag}_truth.log | cut -c1-900

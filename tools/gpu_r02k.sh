#!/bin/bash
# round 2, GPU call K (1 GPU): per-warp uniform decision in the packed kernel - tests + bench
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02k
timeout -k 10 1500 python -m pytest tests/test_staged_levels.py tests/test_uniform_tiles.py tests/test_parity_stages.py tests/test_fused_levels.py tests/test_bands.py -m gpu -q -p no:cacheprovider > $out/${tag}_pytest.txt 2>&1
echo "rc=$?" >> $out/${tag}_pytest.txt
tail -8 $out/${tag}_pytest.txt | cut -c1-400
run() { timeout -k 10 300 python bench.py --steps 64 --warmup 8 --skip-extras --no-cpu-baseline --e2e-steps 0 $2 > $out/${tag}_bench_$1.json 2> $out/${tag}_bench_$1.err; python -c "
import json
d=json.loads(open('$out/${tag}_bench_$1.json').read().splitlines()[-1])
print('$1', d['ms_per_step'], d['stage_ms_per_frame'])
"; }
run default ""
run nostaged "--flags 32"
run general "--flags 8"
run p1080 "--workload 1080p"

#!/bin/bash
# round 2, GPU call R (1 GPU): level 0 with the odd offsets packed across directions - tests + A/B against the previous library
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02r
mkdir -p $out
timeout -k 10 1500 python -m pytest tests/test_staged_levels.py tests/test_uniform_tiles.py tests/test_fused_levels.py tests/test_parity_stages.py tests/test_parity_sequence.py tests/test_reference_kernels.py tests/test_bands.py -m gpu -q -p no:cacheprovider > $out/${tag}_pytest.txt 2>&1
echo "rc=$?" >> $out/${tag}_pytest.txt
tail -8 $out/${tag}_pytest.txt | cut -c1-300
B="python bench.py --no-cpu-baseline --skip-extras --e2e-steps 0"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().splitlines()[-1])
print('$1', d['ms_per_step'], d['stage_ms_per_frame'], d.get('parity'))
"; }
cp svgf_b200/libsvgf_b200.so /tmp/new.so
for i in 1 2; do
  cp /tmp/new.so svgf_b200/libsvgf_b200.so
  timeout -k 10 300 $B > $out/${tag}_new$i.json 2>> $out/${tag}.err; show $out/${tag}_new$i.json
  timeout -k 10 300 $B --flags 8 > $out/${tag}_new_gen$i.json 2>> $out/${tag}.err; show $out/${tag}_new_gen$i.json
  cp svgf_b200/libsvgf_b200_base.so svgf_b200/libsvgf_b200.so
  timeout -k 10 300 $B > $out/${tag}_old$i.json 2>> $out/${tag}.err; show $out/${tag}_old$i.json
done
cp /tmp/new.so svgf_b200/libsvgf_b200.so
timeout -k 10 300 $B --flags 32 > $out/${tag}_new_nostaged.json 2>> $out/${tag}.err; show $out/${tag}_new_nostaged.json
timeout -k 10 300 $B --workload 1080p > $out/${tag}_new_1080p.json 2>> $out/${tag}.err; show $out/${tag}_new_1080p.json

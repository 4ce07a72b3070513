#!/bin/bash
# round 2, final validation (1 GPU): full -m gpu suite, smoke, bench (both arms + variants), parity report, ncu evidence
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02w
mkdir -p $out/profiles
timeout -k 10 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/${tag}_pytest.txt 2>&1
echo "rc=$?" >> $out/${tag}_pytest.txt
tail -6 $out/${tag}_pytest.txt | cut -c1-300
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.txt 2>&1; tail -2 $out/${tag}_smoke.txt
timeout -k 10 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$out/${tag}_bench.json').read().splitlines()[-1])
for k in ('value','ms_per_step','stage_ms_per_frame','general_case','parity','taa','streams_1080p','cpu_baseline','gpu_launches'): print(k, d.get(k))
print('e2e', {k:v for k,v in d['e2e'].items() if k!='api'})
print('roofline', d['roofline'])
"
timeout -k 10 600 python bench.py --impl reference > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"; cut -c1-400 $out/${tag}_bench_ref.json | tail -1
B="python bench.py --no-cpu-baseline --skip-extras --e2e-steps 0"
for v in "general:--flags 8" "nostaged:--flags 32" "1080p:--workload 1080p" "8k:--workload 8k" "f32:--storage f32"; do
  n=${v%%:*}; a=${v#*:}
  timeout -k 10 300 $B $a > $out/${tag}_bench_$n.json 2>> $out/${tag}.err
  python -c "
import json
d=json.loads(open('$out/${tag}_bench_$n.json').read().splitlines()[-1])
print('$n', d['value'], d['ms_per_step'], d['stage_ms_per_frame'])
"
done
timeout -k 10 1200 python tools/parity_report.py --out $out/parity_r02.json > $out/${tag}_parity.log 2>&1; echo "parity rc=$?"; tail -3 $out/${tag}_parity.log | cut -c1-300
bash tools/capture_profiles.sh r02 > $out/${tag}_capture.log 2>&1; tail -3 $out/${tag}_capture.log

#!/bin/bash
# round 2, GPU call V (8 GPUs): the driver's N=8 line with both band transports
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out; tag=r02y
mkdir -p $out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout -k 10 600 $T --master-port 29561 bench.py --gpus 4 > $out/${tag}_bench_n4.json 2> $out/${tag}_bench_n4.err; echo "rc=$?"
python -c "
import json
d=json.loads(open('$out/${tag}_bench_n4.json').read().splitlines()[-1])
print('n4', d['value'], d['ms_per_step'], 'e2e', d['e2e'].get('value'))
print('bands', {k:v for k,v in (d.get('bands') or {}).items() if k not in ('workload',)})
print("other", {k:v for k,v in d.items() if k.startswith("bands_")})
print('streams', (d.get('streams_1080p') or {}).get('value'))
" || tail -20 $out/${tag}_bench_n4.err

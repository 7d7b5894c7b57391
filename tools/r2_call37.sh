#!/bin/bash
# round 2, call 37: SURVEY 8d batch-size / distribution sweep on the final build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for b in 512 4096 262144; do timeout 200 python bench.py --no-cpu-baseline --api-steps 0 --batch $b > $O/c37_bench_b$b.json 2> $O/c37_bench_b$b.err; echo "b$b rc=$?"; done
timeout 200 python bench.py --no-cpu-baseline --api-steps 0 --dist uniform > $O/c37_bench_uniform.json 2> $O/c37_bench_uniform.err; echo "uniform rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c37_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f,'ERR', e)
PY

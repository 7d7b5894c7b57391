#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/c28_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c28_pytest.log
tail -3 $O/c28_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/c28_smoke.log 2>&1; tail -1 $O/c28_smoke.log
for i in 1 2 3; do timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c28_bench$i.json 2> $O/c28_bench$i.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c28_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/c19_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c19_pytest.log
tail -6 $O/c19_pytest.log
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c19_bench.json 2> $O/c19_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 --steps 50 > $O/c19_bench_k50.json 2> $O/c19_bench_k50.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c19_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), d['final_loss'], d.get('e2e_losses_read_on_host'), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 $O/c19_bench.err

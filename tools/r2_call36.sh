#!/bin/bash
# round 2, call 36: range flag of the bounded dedup reported once per unit / block (the in-loop atomic had serialised the id loads): re-validate
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > $O/c36_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c36_pytest.log
tail -3 $O/c36_pytest.log
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c36_bench.json 2> $O/c36_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('c36_bench',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f,'ERR', e)
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c36_launches.csv \
    python bench.py --profile-step --warmup 3 --no-cpu-baseline --api-steps 0 > $O/c36_launches_bench.log 2>&1; echo "launches rc=$?"
python profiles/launch_list.py $O/c36_launches.csv | head -5

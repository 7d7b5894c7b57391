#!/bin/bash
# round 2, call 35: validation of the final tree (full GPU suite, smoke, default bench), paired-record A/B, ncu evidence of the final build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > $O/c35_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c35_pytest.log
tail -4 $O/c35_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/c35_smoke.log 2>&1; tail -2 $O/c35_smoke.log
timeout 400 python bench.py > $O/c35_bench.json 2> $O/c35_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --no-cpu-baseline --api-steps 0 --paired on > $O/c35_bench_paired.json 2> $O/c35_bench_paired.err; echo "paired rc=$?"
python - <<'PY'
import json
for f in ('c35_bench','c35_bench_paired'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
        a=d.get('api_path')
        if a: print({k:(round(v['ms_per_step'],3) if isinstance(v,dict) else v) for k,v in a.items() if k!='what'})
    except Exception as e: print(f,'ERR', e)
PY
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c35_launches.csv \
    python bench.py --profile-step --warmup 3 --no-cpu-baseline --api-steps 0 > $O/c35_launches_bench.log 2>&1; echo "launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/c35_prof_step \
    python bench.py --profile-step --warmup 3 --no-cpu-baseline --api-steps 0 > $O/c35_prof_bench.log 2>&1; echo "ncu full rc=$?"
ls -la $O/c35_prof_step.ncu-rep $O/c35_launches.csv

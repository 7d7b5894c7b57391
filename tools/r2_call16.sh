#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/c16_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c16_pytest.log
tail -3 $O/c16_pytest.log
BK_ONLY=adam timeout 300 python bench_kernels.py > $O/c16_kernels.jsonl 2> $O/c16_kernels.err; echo "kernels rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/c16_kernels.jsonl'):
    d=json.loads(l); print(d['kernel'], d.get('dim'), d.get('unique_ids', d.get('k', d.get('replicas'))), round(d['us'],1), 'us', round(d['frac_of_peak'],3))
PY
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c16_bench.json 2> $O/c16_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c16_bench2.json 2> $O/c16_bench2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c16_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c16_launches_step.csv \
    python bench.py --profile-step --warmup 3 --no-cpu-baseline --api-steps 0 > $O/c16_launches_bench.log 2>&1; echo "ncu list rc=$?"
timeout 300 python tools/bench_census.py --steps 100 > $O/c16_census.jsonl 2> $O/c16_census.err; echo "census rc=$?"; cat $O/c16_census.jsonl; tail -3 $O/c16_census.err

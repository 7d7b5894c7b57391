#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/c13_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c13_pytest.log
tail -3 $O/c13_pytest.log
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c13_bench.json 2> $O/c13_bench.err; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c13_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
BK_ONLY=adam timeout 300 python bench_kernels.py > $O/c13_kernels.jsonl 2> $O/c13_kernels.err; echo "kernels rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/c13_kernels.jsonl'):
    d=json.loads(l); print(d['kernel'], d.get('dim'), d.get('unique_ids', d.get('k', d.get('replicas'))), round(d['us'],1), 'us', round(d['frac_of_peak'],3))
PY
BK_SIZES=1000000 BK_ONLY=adam:64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_copy_flat -s 2 -c 1 -f -o $O/c13_prof_pull64 python bench_kernels.py > $O/c13_prof_pull64.log 2>&1; echo "ncu rc=$?"
BK_SIZES=4000000 BK_ONLY=adam:8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_copy_flat -s 2 -c 1 -f -o $O/c13_prof_pull8 python bench_kernels.py > $O/c13_prof_pull8.log 2>&1; echo "ncu rc=$?"
BK_SIZES=4000000 BK_ONLY=adam:8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push_flat -s 2 -c 1 -f -o $O/c13_prof_push8 python bench_kernels.py > $O/c13_prof_push8.log 2>&1; echo "ncu rc=$?"

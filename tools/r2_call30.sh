#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/c30_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c30_pytest.log
tail -3 $O/c30_pytest.log
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c30_bench$i.json 2> $O/c30_bench$i.err; done
B200_FLAT_U=2 timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c30_bench_u2.json 2> $O/c30_bench_u2.err
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 --dist uniform > $O/c30_bench_uniform.json 2> $O/c30_bench_uniform.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c30_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
BK_ONLY=adam timeout 300 python bench_kernels.py > $O/c30_kernels.jsonl 2>/dev/null; python - <<'PY'
import json
for l in open('gpurun_out/c30_kernels.jsonl'):
    d=json.loads(l); print(d['kernel'], d.get('dim'), d.get('unique_ids'), round(d['us'],1), 'us', round(d['frac_of_peak'],3))
PY

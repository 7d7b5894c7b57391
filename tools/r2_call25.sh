#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/c25_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c25_pytest.log
tail -4 $O/c25_pytest.log
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c25_bench.json 2> $O/c25_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c25_bench2.json 2> $O/c25_bench2.err
B200_UNIQUE_SMALL=0 timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c25_bench_nosmall.json 2> $O/c25_bench_nosmall.err
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 --dist uniform > $O/c25_bench_uniform.json 2> $O/c25_bench_uniform.err
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 --batch 4096 > $O/c25_bench_b4096.json 2> $O/c25_bench_b4096.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c25_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
tail -2 $O/c25_bench.err

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/c10_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c10_pytest.log
tail -4 $O/c10_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/c10_smoke.log 2>&1; tail -1 $O/c10_smoke.log
timeout 300 tools/probes/gather_roof > $O/c10_gather_roof.jsonl 2> $O/c10_gather_roof.err; echo "probe rc=$?"; cat $O/c10_gather_roof.jsonl
timeout 600 python bench.py > $O/c10_bench_default.json 2> $O/c10_bench_default.err; echo "bench rc=$?"
B200_STEP_BRANCHES=0 timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c10_bench_nobranch.json 2> $O/c10_bench_nobranch.err
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c10_bench_branch2.json 2> $O/c10_bench_branch2.err
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 --dist uniform > $O/c10_bench_uniform.json 2> $O/c10_bench_uniform.err
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 --batch 262144 --pool 4 > $O/c10_bench_b262144.json 2> $O/c10_bench_b262144.err
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 --batch 4096 > $O/c10_bench_b4096.json 2> $O/c10_bench_b4096.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c10_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, d.get('api_path'))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c10_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --api-steps 0 > $O/c10_launches_bench.log 2>&1; echo "ncu list rc=$?"

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/c23_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c23_pytest.log
tail -4 $O/c23_pytest.log
timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c23_bench.json 2> $O/c23_bench.err; echo "bench rc=$?"
B200_UNIQUE_SMALL=16384 timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c23_bench_small.json 2> $O/c23_bench_small.err
B200_UNIQUE_SMALL=300 timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c23_bench_small300.json 2> $O/c23_bench_small300.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c23_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 200 python tools/unique_timeline.py > $O/c23_unique_timeline.jsonl 2> $O/c23_unique_timeline.err; echo "timeline rc=$?"; cut -c1-330 $O/c23_unique_timeline.jsonl
B200_UNIQUE_SMALL=16384 timeout 200 python tools/unique_timeline.py > $O/c23_unique_timeline_small.jsonl 2> $O/c23_unique_timeline_small.err; cut -c1-330 $O/c23_unique_timeline_small.jsonl

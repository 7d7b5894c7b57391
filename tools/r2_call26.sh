#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c26_bench_$name.json 2> $O/c26_bench_$name.err; }
run default X=1
run fork_start B200_LOOKAHEAD_FORK=start
run ublocks1 B200_LOOKAHEAD_UNIQUE_BLOCKS=1
run ublocks2 B200_LOOKAHEAD_UNIQUE_BLOCKS=2
run small4096 B200_UNIQUE_SMALL=4096
run nobranch B200_STEP_BRANCHES=0
run tile1 B200_TILE_CTAS=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c26_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 600 python bench.py > $O/c26_bench_full_default.json 2> $O/c26_bench_full_default.err; echo "full bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c26_launches_step.csv \
    python bench.py --profile-step --warmup 3 --no-cpu-baseline --api-steps 0 > $O/c26_launches_bench.log 2>&1; echo "ncu list rc=$?"
timeout 200 python tools/unique_timeline.py > $O/c26_unique_timeline.jsonl 2> $O/c26_unique_timeline.err; cut -c1-300 $O/c26_unique_timeline.jsonl

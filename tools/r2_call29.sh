#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --api-steps 0 > $O/c29_bench_$name.json 2> $O/c29_bench_$name.err; }
run default X=1
run fb5 B200_FLAT_BLOCKS=5
run fb4 B200_FLAT_BLOCKS=4
run fb3 B200_FLAT_BLOCKS=3
run fb5u1 B200_FLAT_BLOCKS=5 B200_FLAT_U=1
run fb10 B200_FLAT_BLOCKS=10
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c29_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
for fb in 8 5 4; do echo "FLAT_BLOCKS=$fb"; B200_FLAT_BLOCKS=$fb BK_ONLY=adam:8 BK_SIZES=122000,4000000 timeout 200 python bench_kernels.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(' ', d['kernel'], d.get('dim'), d.get('unique_ids'), round(d['us'],1), 'us', round(d['frac_of_peak'],3))
"; done

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29621 tests/mgpu_xchg_check.py > $O/c31_xchg_check_w$N.log 2>&1; echo "xchg check rc=$?"; tail -3 $O/c31_xchg_check_w$N.log
timeout 200 $TR --master-port 29622 tools/xchg_profile.py > $O/c31_xprof_fast_n$N.json 2> $O/c31_xprof_fast.err; echo "prof rc=$?"; tail -1 $O/c31_xprof_fast_n$N.json
timeout 420 $TR --master-port 29624 bench.py --gpus $N --steps 20 --warmup 5 > $O/c31_bench_n$N.json 2> $O/c31_bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), (d.get('parity_check') or {}).get('pull'), (d.get('parity_check') or {}).get('push'), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f,'ERR', e)
PY
tail -3 $O/c31_bench_n$N.err

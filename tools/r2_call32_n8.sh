#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29641 bench.py --gpus $N --steps 20 --warmup 5 > $O/c32_bench_n$N.json 2> $O/c32_bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/c32_bench_n$N.json').read().strip().splitlines()[-1])
    print(round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), d.get('parity_check'), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
except Exception as e: print('ERR', e)
PY
tail -3 $O/c32_bench_n$N.err
timeout 200 $TR --master-port 29642 tests/mgpu_xchg_check.py > $O/c32_xchg_check_w$N.log 2>&1; echo "xchg check rc=$?"; tail -2 $O/c32_xchg_check_w$N.log

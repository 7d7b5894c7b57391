#!/bin/bash
# one 8-GPU box: the N=8 bench line (with its parity self-check), the world-8 parity checks, configs 4 and 5
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > $O/c11_bench_n$N.json 2> $O/c11_bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/c11_bench_n$N.json').read().strip().splitlines()[-1])
    print(round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), d.get('parity_check'), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
except Exception as e: print('ERR', e)
PY
tail -3 $O/c11_bench_n$N.err
timeout 300 $TR --master-port 29612 tests/mgpu_xchg_check.py > $O/c11_xchg_check_w$N.log 2>&1; echo "xchg check rc=$?"; tail -4 $O/c11_xchg_check_w$N.log
timeout 300 $TR --master-port 29613 tests/mgpu_allreduce_check.py > $O/c11_allreduce_check_w$N.log 2>&1; echo "allreduce check rc=$?"; tail -4 $O/c11_allreduce_check_w$N.log
timeout 300 $TR --master-port 29614 tools/bench_allreduce.py > $O/c11_allreduce_n$N.json 2> $O/c11_allreduce_n$N.err; echo "allreduce bench rc=$?"; cat $O/c11_allreduce_n$N.json
timeout 420 $TR --master-port 29615 tools/bench_ftrl1b.py --rows 1000000000 > $O/c11_ftrl1b_n$N.json 2> $O/c11_ftrl1b_n$N.err; echo "ftrl rc=$?"; cat $O/c11_ftrl1b_n$N.json; tail -3 $O/c11_ftrl1b_n$N.err
timeout 300 $TR --master-port 29616 tests/mgpu_check.py > $O/c11_mgpu_check_w$N.log 2>&1; echo "mgpu check rc=$?"; tail -3 $O/c11_mgpu_check_w$N.log

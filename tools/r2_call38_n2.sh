#!/bin/bash
# round 2, call 38 (2 GPUs): the final tree through the driver's N=2 launch line, with the parity self-check
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29644 bench.py --gpus 2 --steps 20 --warmup 5 > $O/c38_bench_n2.json 2> $O/c38_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c38_bench_n2.json').read().strip().splitlines()[-1])
    print(round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), d.get('parity_check'), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
except Exception as e: print('ERR', e)
PY
tail -3 $O/c38_bench_n2.err

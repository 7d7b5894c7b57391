#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_layer_trainer.py -m gpu -x -q -k "multi_gpu" > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest.log
tail -8 gpurun_out/c8_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c8_bench_n2.json 2> gpurun_out/c8_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c8_bench_n2.json').read().strip().splitlines()[-1])
    print(round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), d.get('parity_check',{}).get('pull'), d.get('parity_check',{}).get('push'), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
except Exception as e: print('ERR', e)
PY
tail -5 gpurun_out/c8_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 tools/bench_ftrl1b.py --rows 200000000 > gpurun_out/c8_ftrl_n2.json 2> gpurun_out/c8_ftrl_n2.err; echo "ftrl rc=$?"; cat gpurun_out/c8_ftrl_n2.json; tail -3 gpurun_out/c8_ftrl_n2.err

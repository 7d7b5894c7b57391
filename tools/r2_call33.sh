#!/bin/bash
# round 2, call 33: the trainer's CUDA-graph minibatch, the golden-vector kernel test, bounded-unique range flag, N=1 bench with the graphed api leg
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_layer_trainer.py tests/test_gpu_parity.py -m gpu -q -x -k "cuda_graph or reference_kernel_vectors or unique_small or batched_lookups" > $O/c33_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c33_pytest.log
tail -15 $O/c33_pytest.log
timeout 400 python bench.py > $O/c33_bench.json 2> $O/c33_bench.err; echo "bench rc=$?"
tail -3 $O/c33_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c33_bench.json').read().strip().splitlines()[-1])
    print(round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1))
    print(json.dumps(d.get('api_path'), indent=1)[:2500])
except Exception as e: print('ERR', e)
PY

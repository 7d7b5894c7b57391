#!/bin/bash
# round 2, call 41: one bench line of the very last tree (bench.py touched after call 40)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 40 python bench.py --no-cpu-baseline --api-steps 0 --steps 10 > gpurun_out/c41_bench.json 2> gpurun_out/c41_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/c41_bench.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step']*1e3,1), round(d['value']/1e6,1), list(d['roofline'])[-3:])"

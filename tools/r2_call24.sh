#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
B200_UNIQUE_SMALL=16384 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_unique_small -f -o $O/c24_prof_unique_small \
    python bench.py --profile-step --warmup 3 --no-cpu-baseline --api-steps 0 --ids narrow > $O/c24_prof.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_unique -f -o $O/c22_prof_unique \
    python bench.py --profile-step --warmup 3 --no-cpu-baseline --api-steps 0 > $O/c22_prof.log 2>&1; echo "ncu rc=$?"

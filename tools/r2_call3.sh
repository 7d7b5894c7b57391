#!/bin/bash
# round 2, GPU call 3: flat kernels v2 (param-resident metadata), unique v2 (coalesced tiles), tile tower
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
tail -4 gpurun_out/c3_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/c3_bench.json; echo
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/c3_prof_step \
    python bench.py --profile-step --warmup 3 --no-cpu-baseline --tower tile > gpurun_out/c3_prof_bench.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench_kernels.py > gpurun_out/c3_kernels.jsonl 2> gpurun_out/c3_kernels.err; echo "kernels rc=$?"

#!/bin/bash
# round 2, GPU call 1: parity tests, N=1 bench, kernel roofline sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/c1_bench.json
timeout 600 python bench_kernels.py > gpurun_out/c1_kernels.jsonl 2> gpurun_out/c1_kernels.err; echo "kernels rc=$?"
for u in 1 2 4; do B200_FLAT_U=$u BK_ONLY=adam:8 timeout 300 python bench_kernels.py > gpurun_out/c1_kernels_u$u.jsonl 2>> gpurun_out/c1_kernels.err; done
timeout 300 python bench.py --steps 20 --warmup 5 --tower tile --no-cpu-baseline > gpurun_out/c1_bench_tile.json 2> gpurun_out/c1_bench_tile.err; echo "bench tile rc=$?"
tail -c 600 gpurun_out/c1_bench_tile.json

#!/bin/bash
# round 2, GPU call 2: new tests (features, tile tower), tile-tower A/B, ncu of one eager step, graph-timed kernel sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_features.py tests/test_gpu_layer_trainer.py -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -4 gpurun_out/c2_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --tower tile --no-cpu-baseline > gpurun_out/c2_bench_tile.json 2> gpurun_out/c2_bench_tile.err; echo "bench tile rc=$?"
tail -c 700 gpurun_out/c2_bench_tile.json; echo
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/c2_prof_step \
    python bench.py --profile-step --warmup 3 --no-cpu-baseline --tower tile > gpurun_out/c2_prof_bench.log 2>&1; echo "ncu rc=$?"
BK_ONLY=adam:8 timeout 300 python bench_kernels.py > gpurun_out/c2_kernels_adam8.jsonl 2> gpurun_out/c2_kernels.err; echo "kernels rc=$?"
cat gpurun_out/c2_kernels_adam8.jsonl | cut -c1-200
ls -la gpurun_out/*.ncu-rep

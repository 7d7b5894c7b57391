#!/bin/bash
# round 2, 2-GPU call: multi-GPU parity checks, N=2 bench with the parity self-check, fused allreduce
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c7_smi.txt
timeout 900 python -m pytest tests/test_gpu_layer_trainer.py -m gpu -x -q -k "multi_gpu" > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c7_pytest.log
tail -15 gpurun_out/c7_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c7_bench_n2.json 2> gpurun_out/c7_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/c7_bench_n2.json; echo; tail -5 gpurun_out/c7_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 tools/bench_allreduce.py > gpurun_out/c7_allreduce_n2.json 2> gpurun_out/c7_allreduce_n2.err; echo "allreduce rc=$?"
cat gpurun_out/c7_allreduce_n2.json; tail -5 gpurun_out/c7_allreduce_n2.err
timeout 200 python tools/bench_allreduce.py > gpurun_out/c7_allreduce_n1.json 2> gpurun_out/c7_allreduce_n1.err; cat gpurun_out/c7_allreduce_n1.json

#!/bin/bash
# round 2, call 39: full GPU suite on the final tree (trainer: one-off odd minibatches keep the graph, repeated new signature is recaptured)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/c39_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c39_pytest.log
tail -12 gpurun_out/c39_pytest.log

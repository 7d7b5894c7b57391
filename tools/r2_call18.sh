#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/e2e_probe.py > gpurun_out/c18_e2e_probe.json 2> gpurun_out/c18_e2e_probe.err; echo "rc=$?"; cat gpurun_out/c18_e2e_probe.json; tail -5 gpurun_out/c18_e2e_probe.err

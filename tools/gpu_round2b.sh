#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "resnet or recon or auroc or encoder or esvit or gmm_validator or boundary" > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_gpu_b.log | head -40
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 300 python tools/gpu_diag_configs.py > gpurun_out/other_configs.txt 2>&1; echo "configs rc=$?"; tail -12 gpurun_out/other_configs.txt

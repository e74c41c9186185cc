#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_esvit_gpu.py tests/test_encoder_gpu.py tests/test_linear_gpu.py -m gpu -q > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_gpu_i.log | head -20
timeout 300 python tools/gpu_diag_configs.py > gpurun_out/other_configs.txt 2>&1; echo "configs rc=$?"; grep -E "^config|^EsViT|^NF head|attention_t196|decoder alone" gpurun_out/other_configs.txt

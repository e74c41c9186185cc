#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_auroc_gpu.py tests/test_boundary_gpu.py tests/test_gmm_gpu.py -m gpu -q > gpurun_out/pytest_gpu_f.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_gpu_f.log | head -20
timeout 900 python bench.py --workload sweep --steps 5 --warmup 1 --sweep-light-warmup > gpurun_out/bench_sweep_n1.json 2> gpurun_out/bench_sweep.err; echo "sweep rc=$?"; tail -5 gpurun_out/bench_sweep.err
python - <<'PY'
import json
s=json.load(open("gpurun_out/bench_sweep_n1.json")); print("sweep N=1", round(s["value"]), "img/s e2e", round(s["e2e"]["value"]), "ms", round(s["ms_per_step"],1), round(s["e2e"]["ms_per_step"],1), "checksum", s["metrics_checksum"])
PY

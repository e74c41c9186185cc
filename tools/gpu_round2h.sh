#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_encoder_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_gpu_h.log | head -20
if grep -q "failed" gpurun_out/pytest_gpu_h.log; then exit 0; fi
for mode in vnat vt vnat vt; do
  if [ $mode = vt ]; then export VITAD_VNAT=0; else unset VITAD_VNAT; fi
  timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_h.json")); print("$mode", round(d["value"]), "img/s", round(d["ms_per_step"],4), "ms e2e", round(d["e2e"]["value"]), "bs1", d["latency_bs1_ms"]["cuda_graph_p50"], [(k["site"], k["us_per_launch"]) for k in d["kernels"] if "n2304" in k["site"] or "attention" in k["site"]])
PY
done

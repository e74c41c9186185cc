#!/bin/bash
# same-box A/B of one environment switch: gpu_ab.sh VAR  (VAR=0 is the "off" arm)
set -u
VAR=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_encoder_gpu.py -m gpu -q > gpurun_out/pytest_gpu_ab.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_gpu_ab.log | head -20
for mode in on off on off; do
  if [ $mode = off ]; then export $VAR=0; else unset $VAR; fi
  timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab.json")); print("$VAR $mode", round(d["value"]), "img/s", round(d["ms_per_step"],4), "ms e2e", round(d["e2e"]["value"]), "bs1", d["latency_bs1_ms"]["cuda_graph_p50"], [(k["site"], k["us_per_launch"]) for k in d["kernels"][1:6]])
PY
done

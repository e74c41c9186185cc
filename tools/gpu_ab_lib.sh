#!/bin/bash
# same-box A/B of two builds of the library: gpu_ab_lib.sh OTHER.so [pytest selection...]
set -u
OTHER=$1; shift
mkdir -p gpurun_out
if [ $# -gt 0 ]; then
  timeout 900 python -m pytest "$@" -m gpu -q -x > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"
  grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_ab.log | head -20
fi
for which in new other new other; do
  if [ $which = other ]; then export VITAD_LIB=$OTHER; else unset VITAD_LIB; fi
  timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab.json")); print("$which", round(d["value"]), "img/s", round(d["ms_per_step"],4), "ms e2e", round(d["e2e"]["value"]), "bs1", d["latency_bs1_ms"]["cuda_graph_p50"], [(k["site"], k["us_per_launch"]) for k in d["kernels"][1:6]])
PY
done

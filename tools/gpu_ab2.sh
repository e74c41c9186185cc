#!/bin/bash
# same-box A/B of two values of one environment variable: gpu_ab2.sh VAR A B
set -u
VAR=$1; A=$2; B=$3
mkdir -p gpurun_out
for val in $A $B $A $B; do
  export $VAR=$val
  timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab.json")); print("$VAR=$val", round(d["value"]), "img/s", round(d["ms_per_step"],4), "ms e2e", round(d["e2e"]["value"]), "bs1", d["latency_bs1_ms"]["cuda_graph_p50"], [(k["site"], k["us_per_launch"]) for k in d["kernels"][1:7] if "attention" in k["site"] or "gemm_ln_n768_k768" in k["site"]])
PY
done

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_encoder_gpu.py tests/test_graph_gpu.py tests/test_gmm_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_gpu_g.log | head -20
if grep -q "failed" gpurun_out/pytest_gpu_g.log; then exit 0; fi
for rep in 1 2; do
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_g.json")); print("vnat", round(d["value"]), "img/s", round(d["ms_per_step"],4), "ms e2e", round(d["e2e"]["value"]), "bs1", d["latency_bs1_ms"]["cuda_graph_p50"])
print([(k["site"], k["us_per_launch"]) for k in d["kernels"][:8]])
PY
VITAD_CTA_PAIR_VNAT_OFF=1 true
done

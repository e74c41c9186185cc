#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_gpu.py -m gpu -q -x -k "fused" > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest fused rc=$?"
grep -E "^E  |passed|failed|^FAILED|timeout|illegal|Error" gpurun_out/pytest_gpu_c.log | head -30
if grep -q "failed" gpurun_out/pytest_gpu_c.log; then exit 0; fi
timeout 900 python -m pytest tests -m gpu -q -k "encoder or auroc or gmm_validator or esvit or recon or graph" > gpurun_out/pytest_gpu_c2.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_gpu_c2.log | head -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
VITAD_LN_EPI_WARPS=16 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/bench_w16.json 2> gpurun_out/bench_w16.err; echo "bench16 rc=$?"; cut -c1-300 gpurun_out/bench_w16.json
VITAD_FUSED_LN=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; echo "bench unfused rc=$?"; cut -c1-300 gpurun_out/bench_unfused.json
python - <<'PY'
import json
for f in ("bench.json","bench_w16.json","bench_unfused.json"):
    try:
        d=json.load(open("gpurun_out/"+f)); print(f, round(d["value"]), "img/s", round(d["ms_per_step"],3), "ms", d["clocks"]); print([ (k["site"],k["launches_per_step"],k["us_per_launch"]) for k in d["kernels"][:12]])
    except Exception as e: print(f, e)
PY

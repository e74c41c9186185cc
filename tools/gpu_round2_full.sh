#!/bin/bash
# full GPU suite + smoke + headline + sweep (N=1)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_gpu.log | head -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --workload sweep --steps 3 --warmup 1 --sweep-light-warmup > gpurun_out/bench_sweep_n1.json 2> gpurun_out/bench_sweep.err; echo "sweep rc=$?"; tail -5 gpurun_out/bench_sweep.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("headline", round(d["value"]), "img/s", round(d["ms_per_step"],4), "ms; e2e", round(d["e2e"]["value"]), d["e2e"]["repetitions_ms"], "roofline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","peak","frac","peak_regime","frac_of_burst_peak","frac_of_sustained_peak","ms_per_launch","encoder_tflops")})
print("sustained", d.get("sustained")); print("bs1", d.get("latency_bs1_ms")); print("cpu", d.get("cpu_baseline")); print("clocks", d["clocks"], "launches", d["gpu_launches"])
for k in d["kernels"][:14]: print("  ", k)
s=json.load(open("gpurun_out/bench_sweep_n1.json")); print("sweep N=1", round(s["value"]), "img/s e2e", round(s["e2e"]["value"]), "ms", round(s["ms_per_step"],1), "checksum", s["metrics_checksum"])
PY

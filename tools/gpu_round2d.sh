#!/bin/bash
set -u
mkdir -p gpurun_out
VITAD_LIB=vit-ad_b200/lib/libvitad_tl.so timeout 300 python tools/gpu_timeline_ln.py 768 2>&1 | grep -v "^   unit" | head -40
timeout 600 python -m pytest tests/test_linear_gpu.py tests/test_encoder_gpu.py -m gpu -q -x -k "fused or batch_invariant or golden or oracle" > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_gpu_d.log | head
for mode in fused unfused fused unfused; do
  if [ $mode = unfused ]; then export VITAD_FUSED_LN=0; else unset VITAD_FUSED_LN; fi
  timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$mode.json")); print("$mode", round(d["value"]), "img/s", round(d["ms_per_step"],4), "ms e2e", round(d["e2e"]["value"]), "bs1", d.get("latency_bs1_ms"))
PY
done

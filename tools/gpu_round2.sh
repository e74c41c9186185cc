#!/bin/bash
# Round-2 GPU visit: parity tests (not -x: see every failure), smoke, headline bench, sweep workload.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --workload sweep --steps 3 --warmup 1 --sweep-light-warmup > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err; echo "sweep rc=$?"; cat gpurun_out/bench_sweep.json; tail -5 gpurun_out/bench_sweep.err

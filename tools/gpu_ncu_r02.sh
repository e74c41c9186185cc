#!/bin/bash
# Round-2 profiling visit (1 GPU): each target runs plain first (exit 0), then under ncu.
set -u
mkdir -p gpurun_out
timeout 300 python tools/prof_target.py > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_ln_kernel -s 23 -c 2 -o gpurun_out/prof_gemm_ln -f python tools/prof_target.py > gpurun_out/ncu_gemm_ln.log 2>&1; echo "ncu gemm_ln rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm3_tc_kernel|attention_kernel|gemm_ln_kernel" -s 66 -c 6 -o gpurun_out/prof_enc -f python tools/prof_target.py > gpurun_out/ncu_enc.log 2>&1; echo "ncu enc rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm4_tc_kernel -s 1 -c 1 -o gpurun_out/prof_mdn -f python tools/prof_target.py > gpurun_out/ncu_mdn.log 2>&1; echo "ncu mdn rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv

#!/bin/bash
# One GPU-box visit: parity tests, smoke, both bench arms, in-situ launch-site profile, ncu launch list and one
# full-set capture of the dominant kernel.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 300 python tools/gpu_profile_path.py > gpurun_out/profile_path.log 2>&1; cat gpurun_out/profile_path.log
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# full-set captures (second step of the target): fused GMM kernel, the four encoder GEMMs + attention + LayerNorm of one block
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm4_tc_kernel -s 1 -c 1 -o gpurun_out/prof_mdn -f python tools/prof_target.py > gpurun_out/ncu_mdn.log 2>&1; echo "ncu mdn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gemm3_tc_kernel|attention_kernel|layernorm_kernel" -s 60 -c 7 -o gpurun_out/prof_enc -f python tools/prof_target.py > gpurun_out/ncu_enc.log 2>&1; echo "ncu enc rc=$?"
fi
timeout 300 python tools/gpu_diag_configs.py > gpurun_out/other_configs.txt 2>&1; echo "configs rc=$?"
timeout 300 python tools/gpu_diag_resize.py > gpurun_out/resize_decoder.txt 2>&1; echo "resize rc=$?"
timeout 600 python tools/sweep.py > gpurun_out/sweep.json 2> gpurun_out/sweep.err; echo "sweep rc=$?"; cut -c1-300 gpurun_out/sweep.json

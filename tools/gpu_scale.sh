#!/bin/bash
# N-GPU visit: headline bench + sweep workload at N = $1 (torchrun, one rank per GPU)
set -u
N=$1
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 5 --sustained-seconds 0 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -2 gpurun_out/bench_n$N.err
timeout 900 $RUN bench.py --workload sweep --gpus $N --steps 3 --warmup 1 --sweep-light-warmup > gpurun_out/bench_sweep_n$N.json 2> gpurun_out/bench_sweep_n$N.err; echo "sweep N=$N rc=$?"; tail -3 gpurun_out/bench_sweep_n$N.err
if [ "$N" = "2" ]; then timeout 600 $RUN tools/check_gather.py > gpurun_out/check_gather_n2.txt 2>&1; echo "check_gather rc=$?"; grep -E "^[01] " gpurun_out/check_gather_n2.txt; fi
python - <<PY
import json
last=lambda f: json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
d=last("gpurun_out/bench_n$N.json"); print("headline N=$N", round(d["value"]), "img/s", round(d["ms_per_step"],4), "ms; e2e", round(d["e2e"]["value"]))
s=last("gpurun_out/bench_sweep_n$N.json"); print("sweep N=$N", round(s["value"]), "img/s e2e", round(s["e2e"]["value"]), "ms", round(s["ms_per_step"],1), "checksum", s["metrics_checksum"])
PY

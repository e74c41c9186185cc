#!/bin/bash
# the sweep workload alone on the N GPUs of this box: gpu_sweep_n.sh N [extra bench flags]
set -u
N=${1:-2}; shift || true
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload sweep --steps 5 --warmup 2 "$@" > gpurun_out/bench_sweep_n$N.json 2> gpurun_out/bench_sweep_n$N.err; echo "sweep N=$N rc=$?"
tail -3 gpurun_out/bench_sweep_n$N.err
python - <<PY
import json
def last_json(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])
s = last_json("gpurun_out/bench_sweep_n$N.json")
print("sweep    N=$N", round(s["value"]), "img/s  e2e", round(s["e2e"]["value"]), " ms/sweep", round(s["ms_per_step"], 1), " checksum", s.get("metrics_checksum"))
PY

#!/bin/bash
# The two streamline lines of tools/gpu_multi.sh only (weak scaling by frames, strong scaling by seeds) -- re-measured after
# the last change to the streamline kernel so that every N comes from the same build.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29542 bench.py --gpus $N > gpurun_out/r2_bench_topo3a_${N}gpu.log 2>gpurun_out/r2_bench_topo3a_${N}gpu.err
timeout 400 $TR --master-port 29543 bench.py --gpus $N --workload md1m --split seeds --steps 10 > gpurun_out/r2_split_seeds_${N}gpu.log 2>gpurun_out/r2_split_seeds_${N}gpu.err
python - <<PY
import json
for f in ("gpurun_out/r2_bench_topo3a_${N}gpu.log", "gpurun_out/r2_split_seeds_${N}gpu.log"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.4e e2e %.4e ms %.3f parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("parity_checked")))
PY

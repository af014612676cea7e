#!/bin/bash
# 8-GPU extras of round 2: BASELINE configs[3] for real (two passes), and the weak-scaling line with the histogram gathers
# per frame (default), batched 4 frames per call, and disabled (diagnostic: what the exchange costs).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29545 tools/config4.py --frames 1000 --axis 100 --out gpurun_out/r2_config4_8gpu.json > gpurun_out/r2_config4_8gpu.log 2>&1
tail -c 400 gpurun_out/r2_config4_8gpu.log
for g in 1 4 0; do
  timeout 300 $TR --master-port 2955$g bench.py --gpus 8 --gather-every $g > gpurun_out/r2_bench_topo3a_8gpu_gather$g.log 2>gpurun_out/r2_bench_topo3a_8gpu_gather$g.err
done
python - <<'PY'
import json
for g in (1, 4, 0):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_topo3a_8gpu_gather{g}.log").read().strip().splitlines()[-1])
        print(g, "value %.4e ms %.4f parity %s" % (d["value"], d["ms_per_step"], d["parity_checked"]), d["step_ms_by_rank"]["median"], d["step_ms_by_rank"]["kernel_mean"])
    except Exception as e:
        print(g, "failed", e)
PY

#!/bin/bash
# Multi-GPU evidence of round 2, one call per N:  gpurun --gpus N -- 'bash tools/gpu_multi.sh N [full|lite|extra]'
#   (default) nccl_check (sharding + trajectory pipeline over NCCL, bit for bit against one GPU), bench.py default workload at N
#             (weak scaling, gather ring, parity_checked), strong scaling: one md1m frame by seeds, one 464^3 x 100k mesh by slabs
#             (gather inside the timed bracket), a small trajectory through tools/config4.py
#   full      the same, with BASELINE configs[3] for real (1000 frames x 1M seeds) through pycpet_b200.trajectory
#   lite      only the two streamline lines (weak scaling by frames, strong scaling by seeds)
#   extra     BASELINE configs[3] (two passes) and the weak-scaling line with the histogram gathers per frame, batched 4 frames
#             per call, and disabled (diagnostic: what the exchange costs)
N=${1:-2}
MODE=${2:-}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$MODE" = "extra" ]; then
  timeout 600 $TR --master-port 29545 tools/config4.py --frames 1000 --axis 100 --out gpurun_out/r2_config4_${N}gpu.json > gpurun_out/r2_config4_${N}gpu.log 2>&1
  tail -c 400 gpurun_out/r2_config4_${N}gpu.log
  for g in 1 4 0; do
    timeout 300 $TR --master-port 2955$g bench.py --gpus $N --gather-every $g > gpurun_out/r2_bench_topo3a_${N}gpu_gather$g.log 2>gpurun_out/r2_bench_topo3a_${N}gpu_gather$g.err
  done
else
  if [ "$MODE" != "lite" ]; then
    timeout 400 $TR --master-port 29541 tests/nccl_check.py > gpurun_out/r2_nccl_check_${N}gpu.log 2>&1
    grep "nccl_check" gpurun_out/r2_nccl_check_${N}gpu.log | tail -14
  fi
  timeout 400 $TR --master-port 29542 bench.py --gpus $N > gpurun_out/r2_bench_topo3a_${N}gpu.log 2>gpurun_out/r2_bench_topo3a_${N}gpu.err
  timeout 400 $TR --master-port 29543 bench.py --gpus $N --workload md1m --split seeds --steps 10 > gpurun_out/r2_split_seeds_${N}gpu.log 2>gpurun_out/r2_split_seeds_${N}gpu.err
  if [ "$MODE" != "lite" ]; then
    timeout 600 $TR --master-port 29544 bench.py --gpus $N --workload volume464 --split slab --steps 2 > gpurun_out/r2_split_slab_${N}gpu.log 2>gpurun_out/r2_split_slab_${N}gpu.err
    if [ "$MODE" = "full" ]; then
      timeout 900 $TR --master-port 29545 tools/config4.py --frames 1000 --axis 100 --out gpurun_out/r2_config4_${N}gpu.json > gpurun_out/r2_config4_${N}gpu.log 2>&1
    else
      timeout 600 $TR --master-port 29545 tools/config4.py --frames 24 --axis 47 --out gpurun_out/r2_config4_small_${N}gpu.json > gpurun_out/r2_config4_small_${N}gpu.log 2>&1
    fi
    tail -c 2500 gpurun_out/r2_config4_*${N}gpu.log
  fi
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_*_${N}gpu*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    if "value" in d:
        print(f, "value %.4e e2e %.4e ms %.3f parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("parity_checked")), d.get("limiter", d.get("step_ms_by_rank")))
PY

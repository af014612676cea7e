#!/bin/bash
# Round 2, second GPU call (1 GPU): new parity tests, the default bench line with sustained/others, the
# split modes and the config-4 runner at reduced size (code-path validation), nccl_check with one rank.
mkdir -p gpurun_out
{
  echo "== pytest -m gpu"
  timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
} > gpurun_out/r2b_tests.log 2>&1
tail -8 gpurun_out/r2b_tests.log
timeout 600 python bench.py > gpurun_out/r2b_bench_topo3a.log 2>gpurun_out/r2b_bench_topo3a.err
tail -c 600 gpurun_out/r2b_bench_topo3a.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2b_bench_topo3a.log").read().strip().splitlines()[-1])
    print("value %.4e e2e %.4e frac_nominal %.4f parity %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac_nominal"], d["parity_checked"]))
    print("sustained", json.dumps(d.get("sustained"))[:600])
    for k, o in d.get("others", {}).items():
        print(k, "%.4e e2e %.4e" % (o["value"], o["e2e"]["value"]), json.dumps(o["roofline"])[:300])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29541 \
  tests/nccl_check.py > gpurun_out/r2b_nccl1.log 2>&1; tail -14 gpurun_out/r2b_nccl1.log
timeout 300 python tools/config4.py --frames 16 --axis 47 --out gpurun_out/r2b_config4_small.json > gpurun_out/r2b_config4.log 2>&1
tail -c 1500 gpurun_out/r2b_config4.log
timeout 300 python bench.py --workload md1m --split seeds --steps 5 > gpurun_out/r2b_split_seeds_1.log 2>&1; tail -c 1200 gpurun_out/r2b_split_seeds_1.log
timeout 400 python bench.py --workload volume464 --split slab --steps 2 > gpurun_out/r2b_split_slab_1.log 2>&1; tail -c 1200 gpurun_out/r2b_split_slab_1.log

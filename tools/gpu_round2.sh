#!/bin/bash
# Round 2 single-GPU artefacts with the points-packed streamline kernel: tests, smoke, the default bench line
# (sustained + others), the reference arm, launch list, full-set ncu capture of the integrator, compute-sanitizer.
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv
  echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
} > gpurun_out/round2_tests.log 2>&1
tail -8 gpurun_out/round2_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/round2_bench_topo3a.log 2>gpurun_out/round2_bench_topo3a.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/round2_bench_reference.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/round2_bench_topo3a.log").read().strip().splitlines()[-1])
print("value %.4e e2e %.4e frac_nominal %.4f kernel %s parity %s launches %d" % (d["value"], d["e2e"]["value"], d["roofline"]["frac_nominal"], d["roofline"]["kernel"], d["parity_checked"], d["gpu_launches"]))
print("sustained", json.dumps(d.get("sustained"))[:400])
for k, o in d.get("others", {}).items():
    print(k, "%.4e e2e %.4e frac_nominal %.4f" % (o["value"], o["e2e"]["value"], o["roofline"]["frac_nominal"]))
r = json.loads(open("gpurun_out/round2_bench_reference.log").read().strip().splitlines()[-1])
print("reference %.4e cores %s" % (r["value"], r["cpu_baseline"]["cores"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/round2_launches.csv python bench.py --steps 2 --warmup 3 --sustained 0 --others "" > gpurun_out/round2_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k2p_topo_kernel' -c 1 \
  -o gpurun_out/round2_k2p python tools/prof_k2w.py 2 > gpurun_out/round2_prof.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2p_topo_kernel|k1_lattice|k1_grid_kernel' \
  -o gpurun_out/round2_workloads python tools/prof_workloads.py > gpurun_out/round2_prof_workloads.log 2>&1
timeout 300 python bench.py --workload md1m --split seeds --steps 10 > gpurun_out/r2_split_seeds_1gpu.log 2>gpurun_out/r2_split_seeds_1gpu.err
# (the slab split at N = 1 is part of tools/gpu_multi.sh-style runs: bench.py --workload volume464 --split slab --steps 2)
timeout 300 python tools/lattice_nodes_ab.py > gpurun_out/r2_lattice_nodes.log 2>&1
timeout 300 python tools/esp_lattice_quick.py > gpurun_out/r2_esp_lattice_quick.log 2>&1
timeout 300 python tools/k2_ab.py > gpurun_out/r2_k2_ab.log 2>&1
{
  for t in memcheck racecheck initcheck synccheck; do
    echo "== $t"; timeout 900 compute-sanitizer --tool $t python tools/sanitize_target.py 2>&1 | grep -E "sanitize target done|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" | head -8
  done
} > gpurun_out/round2_compute_sanitizer.txt 2>&1
cat gpurun_out/round2_compute_sanitizer.txt
ls -la gpurun_out | tail -12

#!/bin/bash
# One gpurun call that refreshes every single-GPU artefact of a round (about 12 GPU-minutes):
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh round2'
#
# then here:  python tools/results_md.py round2 ; python tools/ncu_summary.py ... (see profiles/README.md)
# Every step runs under its own timeout so that a hang costs that step, not the box.
R=${1:-round2}
mkdir -p gpurun_out
{
  echo "== nvidia-smi"; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv
  echo "== pytest -m gpu"
  timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
} > gpurun_out/${R}_tests.log 2>&1
tail -12 gpurun_out/${R}_tests.log

timeout 400 python bench.py > gpurun_out/bench_topo3a_1gpu.log 2>gpurun_out/bench_topo3a_1gpu.err
for w in md1m topo_fine volume esp101 volume2a; do
  timeout 400 python bench.py --workload $w > gpurun_out/bench_$w.log 2>gpurun_out/bench_$w.err
done
timeout 400 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1
for f in gpurun_out/bench_*.log; do echo "$f: $(tail -1 $f | cut -c1-220)"; done

# launch list of the default bench command (numbers printed under ncu are never bench values)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
# one full-set capture of the dominant kernel of every workload (DRAM traffic, pipe utilisation, source page)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2w_topo_kernel|k1_lattice_kernel|k1_grid_kernel' \
  -o gpurun_out/${R}_workloads python tools/prof_workloads.py > gpurun_out/prof_workloads.log 2>&1
ls -la gpurun_out | tail -20

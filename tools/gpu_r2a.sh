#!/bin/bash
# Round 2, first GPU call: parity of the current tree, A/B of the two streamline kernels, inner-loop
# microbenchmarks, the default bench line and one full-set ncu capture of the hybrid integrator.
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv
  echo "== pytest -m gpu"
  timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
} > gpurun_out/r2a_tests.log 2>&1
tail -12 gpurun_out/r2a_tests.log
timeout 600 python tools/k2_ab.py > gpurun_out/r2a_k2_ab.log 2>&1
tail -30 gpurun_out/r2a_k2_ab.log
timeout 300 tools/ubench2 "" 1 > gpurun_out/r2a_ubench2.log 2>&1
cat gpurun_out/r2a_ubench2.log
timeout 400 python bench.py > gpurun_out/r2a_bench_topo3a.log 2>gpurun_out/r2a_bench_topo3a.err
tail -1 gpurun_out/r2a_bench_topo3a.log | cut -c1-1500
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k2x_topo_kernel' -c 1 \
  -o gpurun_out/r2a_k2x python tools/prof_k2w.py 2 > gpurun_out/r2a_prof.log 2>&1
ls -la gpurun_out | tail

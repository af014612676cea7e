#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python bench.py --workload md1m --split seeds --steps 10 > gpurun_out/r2_split_seeds_1gpu.log 2>gpurun_out/r2_split_seeds_1gpu.err; tail -c 400 gpurun_out/r2_split_seeds_1gpu.log
timeout 400 python bench.py --workload volume464 --split slab --steps 2 > gpurun_out/r2_split_slab_1gpu.log 2>gpurun_out/r2_split_slab_1gpu.err; tail -c 400 gpurun_out/r2_split_slab_1gpu.log
timeout 300 python tools/esp_lattice_quick.py > gpurun_out/r2_esp_lattice_quick.log 2>&1; cat gpurun_out/r2_esp_lattice_quick.log | cut -c1-220

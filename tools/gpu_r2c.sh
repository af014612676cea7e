#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 600 python tools/k2_ab.py '{"k2_cap": 4}' > gpurun_out/r2c_k2_ab.log 2>&1
grep -v direct_vs gpurun_out/r2c_k2_ab.log | cut -c1-260

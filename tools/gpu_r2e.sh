#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25
timeout 600 python tools/k2_ab.py '{"k2_form": 3, "k2_unroll": 4}' '{"k2_form": 3, "k2_unroll": 8}' > gpurun_out/r2e_k2_ab.log 2>&1
grep -v direct_vs gpurun_out/r2e_k2_ab.log | cut -c1-250
grep direct_vs gpurun_out/r2e_k2_ab.log

set -x
for tool in memcheck racecheck initcheck synccheck; do echo "== $tool"; timeout 600 compute-sanitizer --tool $tool python tools/sanitize_target.py 2>&1 | grep -v "^=========$" | tail -4; done > gpurun_out/sanitizer.log 2>&1
cat gpurun_out/sanitizer.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/launches_full*.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 700 --csv --log-file gpurun_out/launches_full.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-clocks > gpurun_out/ncu_bench_full.log 2>&1
tail -1 gpurun_out/ncu_bench_full.log | cut -c1-200
wc -l gpurun_out/launches_full*.csv

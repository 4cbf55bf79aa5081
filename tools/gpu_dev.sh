#!/bin/bash
mkdir -p gpurun_out
nproc; grep -m1 "model name" /proc/cpuinfo
( time timeout 1200 python bench.py --steps 5 --warmup 3 ) 2>&1 | tail -5 | tee gpurun_out/bench_default.log | cut -c1-2500
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) 2>&1 | tail -5 | tee gpurun_out/bench_reference.log | cut -c1-1200

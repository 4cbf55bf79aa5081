#!/bin/bash
# Dev loop on the GPU box: kernel tests in separate processes (a trapped kernel poisons the context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/smi.txt
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 | tee gpurun_out/test_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
timeout 600 python bench.py --steps 5 --warmup 3 --fast --no-cpu 2>&1 | tail -5 | tee gpurun_out/bench_fast.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log

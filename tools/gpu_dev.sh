#!/bin/bash
# Dev loop on the GPU box: kernel tests in separate processes (a trapped kernel poisons the context).
mkdir -p gpurun_out
timeout 300 python tools/gemm_probe.py 2>&1 | tail -20 | tee gpurun_out/gemm_probe.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30 | tee gpurun_out/test_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log

#!/bin/bash
# Dev loop on the GPU box: kernel tests in separate processes (a trapped kernel poisons the context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/smi.txt
timeout 300 python tools/gemm_probe.py 2>&1 | tee gpurun_out/gemm_probe.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not conv_gemm and not linear and not stem and not fpn" 2>&1 | tail -30 | tee gpurun_out/test_misc.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv_gemm or linear or stem or fpn" 2>&1 | tail -40 | tee gpurun_out/test_gemm.log

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roi_heads.py -q -m gpu 2>&1 | tail -60 | tee gpurun_out/test_heads.log

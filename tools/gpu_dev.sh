#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rpn.py -q -m gpu 2>&1 | tail -40 | tee gpurun_out/test_rpn.log

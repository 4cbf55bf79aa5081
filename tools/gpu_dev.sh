#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_runner.py -q -m gpu 2>&1 | grep -E "^E   |passed|failed|^FAILED|Error" | cut -c1-300

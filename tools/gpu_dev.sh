#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_edge_cases.py -q -m gpu 2>&1 | grep -E "^E   |passed|failed|^FAILED|Error" | cut -c1-300 | tee gpurun_out/test_edge.log

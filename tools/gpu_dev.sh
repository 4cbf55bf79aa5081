#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('full: img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'gemm_ms', round(d['roofline']['kernel_ms_per_step'],2), 'TF', round(d['roofline']['achieved'],1))"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu 2>&1 | tail -1

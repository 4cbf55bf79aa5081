#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_probe.py 2>&1 | grep -c "max_err=0 nan=0" | tee gpurun_out/gemm_probe.log
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | grep -E "^E   |passed|failed|^FAILED" | cut -c1-300 | tee gpurun_out/test_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('full: img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'gemm_ms', round(d['roofline']['kernel_ms_per_step'],2), 'TF', round(d['roofline']['achieved'],1), 'words', d['config']['words_per_step'])"
timeout 900 python bench.py --workload backbone_bs8 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('backbone: img/s', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'gemm_ms', round(d['roofline']['kernel_ms_per_step'],2), 'TF', round(d['roofline']['achieved'],1))"

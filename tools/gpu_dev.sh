#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gemm_probe.py 2>&1 | grep -c "max_err=0 nan=0"
GLASS_PAIR_MODE=2 timeout 120 python tools/gemm_probe.py 2>&1 | grep -c "max_err=0 nan=0"
for tm in 1 0; do
GLASS_TAP_MODE=$tm timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('tap_mode $tm full: img/s', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'gemm_ms', round(d['roofline']['kernel_ms_per_step'],2))"
done
GLASS_TAP_MODE=0 timeout 900 python bench.py --workload backbone_bs8 --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('tap_mode 0 backbone: img/s', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'gemm_ms', round(d['roofline']['kernel_ms_per_step'],2))"
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | grep -E "^E   |passed|failed|^FAILED" | cut -c1-300 | tee gpurun_out/test_gpu.log

#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gemm_probe.py 2>&1 | grep -c "max_err=0 nan=0"
for rep in 1 2; do
for lib in libglass_b200_base.so libglass_b200.so; do
GLASS_B200_LIB=$PWD/glass_text_spotting_b200/_lib/$lib timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$lib full: img/s', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'gemm_ms', round(d['roofline']['kernel_ms_per_step'],2))"
done
done
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | grep -E "^E   |passed|failed|^FAILED" | cut -c1-300 | tee gpurun_out/test_gpu.log

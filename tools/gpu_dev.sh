#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for lib in libglass_b200.so libglass_b200_alt.so; do
GLASS_B200_LIB=$PWD/glass_text_spotting_b200/_lib/$lib timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$lib full: img/s', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'gemm_ms', round(d['roofline']['kernel_ms_per_step'],2), d['clocks'])"
done
done

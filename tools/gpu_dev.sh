#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload roialign_512 --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_roialign.json | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_roi_heads.py -q -m gpu 2>&1 | tail -1

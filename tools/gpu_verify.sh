#!/bin/bash
# One GPU-box call: parity tests, then the three bench workloads.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/test_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_full.err | tail -1 | tee gpurun_out/bench_full.json
timeout 600 python bench.py --workload backbone_bs8 --steps 10 --warmup 3 2>gpurun_out/bench_bb.err | tail -1 | tee gpurun_out/bench_bb.json
timeout 300 python bench.py --workload roialign_512 --steps 20 --warmup 3 2>gpurun_out/bench_roi.err | tail -1 | tee gpurun_out/bench_roi.json

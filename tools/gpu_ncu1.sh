#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:roi_align_rotated_split8 -s 3 -c 1 -o gpurun_out/ncu_roi -f python bench.py --workload roialign_512 --steps 2 --warmup 3 > gpurun_out/ncu_roi.log 2>&1
$NCU -k regex:conv_gemm -s 4 -c 1 -o gpurun_out/ncu_bb_conv3 -f python bench.py --workload backbone_bs8 --steps 1 --warmup 1 --no-cpu --no-clocks > gpurun_out/ncu_bb.log 2>&1
$NCU -k regex:conv_gemm -s 76 -c 2 -o gpurun_out/ncu_local01 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-clocks > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep

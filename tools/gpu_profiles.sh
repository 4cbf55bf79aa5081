#!/bin/bash
# Round profiles: launch list of one full_bs4 step, DRAM traffic of every conv GEMM launch of a step, full captures of
# the MMA-bound GEMM, the HBM-bound GEMM and the RoIAlign kernel.  Outputs in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-clocks"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launch_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.md
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:conv_gemm -s 460 -c 115 --csv --log-file gpurun_out/gemm_traffic.csv $B > gpurun_out/gemm_traffic.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 436 -c 2 -o gpurun_out/ncu_gemm_l3 -f $B > gpurun_out/ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_rotated_split8 -s 3 -c 1 -o gpurun_out/ncu_roi_final -f python bench.py --workload roialign_512 --steps 8 --warmup 3 --no-clocks > gpurun_out/ncu_roi.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv

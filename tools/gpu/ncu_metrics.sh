#!/bin/bash
# round profiles: launch list of one eager full_bs4 step, per-launch GEMM metrics (time, DRAM bytes, tensor pipe) of one step
# of full_bs4 and backbone_bs8, the RoIAlign launch's DRAM bytes, and --set full captures of an MMA-bound GEMM, an
# HBM-bound GEMM with residual (TMA epilogue) and the RoIAlign kernel.   usage: ncu_metrics.sh <tag>
T=${1:-prof}
O=gpurun_out/$T
mkdir -p $O
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
F="python bench.py --steps 1 --warmup 1 --no-cpu --no-clocks --no-submetrics --no-graph"
B="python bench.py --workload backbone_bs8 --steps 1 --warmup 1 --no-cpu --no-clocks"
R="python bench.py --workload roialign_512 --steps 8 --warmup 3 --no-clocks"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 900 --csv --log-file $O/launches.csv $F > $O/launch_bench.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches_summary.md 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:conv_gemm -s 116 -c 116 --csv --log-file $O/full_gemm_metrics.csv $F > $O/full_gemm.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:conv_gemm -s 61 -c 61 --csv --log-file $O/bb_gemm_metrics.csv $B > $O/bb_gemm.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:roi_align_rotated_split8 -s 4 -c 8 --csv --log-file $O/roi_metrics.csv $R > $O/roi.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 206 -c 2 -o $O/ncu_gemm_l3 -f $F > $O/ncu_gemm_l3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 65 -c 1 -o $O/ncu_bb_res2conv3 -f $B > $O/ncu_bb.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_rotated_split8 -s 5 -c 1 -o $O/ncu_roi -f $R > $O/ncu_roi.log 2>&1
ls -la $O; head -30 $O/launches_summary.md

#!/bin/bash
# round 2, call 4: strict GPU suite; ncu --set full of two HBM-bound GEMM layers (res2 shortcut 64->256 no residual, res2 conv3 + residual)
# and the launch list of one eager step
mkdir -p gpurun_out/r2c4
O=gpurun_out/r2c4
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | cut -c1-300 > $O/tests_strict.log
B="python bench.py --workload backbone_bs8 --steps 1 --warmup 1 --no-cpu --no-clocks"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 63 -c 3 -o $O/ncu_bb_hbm -f $B > $O/ncu_bb.log 2>&1
F="python bench.py --steps 1 --warmup 1 --no-cpu --no-clocks --no-submetrics --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 900 --csv --log-file $O/launches.csv $F > $O/launch_bench.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches_summary.md 2>&1
tail -3 $O/tests_strict.log; ls -la $O; head -40 $O/launches_summary.md

#!/bin/bash
# recognizer-side A/B: parity tests of the recurrent kernels + a launch list of one step.  usage: rec.sh <tag>
T=${1:-rec}
O=gpurun_out/$T
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_roi_heads.py tests/test_gpu_fullsize_parity.py tests/test_gpu_e2e.py -q -m gpu 2>&1 | tail -8 | cut -c1-300 > $O/tests.log
cp gpurun_out/parity_report.json $O/ 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-submetrics 2>$O/bench.err | tail -1 > $O/bench.json
F="python bench.py --steps 1 --warmup 1 --no-cpu --no-clocks --no-submetrics --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 900 --csv --log-file $O/launches.csv $F > $O/launch_bench.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches_summary.md 2>&1
tail -3 $O/tests.log
python - <<PY
import json
try:
    d=json.load(open("$O/bench.json")); print("bench", round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3))
except Exception as e: print("bench ERR", e)
PY
sed -n 1,16p $O/launches_summary.md

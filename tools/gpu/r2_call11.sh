#!/bin/bash
O=gpurun_out/r2c11
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_roi_heads.py tests/test_gpu_fullsize_parity.py tests/test_gpu_e2e.py -q -m gpu -x 2>&1 | tail -5 | cut -c1-300 > $O/tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-submetrics 2>$O/bench.err | tail -1 > $O/bench.json
F="python bench.py --steps 1 --warmup 1 --no-cpu --no-clocks --no-submetrics --no-graph"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lstm_cluster|aster_decode" -s 3 -c 3 -o $O/ncu_rec -f $F > $O/ncu_rec.log 2>&1
tail -3 $O/tests.log
python - <<PY
import json
try:
    d=json.load(open("$O/bench.json")); print("bench", round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3))
except Exception as e: print("bench ERR", e)
PY
tail -2 $O/bench.err; ls -la $O

#!/bin/bash
# quick A/B: GEMM-level parity tests, per-layer tables, one bench line.  usage: quick.sh <tag>
T=${1:-quick}
O=gpurun_out/$T
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backbone.py tests/test_gpu_device_counts.py -q -m gpu -x 2>&1 | tail -4 | cut -c1-300 > $O/tests.log
timeout 200 python tools/layer_profile.py full_bs4 > $O/layers_full.txt 2>&1
timeout 200 python tools/layer_profile.py backbone_bs8 > $O/layers_bb.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-submetrics 2>$O/bench.err | tail -1 > $O/bench.json
tail -2 $O/tests.log; head -1 $O/layers_full.txt; head -1 $O/layers_bb.txt
python - <<PY
import json
try:
    d=json.load(open("$O/bench.json")); print("bench", round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "gemm_ms", round(d["roofline"]["kernel_ms_per_step"],2), "frac", round(d["roofline"]["frac"],4))
except Exception as e: print("bench ERR", e)
PY
tail -3 $O/bench.err

#!/bin/bash
# full check: strict GPU suite, smoke, bench line (with submetrics + CPU leg).  usage: full.sh <tag>
T=${1:-full}
O=gpurun_out/$T
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 | cut -c1-300 > $O/tests.log
cp gpurun_out/parity_report.json $O/ 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 2>$O/bench.err | tail -1 > $O/bench.json
tail -3 $O/tests.log; tail -1 $O/smoke.log
python - <<PY
import json
try:
    d=json.load(open("$O/bench.json")); print("bench", round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "bb issued", round(d["submetrics"]["backbone"]["issued_frac"],3), "roi", round(d["submetrics"]["roialign"]["frac_algorithmic"],3), "cpu", d["cpu_baseline"]["value"])
except Exception as e: print("bench ERR", e)
PY
tail -3 $O/bench.err

#!/bin/bash
# round 2, call 3: ABI v2 (device-side counts, plans, graph step): GPU suite in report-only mode + strict, bench
mkdir -p gpurun_out/r2c3
O=gpurun_out/r2c3
GLASS_PARITY_REPORT_ONLY=1 timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -60 | cut -c1-600 > $O/tests_report_only.log
cp gpurun_out/parity_report.json $O/parity_report.json 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-submetrics 2>$O/bench_graph.err | tail -1 > $O/bench_graph.json
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-submetrics --no-graph 2>$O/bench_eager.err | tail -1 > $O/bench_eager.json
GLASS_PLAN_CACHE=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-submetrics --no-graph 2>$O/bench_eager_noplan.err | tail -1 > $O/bench_eager_noplan.json
tail -25 $O/tests_report_only.log; tail -2 $O/smoke.log
for f in $O/bench_*.json; do python - <<PY
import json
try:
    d=json.load(open("$f")); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], d["gpu_launches"], d["run"].get("words_per_step"))
except Exception as e: print("$f", "ERR", e)
PY
done
tail -5 $O/bench_graph.err

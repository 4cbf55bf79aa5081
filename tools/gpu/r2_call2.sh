#!/bin/bash
# round 2, call 2: whole GPU suite in report-only mode (status of every tap vs the literal tolerance), smoke, bench lines
mkdir -p gpurun_out/r2c2
O=gpurun_out/r2c2
GLASS_PARITY_REPORT_ONLY=1 timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 | cut -c1-400 > $O/tests_report_only.log
cp gpurun_out/parity_report.json $O/parity_report.json 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 2>$O/bench_full.err | tail -1 > $O/bench_full.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>$O/bench_ref.err | tail -1 > $O/bench_ref.json
tail -5 $O/tests_report_only.log; tail -2 $O/smoke.log; cut -c1-600 $O/bench_full.json; tail -3 $O/bench_full.err; cut -c1-300 $O/bench_ref.json

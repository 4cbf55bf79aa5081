#!/bin/bash
# End-of-round verification on one GPU: parity tests, smoke, every bench workload (JSON lines into gpurun_out/final_*.json)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/final_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/final_full.err | tail -1 > gpurun_out/final_full.json
timeout 600 python bench.py --workload backbone_bs8 --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final_bb.json
timeout 300 python bench.py --workload roialign_512 --steps 200 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final_roi.json
timeout 300 python bench.py --workload postprocess_bs4 --steps 50 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final_pp.json
timeout 300 python bench.py --workload totaltext_loop --steps 10 --warmup 3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/final_tt.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/final_ref.json
for f in full bb roi pp tt ref; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/final_{f}.json"))
    print(f, round(d["value"],2), d["unit"], "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"), "frac", d.get("roofline",{}).get("frac"), "cpu", d.get("cpu_baseline",{}).get("value"), "launches", d.get("gpu_launches"))
except Exception as e:
    print(f, "FAILED", e)
PY
done

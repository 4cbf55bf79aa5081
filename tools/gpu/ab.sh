#!/bin/bash
# same-box A/B of environment switches: ab.sh <tag> "<ENV=a>" "<ENV=b>" ...   (each variant benched twice, interleaved)
T=$1; shift
O=gpurun_out/$T; mkdir -p $O
for rep in 1 2; do
  i=0
  for v in "$@"; do
    i=$((i+1))
    env $v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-submetrics 2>>$O/err.log | tail -1 > $O/v${i}_r${rep}.json
    python - <<PY
import json
try:
    d=json.load(open("$O/v${i}_r${rep}.json")); print("$v", "rep$rep", round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3))
except Exception as e: print("$v", "ERR", e)
PY
  done
done
tail -3 $O/err.log

#!/bin/bash
# 8-GPU evidence: the scaling bench (full_bs4) and the TotalText-shape loop (configs[4]) with the all-gather timed
O=gpurun_out/${1:-n8}; mkdir -p $O
N=${2:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 2>$O/full.err | tail -1 > $O/full_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --workload totaltext_loop --gpus $N --steps 20 --warmup 5 2>$O/tt.err | tail -1 > $O/totaltext_n$N.json
for f in $O/full.err $O/tt.err; do tail -n 2 $f; done
python - <<PY
import json
for f in ("$O/full_n$N.json","$O/totaltext_n$N.json"):
    try:
        d=json.load(open(f)); print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), {k:v for k,v in d["run"].items() if "allgather" in k})
    except Exception as e: print(f,"ERR",e)
PY

#!/bin/bash
# round 2, call 1: validate / time the two opt-in recurrent kernels left from round 1, baseline bench, per-layer GEMM table
mkdir -p gpurun_out/r2c1
O=gpurun_out/r2c1
timeout 200 python tools/lstm_ab.py > $O/lstm_ab.txt 2>&1
GLASS_TEST_OPTIN=1 timeout 500 python -m pytest tests/test_gpu_decoder_pre.py tests/test_gpu_lstm_cluster.py -q -m gpu -x 2>&1 | tail -15 | cut -c1-400 > $O/optin_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>$O/bench_base.err | tail -1 > $O/bench_base.json
GLASS_LSTM_CLUSTER=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>$O/bench_lstmc.err | tail -1 > $O/bench_lstmc.json
GLASS_DEC_PRE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>$O/bench_decpre.err | tail -1 > $O/bench_decpre.json
GLASS_LSTM_CLUSTER=1 GLASS_DEC_PRE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>$O/bench_both.err | tail -1 > $O/bench_both.json
timeout 200 python tools/layer_profile.py full_bs4 > $O/layers_full.txt 2>&1
timeout 200 python tools/layer_profile.py backbone_bs8 > $O/layers_bb.txt 2>&1
nproc > $O/host.txt; grep -m1 "model name" /proc/cpuinfo >> $O/host.txt; free -g >> $O/host.txt
tail -3 $O/lstm_ab.txt $O/optin_tests.log; for f in $O/bench_*.json; do python - <<PY
import json
try:
    d=json.load(open("$f")); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"])
except Exception as e: print("$f", "ERR", e)
PY
done

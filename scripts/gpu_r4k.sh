#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/coarse_r4k.jsonl
DBG_OBJECTS=125 DBG_CONFIGS=1:0:1 DBG_SPLITS=0,2,3,4 DBG_ITERS=12 timeout 200 python scripts/gpu_coarse_dbg.py splits_125k >> gpurun_out/coarse_r4k.jsonl 2>> gpurun_out/coarse_r4k.err
DBG_OBJECTS=500 DBG_CONFIGS=1:0:1 DBG_SPLITS=0,3,4,6 DBG_ITERS=10 timeout 200 python scripts/gpu_coarse_dbg.py splits_500k >> gpurun_out/coarse_r4k.jsonl 2>> gpurun_out/coarse_r4k.err
DBG_CONFIGS=1:0:1 DBG_SPLITS=0,4,6 DBG_ITERS=8 timeout 300 python scripts/gpu_coarse_dbg.py splits_1m >> gpurun_out/coarse_r4k.jsonl 2>> gpurun_out/coarse_r4k.err
python - <<'PY'
import json
for l in open("gpurun_out/coarse_r4k.jsonl"):
    d = json.loads(l)
    print(d["label"], "splits", d["splits"], "coarse_ms %.3f match_ms %.3f" % (d["coarse_ms"], d["match_ms"]), d["tiers"], d["same_bits_as_first"])
PY

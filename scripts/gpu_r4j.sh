#!/bin/bash
# database splits per query tile of the coarse pass: 125 k-row shard (the 8-GPU case), 250 k, 1 M
mkdir -p gpurun_out; : > gpurun_out/coarse_r4j.jsonl
DBG_OBJECTS=125 DBG_CONFIGS=1:0:1 DBG_SPLITS=0,1,2,3,4,6 DBG_ITERS=10 timeout 200 python scripts/gpu_coarse_dbg.py splits_125k >> gpurun_out/coarse_r4j.jsonl 2>> gpurun_out/coarse_r4j.err
DBG_OBJECTS=250 DBG_CONFIGS=1:0:1 DBG_SPLITS=0,2,4,6 DBG_ITERS=10 timeout 200 python scripts/gpu_coarse_dbg.py splits_250k >> gpurun_out/coarse_r4j.jsonl 2>> gpurun_out/coarse_r4j.err
DBG_CONFIGS=1:0:1 DBG_SPLITS=0,4,6,12 DBG_ITERS=8 timeout 300 python scripts/gpu_coarse_dbg.py splits_1m >> gpurun_out/coarse_r4j.jsonl 2>> gpurun_out/coarse_r4j.err
python - <<'PY'
import json
for l in open("gpurun_out/coarse_r4j.jsonl"):
    d = json.loads(l)
    print(d["label"], "splits", d["splits"], "coarse_ms %.3f match_ms %.3f" % (d["coarse_ms"], d["match_ms"]), d["tiers"], d["stats"], d["same_bits_as_first"])
PY

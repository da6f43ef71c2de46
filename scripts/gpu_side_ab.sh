#!/bin/bash
# A/B of the side-stream weight-gradient GEMMs (AC_TRAIN_SIDE=0 keeps everything on the main stream)
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu 2>&1 | tail -5
for side in 0 1; do
  AC_TRAIN_SIDE=$side timeout 600 python bench.py --workload train --steps 30 --warmup 5 > $OUT/train_side$side.json 2> $OUT/train_side$side.err
  python - <<P
import json
for l in open("$OUT/train_side$side.json"):
    if l.startswith("{"):
        d = json.loads(l); print("side=$side", d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["ms_per_step_ss_ratio_0.85"], d["roofline"]["spans_ms_per_step"])
P
done

#!/bin/bash
# training-path regression + bench line of the training workload
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-t}
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_entry_points.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --workload train --steps 30 --warmup 5 > $OUT/train_$TAG.json 2> $OUT/train_$TAG.err
python - <<P
import json
for l in open("$OUT/train_$TAG.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("unpipelined_ms_per_step"), "ss085", d["ms_per_step_ss_ratio_0.85"], d["roofline"]["spans_ms_per_step"])
P
tail -3 $OUT/train_$TAG.err

#!/bin/bash
# look-ahead encoder (TrainStep.prefetch): equivalence test, then the training bench for several convolution SM caps
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu 2>&1 | tail -4
for sms in ${@:-84 108 132}; do
  AC_TRAIN_CNN_SMS=$sms timeout 600 python bench.py --workload train --steps 30 --warmup 5 > $OUT/train_la$sms.json 2> $OUT/train_la$sms.err
  python - <<P
import json
for l in open("$OUT/train_la$sms.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("cnn_sms=$sms pipelined", round(d["ms_per_step"], 4), round(d["value"]), "inline", round(d["pipeline"]["unpipelined_ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "e2e-sync", round(d["e2e"]["unpipelined_ms_per_step"], 4))
P
  tail -2 $OUT/train_la$sms.err
done

#!/bin/bash
# n-tile width of the bf16 convolution (AC_CONV_BF16_BN=128|256): parity tests, then training step and configs[4] timings
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu -k "bf16 or conv3x3 or precision" 2>&1 | tail -3
for bn in 128 256; do
  export AC_CONV_BF16_BN=$bn
  timeout 600 python bench.py --workload train --steps 30 --warmup 5 > $OUT/train_bn$bn.json 2> $OUT/train_bn$bn.err
  python - <<P
import json
for l in open("$OUT/train_bn$bn.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("BN=$bn train", round(d["ms_per_step"], 4), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"], 4), "conv ms", r["ms_per_step"], "TF/s", round(r["achieved"], 1), "frac", round(r["frac"], 3))
P
  python scripts/c5_time.py 20 2>&1 | tail -1
done

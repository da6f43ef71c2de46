#!/bin/bash
# config-5 (TempGRU beam-4 + SED) with and without the tagger / encoder overlap, plus the GPU parity tests of that path
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "temp or sed or hf or config5 or tag" 2>&1 | tail -3
for ov in 0 1; do
  AC_SED_OVERLAP=$ov timeout 600 python bench.py --no-train --steps 20 --warmup 3 > $OUT/c5_ov$ov.json 2> $OUT/c5_ov$ov.err
  python - <<P
import json
for l in open("$OUT/c5_ov$ov.json"):
    if l.startswith("{"):
        d = json.loads(l); c = d["config5_tempgru_beam4"]
        print("overlap=$ov", {k: (round(c[k]["ms_per_step"], 3), round(c[k]["e2e"]["ms_per_step"], 3)) for k in ("bf16", "tf32", "fp32")}, "headline", d["ms_per_step"], d["e2e"]["ms_per_step"])
P
done

#!/bin/bash
# config-5 (TempGRU beam-4 + SED, 16 clips) with and without the tagger / encoder overlap, plus the GPU parity tests of that path
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "temp or sed or hf or config5 or tag" 2>&1 | tail -3
for ov in 0 1; do
  echo "AC_SED_OVERLAP=$ov"
  AC_SED_OVERLAP=$ov timeout 600 python scripts/c5_time.py 20 2>&1 | tail -1
done

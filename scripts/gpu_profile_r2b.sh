#!/bin/bash
# Round-2 (final) ncu evidence: launch lists + DRAM traffic of one inference step and one training step, full captures of the
# kernels written or rewritten late in the round (head-split greedy / beam decode, stem, log-mel, SE).
OUT=gpurun_out; TAG=${1:-r2}
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_traffic.csv python scripts/one_step.py 1 > $OUT/${TAG}_traffic.log 2>&1; echo "infer traffic rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_train_traffic.csv python scripts/train_one_step.py 1 tf32 > $OUT/${TAG}_train_traffic.log 2>&1; echo "train traffic rc=$?"
cap() {  # name, kernel regex, skip, count, script...
  name=$1; re=$2; skip=$3; cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c $cnt -f -o $OUT/${TAG}_$name \
      "$@" > $OUT/${TAG}_ncu_$name.log 2>&1; echo "$name rc=$?"
}
cap greedy_heads ^greedy_heads_kernel 1 1 python scripts/one_step.py 1
cap stem ^stem_kernel 1 1 python scripts/one_step.py 1
cap logmel ^logmel_kernel 1 1 python scripts/one_step.py 1
cap se ^se_kernel 40 1 python scripts/one_step.py 1
cap beam_heads ^beam_heads_kernel 1 1 python scripts/beam_time.py
ls -la $OUT/${TAG}_*.ncu-rep; du -sh $OUT

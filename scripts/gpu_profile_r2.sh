#!/bin/bash
# Round-2 ncu evidence for profiles/: launch list + DRAM traffic of ONE training step and of one inference step, and full
# captures of the new kernels (GRU backward, attention backward, TF32 convolution, loss, optimizer).
OUT=gpurun_out; TAG=${1:-r2}
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_train_traffic.csv python scripts/train_one_step.py 1 tf32 > $OUT/${TAG}_train_traffic.log 2>&1; echo "train traffic rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_traffic.csv python scripts/one_step.py 1 > $OUT/${TAG}_traffic.log 2>&1; echo "infer traffic rc=$?"
cap() {  # name, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o $OUT/${TAG}_$1 \
      python scripts/train_one_step.py 1 tf32 > $OUT/${TAG}_ncu_$1.log 2>&1; echo "$1 rc=$?"
}
cap bigru_bwd bigru_bwd_kernel 3 1
cap bigru_fwd bigru_recurrence_kernel 3 1
cap attn_bwd attn_bwd_kernel 4 1
cap conv_tf32 "gemm_tc_kernel.*true" 18 1
cap ls_ce ls_ce_kernel 1 1
cap adam adam_kernel 1 1
ls -la $OUT/*.ncu-rep; du -sh $OUT

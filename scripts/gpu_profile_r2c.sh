#!/bin/bash
# ncu evidence for the bf16 precision mode: launch list + DRAM traffic of one training step in bf16, full capture of the
# bf16 convolution kernel (block-4 conv2 shape of the training step).
OUT=gpurun_out; TAG=${1:-r2}
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_train_traffic.csv python scripts/train_one_step.py 1 bf16 > $OUT/${TAG}_train_traffic.log 2>&1; echo "train traffic rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^conv3x3_bf16_kernel -s 17 -c 1 -f -o $OUT/${TAG}_conv_bf16 \
    python scripts/train_one_step.py 1 bf16 > $OUT/${TAG}_ncu_conv_bf16.log 2>&1; echo "conv_bf16 rc=$?"
ls -la $OUT/${TAG}_conv_bf16.ncu-rep

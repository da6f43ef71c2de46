#!/bin/bash
# ncu evidence for profiles/: per-launch DRAM traffic + time of one step, and full captures of the top kernels.
OUT=gpurun_out; TAG=${1:-r1}
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_traffic.csv python scripts/one_step.py 1 > $OUT/${TAG}_traffic.log 2>&1; echo "traffic rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 50 -c 2 -f -o $OUT/${TAG}_gemm_b2 \
    python scripts/one_step.py 1 > $OUT/${TAG}_ncu_gemm.log 2>&1; echo "gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_tma_kernel -s 25 -c 1 -f -o $OUT/${TAG}_dw_b2 \
    python scripts/one_step.py 1 > $OUT/${TAG}_ncu_dw.log 2>&1; echo "dw rc=$?"
ls -la $OUT/*.ncu-rep; du -sh $OUT

#!/bin/bash
# One GPU-box visit: parity tests, headline bench, ncu launch list, ncu full capture of selected kernels.
# usage (under gpurun): bash scripts/gpu_round.sh <tag> [ncu-kernel-regex] [skip] [count]
TAG=${1:-run}
KRE=${2:-gemm_tc_kernel}
SKIP=${3:-48}
CNT=${4:-6}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
tail -3 $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python scripts/one_step.py 1 > $OUT/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
if [ "$CNT" != "0" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s $SKIP -c $CNT -f -o $OUT/${TAG}_prof \
    python scripts/one_step.py 1 > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
ls -la $OUT | head -30
du -sh $OUT

#!/bin/bash
# ncu --set full of the training-step kernels (configs[2]): scripts/gpu_ncu_train.sh <tag> <kernel-regex> [variant]
TAG=$1; KRE=$2; VAR=${3:-vm}
OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python bench.py --workload train --variant $VAR --steps 2 --warmup 3"
timeout 600 $CMD > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s 6 -c 3 -f -o $OUT/train $CMD > $OUT/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu.log

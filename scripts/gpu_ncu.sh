#!/bin/bash
# one `ncu --set full` capture of a kernel of a bench workload, after the plain run has exited 0
# usage: scripts/gpu_ncu.sh <tag> <name> <kernel-regex> <skip> <count> -- <bench args...>
TAG=$1; NAME=$2; KRE=$3; SKIP=$4; COUNT=$5; shift 6
OUT=gpurun_out/$TAG
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras $@"
timeout 600 $CMD > $OUT/${NAME}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${NAME}_plain.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s $SKIP -c $COUNT -f -o $OUT/$NAME $CMD > $OUT/${NAME}_ncu.log 2>&1
echo "ncu $NAME rc=$?"; tail -2 $OUT/${NAME}_ncu.log

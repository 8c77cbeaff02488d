#!/bin/bash
# source-level counters (instructions executed / stall samples per SASS line) of every launch of one kernel in a bench run
# usage: scripts/gpu_ncu_src.sh <tag> <name> <kernel-regex> <count> -- <bench args...>
TAG=$1; NAME=$2; KRE=$3; COUNT=$4; shift 5
OUT=gpurun_out/$TAG
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline $@"
timeout 600 $CMD > $OUT/${NAME}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${NAME}_plain.log; exit 1; }
timeout 1200 ncu --section SourceCounters --section LaunchStats --section SpeedOfLight --section WarpStateStats --clock-control none --import-source on -k "regex:$KRE" -c $COUNT -f -o $OUT/$NAME $CMD > $OUT/${NAME}_ncu.log 2>&1
echo "ncu $NAME rc=$?"; tail -2 $OUT/${NAME}_ncu.log

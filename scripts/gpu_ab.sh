#!/bin/bash
# frame bench of the default library and of variant libraries, one gpurun call: scripts/gpu_ab.sh <tag> [variant ...]
# env: BARGS = extra bench.py arguments; ENVS = "NAME=VALUE ..." extra runs of the default library under that environment
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
V=$PWD/jittor-myc-nerfs_b200/variants
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras $BARGS"
timeout 300 $B > $OUT/bench_default.json 2> $OUT/bench_default.err || tail -3 $OUT/bench_default.err
for v in "$@"; do
  TVM_LIB=$V/libtvmrender_$v.so timeout 300 $B > $OUT/bench_$v.json 2> $OUT/bench_$v.err || tail -3 $OUT/bench_$v.err
done
for e in $ENVS; do
  env $e timeout 300 $B > $OUT/bench_env_$e.json 2> $OUT/bench_env_$e.err || tail -3 $OUT/bench_env_$e.err
done
python scripts/bshow.py $OUT/bench_*.json
grep -o "stage ms/step {[^}]*}" $OUT/*.err /dev/null

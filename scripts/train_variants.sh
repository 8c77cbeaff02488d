#!/bin/bash
# stage times of the captured training step (configs[2]) for the default library and variant libraries: scripts/train_variants.sh v1 v2 ...
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], {k: round(v, 4) for k, v in d["stage_ms_per_step"].items()}, "graph step %.4f ms" % d["full_step_cuda_graph"]["ms_per_step"])'
for v in default "$@" default; do
  if [ $v = default ]; then L=""; else L=$PWD/jittor-myc-nerfs_b200/variants/libtvmrender_$v.so; fi
  TVM_LIB=$L python bench.py --workload train --steps ${STEPS:-200} --warmup 3 2>/dev/null | python -c "$show" $v
done

#!/bin/bash
# side workloads (configs[2], configs[3]) -> gpurun_out/<tag>/side_*.json
TAG=${1:-side}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for w in ${WORKLOADS:-train ref npp}; do
  for r in ${REGIMES:-R1}; do
    timeout ${SIDE_TIMEOUT:-600} python bench.py --workload $w --regime $r --steps ${STEPS:-5} --warmup 3 ${SIDE_ARGS} > $OUT/side_${w}_${r}.json 2> $OUT/side_${w}_${r}.err
    echo "$w $r rc=$?"; tail -2 $OUT/side_${w}_${r}.err; cat $OUT/side_${w}_${r}.json
  done
done

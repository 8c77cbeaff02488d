#!/bin/bash
# captured training step (configs[2]) of the default library and of variant libraries, alternating: scripts/train_ab.sh <variant> [steps]
V=$PWD/jittor-myc-nerfs_b200/variants/libtvmrender_$1.so
S=${2:-200}
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], {k: round(v, 4) for k, v in d["stage_ms_per_step"].items()}, "graph step %.4f ms" % d["full_step_cuda_graph"]["ms_per_step"])'
for i in 1 2; do
  python bench.py --workload train --steps $S --warmup 3 2>/dev/null | python -c "$show" default
  TVM_LIB=$V python bench.py --workload train --steps $S --warmup 3 2>/dev/null | python -c "$show" $1
done
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c '
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("frame", d["ms_per_step"], d["roofline"]["stage_ms_per_step"], "train block", d["train"]["ms_per_step"])'

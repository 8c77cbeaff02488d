#!/bin/bash
# training block of the bench line at N ranks, peer all-reduce vs NCCL: scripts/gpu_train_scale.sh <tag> <N> [ctas...]
TAG=$1; N=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_$name.json 2> $OUT/bench_$name.err || tail -5 $OUT/bench_$name.err
  python - <<PY
import json
try:
    j = json.loads(open("$OUT/bench_$name.json").read().strip().splitlines()[-1])
    t = j["train"]
    print("$name", "frame %.1f M rays/s" % (j["value"] / 1e6), "| train ms %.4f" % t["ms_per_step"], "rays/s %.2f M" % (t["value"] / 1e6), t["gradient_exchange"][:90])
except Exception as e:
    print("$name ERR", e)
PY
}
run peer TVM_AR=peer
run nccl TVM_AR=nccl
for c in "$@"; do run peer_c$c TVM_AR=peer TVM_AR_CTAS=$c; done

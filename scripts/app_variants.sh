#!/bin/bash
# Kernel-tuning harness for tvm_mlp_tc.cu: builds variant libraries from -D switches HERE (no GPU needed), then one
# gpurun call benches the default library and every variant on the frame workload and prints their stage times.
#   scripts/app_variants.sh build  NOGATHER NOMLP ...   -> jittor-myc-nerfs_b200/variants/libtvmrender_<V>.so  (-DTVM_EXP_<V>)
#   scripts/app_variants.sh bench  NOGATHER NOMLP ...   -> run on the GPU box (inside gpurun): stage ms per variant
# TVM_EXP_NOGATHER: the gather group writes zeros (MLP group alone); TVM_EXP_NOMLP: the MLP group only consumes the GEMM0
# operand (gather group alone).  Their pixels are wrong on purpose: bench.py refuses the line and prints the stage times
# in its error message, which is what this script greps (profiles/r01q_notes.txt has the numbers of round 1).
set -e
cd "$(dirname "$0")/.."
MODE=$1; shift
CSRC=jittor-myc-nerfs_b200/csrc
VDIR=jittor-myc-nerfs_b200/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
if [ "$MODE" = build ]; then
  make -C $CSRC > /dev/null
  mkdir -p $VDIR
  for v in "$@"; do
    nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DTVM_EXP_$v -c $CSRC/tvm_mlp_tc.cu -o /tmp/tvm_mlp_tc_$v.o
    OBJS=$(ls $CSRC/*.o | grep -v tvm_mlp_tc.o)
    nvcc $ARCH -shared -o $VDIR/libtvmrender_$v.so $OBJS /tmp/tvm_mlp_tc_$v.o -lcudart
    echo "built $VDIR/libtvmrender_$v.so"
  done
else
  show() { python -c "import json,sys; j=json.load(sys.stdin); print('$1', round(j['value']/1e6,1), 'M rays/s', j['roofline']['stage_ms_per_step'], 'err', j['max_abs_err_vs_oracle_512rays'])"; }
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | show default
  for v in "$@"; do
    TVM_LIB=$PWD/$VDIR/libtvmrender_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/tmp/err_$v.log > /tmp/out_$v.json \
      && show $v < /tmp/out_$v.json || { echo -n "$v: "; grep -o "stage ms/step {[^}]*}" /tmp/err_$v.log; }
  done
fi

#!/bin/bash
# Build a variant libtvmrender from -D switches on ONE translation unit (no GPU needed):
#   scripts/variant.sh <name> <file.cu> <nvcc flags...>   -> jittor-myc-nerfs_b200/variants/libtvmrender_<name>.so
# select it at run time with TVM_LIB=$PWD/jittor-myc-nerfs_b200/variants/libtvmrender_<name>.so
set -e
cd "$(dirname "$0")/.."
NAME=$1; SRC=$2; shift 2
CSRC=jittor-myc-nerfs_b200/csrc
VDIR=jittor-myc-nerfs_b200/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
make -s -C $CSRC > /dev/null
mkdir -p $VDIR
nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $CSRC/$SRC -o /tmp/variant_$NAME.o
OBJS=$(ls $CSRC/*.o | grep -v "/${SRC%.cu}.o")
nvcc $ARCH -shared -o $VDIR/libtvmrender_$NAME.so $OBJS /tmp/variant_$NAME.o -lcudart
echo "built $VDIR/libtvmrender_$NAME.so"

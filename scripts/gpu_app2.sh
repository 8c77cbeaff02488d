#!/bin/bash
# A/B of the appearance-head kernels in one gpurun call: tests of the tensor-core head, then the frame bench with the default
# library (k_app_tc2), the single-group kernel (TVM_APP_TC=1) and every variant library given on the command line.
TAG=${1:-app2}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
V=$PWD/jittor-myc-nerfs_b200/variants
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q -k "tensor_core or bf16_appearance or reftensorf or full_frame or no_write or streamed" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B > $OUT/bench_default.json 2> $OUT/bench_default.err || tail -3 $OUT/bench_default.err
TVM_APP_TC=1 timeout 300 $B > $OUT/bench_v1.json 2> $OUT/bench_v1.err || tail -3 $OUT/bench_v1.err
for v in "$@"; do
  TVM_LIB=$V/libtvmrender_$v.so timeout 300 $B > $OUT/bench_$v.json 2> $OUT/bench_$v.err || tail -3 $OUT/bench_$v.err
done
if [ -n "$R2" ]; then
  timeout 300 $B --regime R2 > $OUT/bench_default_R2.json 2> $OUT/bench_default_R2.err || tail -3 $OUT/bench_default_R2.err
fi
python scripts/bshow.py $OUT/bench_*.json

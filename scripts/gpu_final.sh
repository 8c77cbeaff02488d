#!/bin/bash
# One gpurun call for the round's records: GPU tests, smoke, the contract bench line, launch list, ncu --set full of the three
# frame kernels (each the whole frame in one launch), side workloads.   usage: scripts/gpu_final.sh <tag>
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -3 $OUT/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -3 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -2 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 $CMD > $OUT/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
# TVM_WS_GIB=80: the worst-case workspace fits, every resident frame is one launch per kernel (what the roofline legs time)
TVM_WS_GIB=80 timeout 600 $CMD > $OUT/plain80.log 2>&1 &&
TVM_WS_GIB=80 timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_march|k_app_tc2|k_composite" -s 0 -c 3 -f -o $OUT/prof $CMD > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 $OUT/ncu_full.log
for w in ${WORKLOADS:-train ref npp}; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > $OUT/side_$w.json 2> $OUT/side_$w.err; echo "$w rc=$?"
done
for v in ref npp; do
  timeout 600 python bench.py --workload train --variant $v --steps 5 --warmup 3 > $OUT/train_$v.json 2> $OUT/train_$v.err; echo "train $v rc=$?"
done
timeout 600 python bench.py --workload maintain --steps 5 --warmup 3 > $OUT/side_maintain.json 2> $OUT/side_maintain.err; echo "maintain rc=$?"
ls -la $OUT

#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, launch list, one full ncu capture of the top kernel.
# usage: scripts/gpu_round.sh <tag> [pytest-args...]
TAG=${1:-r}
shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest" 
timeout 900 python -m pytest tests -q -m gpu "$@" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -25 $OUT/pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
tail -8 $OUT/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
tail -3 $OUT/bench.err; cat $OUT/bench.json
if [ -n "$BENCH_EXTRA" ]; then
  for extra in $BENCH_EXTRA; do
    timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $(echo $extra | tr ',' ' ') > $OUT/bench_$extra.json 2> $OUT/bench_$extra.err; echo "bench $extra rc=$?"; cat $OUT/bench_$extra.json
  done
fi
if [ -z "$NO_NCU" ]; then
echo "== ncu launches"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 $CMD > $OUT/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:${NCU_KERNEL:-k_march|k_app_tc2}" -s ${NCU_SKIP:-36} -c ${NCU_COUNT:-2} -f -o $OUT/prof $CMD > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 $OUT/ncu_full.log
fi
ls -la $OUT

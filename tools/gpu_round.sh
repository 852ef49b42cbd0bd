#!/bin/bash
# One GPU-box visit (run under gpurun): parity tests, bench lines, ncu launch list, full captures.
# Usage: bash tools/gpu_round.sh <tag> [steps...]   steps: tests bench driver ref launches icp tps trace
TAG=${1:-rX}; shift
STEPS=${@:-tests driver bench ref launches icp}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_$TAG.csv 2>&1
for S in $STEPS; do
case $S in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu_$TAG.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
  tail -15 $OUT/pytest_gpu_$TAG.log ;;
driver)   # the driver's invocation
  timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_driver_$TAG.json 2> $OUT/bench_driver_$TAG.err
  echo "bench(driver) exit $?"; cat $OUT/bench_driver_$TAG.json; tail -3 $OUT/bench_driver_$TAG.err ;;
bench)
  timeout 600 python bench.py --skip-extras > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
  echo "bench exit $?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err ;;
ref)
  timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
  cat $OUT/bench_ref_$TAG.json ;;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 130 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 6 --warmup 3 --reps 1 --skip-extras --no-pipeline > $OUT/ncu_launches_$TAG.log 2>&1 ;;
icp)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_system -s 4 -c 1 -f -o $OUT/icp_$TAG \
    python bench.py --roofline-only > $OUT/ncu_icp_$TAG.log 2>&1 ;;
tps)      # full captures of the segmentation / registration kernels inside the synchronous frame graph (-s skips frame 1)
  for K in tps_pass_tile_kernel:110 tps_filter_kernel:2 tps_init_samples_kernel:2 icp_loop_kernel:2; do
    N=${K%%:*}; SK=${K##*:}
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$N -s $SK -c 1 -f -o $OUT/${N}_$TAG \
      python bench.py --steps 6 --warmup 3 --reps 1 --skip-extras --no-pipeline > $OUT/ncu_${N}_$TAG.log 2>&1
  done ;;
trace)
  bash tools/tps_ab.sh > $OUT/tps_ab_$TAG.log 2>&1; tail -40 $OUT/tps_ab_$TAG.log ;;
esac
done
ls -la $OUT | tail -25

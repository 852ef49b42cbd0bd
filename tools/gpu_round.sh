#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, one full capture of the roofline kernel.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_$TAG.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?"; cat $OUT/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
cat $OUT/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 6 --warmup 3 --skip-extras > $OUT/ncu_launches_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_system -s 4 -c 1 -f -o $OUT/icp_$TAG \
  python bench.py --roofline-only > $OUT/ncu_icp_$TAG.log 2>&1
ls -la $OUT | tail -20

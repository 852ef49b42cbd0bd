#!/bin/bash
# sweep of the ICP system kernel's tuning knobs at the roofline sizing (run under gpurun)
mkdir -p gpurun_out
for cfg in "3 1" "4 1" "3 2"; do
  set -- $cfg
  echo "occ=$1 stages=$2" ; SSF_ICP_OCC=$1 SSF_ICP_STAGES=$2 timeout 300 python bench.py --roofline-only 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['us_per_launch'], d['frac'], d['latency_us'])"
done

#!/bin/bash
# A/B of the segmentation schedules on one GPU (run under gpurun): graph of small kernels vs the
# persistent cooperative kernel; frames/s of the whole frame + the persistent kernel's phase trace
for p in 0 1; do
  echo "SSF_TPS_PERSISTENT=$p"
  SSF_TPS_PERSISTENT=$p python bench.py --steps 300 --warmup 10 --skip-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['last_frame_stats']['gpu_ms'], d['gpu_launches'])"
done
SSF_TPS_PERSISTENT=1 python tools/tps_trace.py 2>&1 | tail -30

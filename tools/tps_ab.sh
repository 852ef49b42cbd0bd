#!/bin/bash
# A/B of the segmentation schedules on one GPU (run under gpurun): fused passes (default) vs the
# round-1 pass + merge launches vs the persistent cooperative kernel; frames/s of the whole frame
for v in "SSF_TPS_FUSED=1" "SSF_TPS_FUSED=0" "SSF_TPS_FUSED=1 SSF_PDL=1" "SSF_TPS_PERSISTENT=1"; do
  echo "$v"
  env $v python bench.py --steps 200 --warmup 10 --reps 3 --skip-extras 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('  pipelined value %.0f e2e %.0f fps | last frame gpu_ms %.3f | launches/200 frames %d' % (d['value'], d['e2e']['value'], d['last_frame_stats']['gpu_ms'], d['gpu_launches']))"
  env $v python bench.py --steps 200 --warmup 10 --reps 3 --skip-extras --no-pipeline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('  synchronous value %.0f e2e %.0f fps | ms/frame %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done

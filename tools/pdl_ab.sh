#!/bin/bash
# A/B of programmatic dependent launch on the synchronous frame (run under gpurun): SSF_PDL 0 (plain graph edges),
# 1 (every kernel lets its successor launch at its top), 2 (the fused segmentation pass triggers after its decisions)
for v in ${MODES:-0 1 2 3}; do
  echo "SSF_PDL=$v"
  SSF_PDL=$v python bench.py --steps 200 --warmup 10 --reps 3 --skip-extras --no-pipeline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('  synchronous value %.0f e2e %.0f fps | ms/frame %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
  SSF_PDL=$v python bench.py --steps 200 --warmup 10 --reps 3 --skip-extras 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('  pipelined value %.0f e2e %.0f fps' % (d['value'], d['e2e']['value']))"
done

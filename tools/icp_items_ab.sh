#!/bin/bash
# A/B of the ICP system kernel at the roofline sizing (run under gpurun): library build (default, or
# variants/libssf_*.so from tools/build_variant.sh) x resident CTAs per SM the kernel is compiled for
# (OCCS) x staging of the streamed planes (STAGES: 1 direct loads, 2..4 TMA ring, -2..-4 cp.async ring)
mkdir -p gpurun_out
for lib in ${LIBS:-supersurfel_fusion_b200/libssf.so variants/libssf_*.so}; do
  [ -f "$lib" ] || continue
  for occ in ${OCCS:-3 4 5}; do
    for st in ${STAGES:-1}; do
      echo "lib=$lib occ=$occ stages=$st"
      SSF_LIB=$PWD/$lib SSF_ICP_OCC=$occ SSF_ICP_STAGES=$st timeout 300 python bench.py --roofline-only 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('  us', round(d['us_per_launch'],1), 'binned', round(d['binned_by_tile']['us_per_launch'],1), 'frac', round(d['frac'],3), 'lat', d['latency_us'])"
    done
  done
done

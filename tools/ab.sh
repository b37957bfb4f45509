#!/bin/bash
# A/B runner on the GPU box: tools/ab.sh "<workload> <workload> ..." [variant .so ...]; prints ms/step, K2 ms and roofline fraction per (lib, workload)
WLS="$1"; shift
for lib in default "$@"; do
  for w in $WLS; do
    if [ "$lib" = default ]; then unset SBM_LIB_PATH; else export SBM_LIB_PATH=$lib; fi
    python bench.py --workload $w --steps 30 --warmup 5 --blocks 3 --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read()
try:
    d=json.loads(l); r=d['roofline']
    print('$lib $w ms %.3f e2e %.3f k2 %.3f frac %.3f k1 %.3f tail %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step'], r['kernel_ms'], (r['frac'] or 0), r['other_kernels_ms']['k1_params'], r['other_kernels_ms']['k3_reduce+collective+k4_finish']))
except Exception as e: print('$lib $w ERR', l[-300:])
"
  done
done

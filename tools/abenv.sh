#!/bin/bash
# tools/abenv.sh "<ENV=VAL ...>;<ENV=VAL ...>;..." [bench args]: one short bench run per environment setting (development A/B on the GPU box)
IFS=';' read -ra SETS <<< "$1"; shift
i=0
for s in "${SETS[@]}"; do
  env $s timeout 300 python bench.py --no-cpu --no-also --blocks 3 "$@" > gpurun_out/abenv_$i.json 2> gpurun_out/abenv_$i.err
  python - "$s" gpurun_out/abenv_$i.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    r = d["roofline"]
    print("%-40s ms/step %.4f  e2e %.4f  k1 %.4f k2 %.4f tail %.4f frac %s" % (sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], r["other_kernels_ms"]["k1_params"], r["kernel_ms"], r["other_kernels_ms"]["k3_reduce+collective+k4_finish"], r["frac"]))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
  i=$((i+1))
done

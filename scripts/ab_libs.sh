#!/bin/bash
# A/B on the GPU box between library variants (scripts/build_variant.sh NAME ...): bench.py for each NAME given ("-" = the regular build)
for rep in 1 2; do
for name in "$@"; do
  if [ "$name" = "-" ]; then lib=$PWD/pitchvis_b200/lib/libpvqt.so; else lib=$PWD/pitchvis_b200/lib/libpvqt_$name.so; fi
  PVQT_LIB=$lib timeout 300 python bench.py --steps 30 --configs none --no-cpu-baseline > gpurun_out/abl.json 2> gpurun_out/abl.err
  python - <<P
import json
d=json.load(open("gpurun_out/abl.json"))
print("[$name]", round(d["value"]/1e6,2), d["step_ms"]["median"], {k:round(v["avg_ms"]*1e3,1) for k,v in d["roofline"]["kernels"].items()}, round(d["e2e"]["value"]/1e6,2))
P
done; done

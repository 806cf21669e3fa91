#!/bin/bash
# streams4096 (16 launches of 131,072 frames per step) against the number of launch lanes (scratch sets + streams) and the chunk size
for setting in "PVQT_LANES=1" "PVQT_LANES=2" "PVQT_LANES=3" "PVQT_LANES=2 PVQT_CHUNK_FRAMES=65536" "PVQT_LANES=1 PVQT_CHUNK_FRAMES=262144"; do
  env $setting timeout 300 python bench.py --workload streams4096 --steps 5 --warmup 3 --configs none --no-cpu-baseline --sustain 0 > gpurun_out/ls.json 2> gpurun_out/ls.err
  python - <<P
import json
d=json.load(open("gpurun_out/ls.json"))
print("[$setting] %.2f M frames/s  %.2f ms/step  e2e %.2f M" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6))
P
done

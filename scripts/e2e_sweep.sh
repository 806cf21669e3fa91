#!/bin/bash
# end-to-end (host buffers) throughput of chords60 against the host-pipeline knobs
for seg in 2 3 4 6; do for hl in 1 2 3; do
  PVQT_SEGMENTS=$seg PVQT_HOST_LANES=$hl timeout 300 python bench.py --steps 20 --configs none --no-cpu-baseline > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - <<P
import json
d=json.load(open("gpurun_out/sw.json"))
print("segments $seg host_lanes $hl  e2e %.2f M frames/s  device %.2f M" % (d["e2e"]["value"]/1e6, d["value"]/1e6))
P
done; done

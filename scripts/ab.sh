#!/bin/bash
# A/B on the GPU box: bench.py (and optionally the large-batch script) for each library variant named.
#   scripts/ab.sh mb1 mb5 mb6   -> gpurun_out/ab_<name>.json ; prints value / step / kernels
mkdir -p gpurun_out
for n in "$@"; do
  PVQT_LIB=$PWD/pitchvis_b200/lib/libpvqt_$n.so timeout 300 python bench.py --no-cpu-baseline --steps 50 > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err || echo "$n FAILED"
  PVQT_LIB=$PWD/pitchvis_b200/lib/libpvqt_$n.so timeout 300 python scripts/large_batch.py > gpurun_out/ab_${n}_large.txt 2>&1 || echo "$n large FAILED"
  python - "$n" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/ab_{n}.json'))
    print(n, 'value %.2f M/s step %.2f us'%(d['value']/1e6, d['step_ms']['median']*1e3), {k:round(v*1e3,1) for k,v in d['roofline']['kernels_avg_ms'].items()}, 'e2e %.2f M/s'%(d['e2e']['value']/1e6))
except Exception as e: print(n,'no result',e)
P
  tail -3 gpurun_out/ab_${n}_large.txt
done

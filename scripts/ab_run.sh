#!/bin/bash
# A/B on the GPU box: parity tests, then bench.py for each environment setting given as arguments ("VAR=value", or "-" for none)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_properties.py tests/test_gpu_analysis.py -m gpu -q -x > gpurun_out/ab_pytest.log 2>&1; tail -3 gpurun_out/ab_pytest.log
i=0
for setting in "$@"; do
  i=$((i+1))
  if [ "$setting" = "-" ]; then envs=""; else envs="$setting"; fi
  env $envs timeout 300 python bench.py --steps 30 --configs none --no-cpu-baseline > gpurun_out/ab_bench_$i.json 2> gpurun_out/ab_bench_$i.err
  python - <<P
import json
d=json.load(open("gpurun_out/ab_bench_$i.json"))
print("[$setting]", round(d["value"]/1e6,2), d["step_ms"]["median"], {k:round(v["avg_ms"]*1e3,1) for k,v in d["roofline"]["kernels"].items()}, round(d["e2e"]["value"]/1e6,2))
P
done

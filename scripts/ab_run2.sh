#!/bin/bash
# parity tests, then chords60 and a 1024-stream job with the regular build
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_properties.py tests/test_gpu_analysis.py -m gpu -q -x > gpurun_out/ab_pytest.log 2>&1; tail -2 gpurun_out/ab_pytest.log
for rep in 1 2; do
timeout 300 python bench.py --steps 30 --configs none --no-cpu-baseline > gpurun_out/ab2.json 2> gpurun_out/ab2.err
python - <<P
import json
d=json.load(open("gpurun_out/ab2.json"))
print("chords60", round(d["value"]/1e6,2), d["step_ms"]["median"], {k:round(v["avg_ms"]*1e3,1) for k,v in d["roofline"]["kernels"].items()})
P
done
timeout 300 python bench.py --workload streams4096 --streams 1024 --steps 5 --warmup 3 --configs none --no-cpu-baseline --sustain 0 > gpurun_out/ab2s.json 2> gpurun_out/ab2s.err
python - <<P
import json
d=json.load(open("gpurun_out/ab2s.json"))
print("streams1024 %.2f M frames/s" % (d["value"]/1e6), {k:round(v["avg_ms"]*1e3,1) for k,v in d["roofline"]["kernels"].items()})
P

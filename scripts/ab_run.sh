timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_properties.py -m gpu -q -x > gpurun_out/ab_pytest.log 2>&1; tail -3 gpurun_out/ab_pytest.log
for n in _base ""; do PVQT_LIB=$PWD/pitchvis_b200/lib/libpvqt$n.so timeout 300 python bench.py --steps 30 --configs none --no-cpu-baseline > gpurun_out/ab_bench$n.json 2> gpurun_out/ab_bench$n.err; python - <<P
import json
d=json.load(open("gpurun_out/ab_bench$n.json"))
print("lib$n", round(d["value"]/1e6,2), d["step_ms"]["median"], {k:round(v["avg_ms"]*1e3,1) for k,v in d["roofline"]["kernels"].items()}, round(d["e2e"]["value"]/1e6,2))
P
done

"""Per-frame latency of pvqt_calc_instant_db (the reference's real-time call, vqt.rs:866) against the CPU port."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import orc
import pitchvis_b200 as pv

v = pv.Vqt()
o = orc.OracleVqt()
x = orc.test_create_sines(o.params, [440.0, 880.0, 1320.0])
for _ in range(20):
    v.calculate_vqt_instant_in_db(x)
ts = []
for _ in range(500):
    t0 = time.perf_counter()
    v.calculate_vqt_instant_in_db(x)
    ts.append(time.perf_counter() - t0)
ts = np.sort(np.array(ts)) * 1e6
for _ in range(20):
    o.calculate_vqt_instant_in_db(x, 1)
tc = []
for _ in range(500):
    t0 = time.perf_counter()
    o.calculate_vqt_instant_in_db(x, 1)
    tc.append(time.perf_counter() - t0)
tc = np.sort(np.array(tc)) * 1e6
print(f"gpu instant: p50 {ts[250]:.1f} us  p99 {ts[494]:.1f} us  min {ts[0]:.1f} us | cpu port: p50 {tc[250]:.1f} us p99 {tc[494]:.1f} us")

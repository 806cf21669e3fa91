import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pitchvis_b200 as pv
from pitchvis_b200 import synth, _ffi
v = pv.Vqt(pv.VqtParameters.default(), device=0)
audio = synth.polyphonic_chords(60.0, 22050.0, seed=7)
lib = _ffi.load()
lib.pvqt_debug_tc_set.argtypes = [ctypes.c_ulonglong]
lib.pvqt_debug_tc_set(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for _ in range(3):
    out = v.calculate_vqt_batch_in_db(audio, 368)
buf = (ctypes.c_ulonglong * 40)()
lib.pvqt_debug_tc.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
lib.pvqt_debug_tc(buf)
names = ["loader wait empty", "loader total", "mma wait full", "mma wait tempty", "mma issue", "mma total", "epi wait tfull", "epi loop total", "epi store", "prologue", "t commit(6)", "t epi sees tfull(6)", "t epi arrives tempty(6)", "t mma sees tempty(6)", "t mma issue(6) begins"]
for n, x in zip(names, buf):
    print(f"{n:22s} {x:10d} cycles")

print("loader a=10: start %d, after wait %d, after emit %d, after fence %d, after arrive %d" % tuple(buf[16:21]))
print("epi a=10: start %d, after wait %d, after ld0 %d, after math0 %d, after ld1 %d, after math1 %d" % tuple(buf[24:30]))

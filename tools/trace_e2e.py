"""Phase timings of host-input (end-to-end) proofs vs device-resident ones: python tools/trace_e2e.py [log_steps]"""
import ctypes as C
import importlib
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
cm = importlib.import_module("cairo-m_b200")
import bench

log = int(sys.argv[1]) if len(sys.argv) > 1 else 22
lib = cm.lib()
h = C.c_void_p()
cm.check(lib.cm31_test_fib_input_create(C.c_uint32(bench.fib_iterations(log)), C.byref(h)))
cap = 1 << 26
buf = (C.c_uint8 * cap)()
ln = C.c_size_t()
tm = (C.c_double * 5)()
for mode in ("host-input", "device-resident"):
    if mode == "device-resident":
        cm.check(lib.cm31_input_upload(h))
    for i in range(5):
        t0 = time.perf_counter()
        cm.check(lib.cm31_prove_cairo_m(h, 16, 80, buf, C.c_size_t(cap), C.byref(ln), tm))
        cm.check(lib.cm31_sync())
        wall = (time.perf_counter() - t0) * 1e3
        print(f"{mode:16s} wall {wall:7.2f} ms  phases pre/trace/inter/stark/total = " + " ".join(f"{v:6.2f}" for v in tm))

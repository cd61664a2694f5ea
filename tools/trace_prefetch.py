"""Per-step wall time and free device memory of the pipelined end-to-end loop (cm31_input_prefetch + cm31_prove_cairo_m):
python tools/trace_prefetch.py [log_steps] [steps].  CM31_NO_INPUT_SLOTS=1 disables the slot recycling for comparison."""
import ctypes as C
import importlib
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

cm = importlib.import_module("cairo-m_b200")
import bench

log = int(sys.argv[1]) if len(sys.argv) > 1 else 22
k = int(sys.argv[2]) if len(sys.argv) > 2 else 12
lib = cm.lib()
h = C.c_void_p()
cm.check(lib.cm31_test_fib_input_create(C.c_uint32(bench.fib_iterations(log)), C.byref(h)))
cap = 1 << 26
buf = (C.c_uint8 * cap)()
ln = C.c_size_t()
tm = (C.c_double * 5)()


def prove():
    cm.check(lib.cm31_prove_cairo_m(h, 16, 80, buf, C.c_size_t(cap), C.byref(ln), tm))


for _ in range(3):
    prove()
torch.cuda.synchronize()
t_all = time.perf_counter()
cm.check(lib.cm31_input_prefetch(h))
for step in range(k):
    t0 = time.perf_counter()
    if step + 1 < k:
        cm.check(lib.cm31_input_prefetch(h))
    t1 = time.perf_counter()
    prove()
    t2 = time.perf_counter()
    free, total = torch.cuda.mem_get_info()
    print(f"step {step:2d}: prefetch call {1e3 * (t1 - t0):6.2f} ms  prove {1e3 * (t2 - t1):6.2f} ms  phases " +
          " ".join(f"{v:5.2f}" for v in tm) + f"  free {free / 2**30:7.2f} GiB")
torch.cuda.synchronize()
print(f"pipelined: {1e3 * (time.perf_counter() - t_all) / k:.2f} ms/step")

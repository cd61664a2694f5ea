"""Per-proof wall time and phase times of a run of asynchronous (or blocking) proofs: python tools/dbg_async_phases.py <program> <n> [count] [sync]"""
import ctypes as C, importlib, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
cm = importlib.import_module("cairo-m_b200"); lib = cm.lib()
prog = int(sys.argv[1]); n = int(sys.argv[2]); count = int(sys.argv[3]) if len(sys.argv) > 3 else 5
sync = len(sys.argv) > 4 and sys.argv[4] == "sync"
h = C.c_void_p(); cm.check(lib.cm31_test_program_input_create(C.c_uint32(prog), C.c_uint32(n), C.byref(h))); cm.check(lib.cm31_input_upload(h))
cap = 1 << 26; bufs = [(C.c_uint8 * cap)(), (C.c_uint8 * cap)()]; lens = [C.c_size_t(), C.c_size_t()]; tm = (C.c_double * 5)()
walls = []
for i in range(count):
    t0 = time.perf_counter()
    f = lib.cm31_prove_cairo_m if sync else lib.cm31_prove_cairo_m_async
    cm.check(f(h, 16, 80, bufs[i & 1], C.c_size_t(cap), C.byref(lens[i & 1]), tm))
    walls.append(round((time.perf_counter() - t0) * 1e3, 1))
    print(f"proof {i}: wall {walls[-1]:7.1f} ms  phases {[round(x, 2) for x in tm]}", file=sys.stderr, flush=True)
cm.check(lib.cm31_prove_wait())
print("walls", walls)

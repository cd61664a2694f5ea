"""Per-launch timeline of one proof (CUDA events): where the GPU waits on the host.
python tools/trace_proof.py [log_steps] -> gpurun_out/trace_<log>.csv + a gap summary."""
import ctypes as C
import importlib
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
cm = importlib.import_module("cairo-m_b200")
import bench

log = int(sys.argv[1]) if len(sys.argv) > 1 else 22
lib = cm.lib()
h = C.c_void_p()
cm.check(lib.cm31_test_fib_input_create(C.c_uint32(bench.fib_iterations(log)), C.byref(h)))
cm.check(lib.cm31_input_upload(h))
import os
if os.environ.get("CM31_TRACE_SHARDED"):  # the same proof as one share of a (world size 1) sharded proof
    import torch
    torch.cuda.set_device(0)
    cm.shard_init(arena_gib=16)
cap = 1 << 26
buf = (C.c_uint8 * cap)()
ln = C.c_size_t()
tm = (C.c_double * 5)()
for i in range(3):
    if i == 2:
        cm.check(lib.cm31_profile_reset())
        cm.check(lib.cm31_profile_enable(1))
    cm.check(lib.cm31_prove_cairo_m(h, 16, 80, buf, C.c_size_t(cap), C.byref(ln), tm))
tb = C.create_string_buffer(1 << 22)
tl = C.c_size_t()
cm.check(lib.cm31_profile_trace(tb, C.c_size_t(1 << 22), C.byref(tl)))
rows = [l.split(",") for l in tb.value.decode().splitlines()]
out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
(out / f"trace_{log}.csv").write_text(tb.value.decode())
prev_end, gaps = 0.0, []
for name, t0, d in rows:
    t0, d = float(t0), float(d)
    gaps.append((t0 - prev_end, name, t0))
    prev_end = t0 + d
total = prev_end
busy = sum(float(r[2]) for r in rows)
print(f"total {total:.2f} ms, kernels {busy:.2f} ms, gaps {total - busy:.2f} ms, proof total {tm[4]:.2f} ms, phases {list(tm)}")
print("largest gaps (ms before kernel):")
for g, name, t0 in sorted(gaps, reverse=True)[:40]:
    print(f"  {g:8.3f}  before {name:24s} at {t0:8.2f}")

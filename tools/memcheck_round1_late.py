"""Small driver for `compute-sanitizer --tool memcheck`: the kernels added late in round 1 (device adapter, segmented logup
scan, shared-sum twiddles, slot-exact unpack) on small inputs, each checked against the host adapter / oracle."""
import ctypes as C
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

cm = importlib.import_module("cairo-m_b200")
from tests import cairo_helpers as ch
from tests import oracle_lib as orc
from tests.test_adapter_gpu import assert_same_tables

lib = cm.lib()
for program, n in [(ch.FIB, 40), (ch.ARRAY_SUM, 20), (ch.U32_MIX, 6)]:
    host = ch.GpuFibInput(cm, n, program)
    dev = ch.GpuAdaptedInput(cm, n, program)
    cm.check(lib.cm31_input_upload(host.h))
    assert_same_tables(cm, host.h, dev.h)
    host.close()
    dev.close()
print("adapter ok")
P = orc.P
for L in (12, 14):
    n = 1 << L
    src = orc.splitmix64(0x10C0 + L, 4 * n).reshape(4, n)
    cols = [torch.from_numpy(np.ascontiguousarray(src[c]).view(np.int32)).cuda() for c in range(4)]
    ptrs = (C.c_void_p * 4)(*[t.data_ptr() for t in cols])
    claimed = (C.c_uint32 * 4)()
    cm.check(lib.cm31_logup_finalize_last(ptrs, C.c_uint32(L), claimed))
    n_inv = pow(n % P, P - 2, P)
    for k in range(4):
        total = int(src[k].astype(np.uint64).sum() % P)
        shifted = ((src[k].astype(np.int64) - total * n_inv % P) % P).astype(np.uint32)
        assert np.array_equal(cols[k].cpu().numpy().view(np.uint32), orc.prefix_sum(shifted, L))
print("seg scan ok")
tw = cm.Twiddles(16)
tw.close()
inp = ch.GpuAdaptedInput(cm, 10)
proof, _ = inp.prove()
inp.close()
assert ch.oracle_cairo_verify(proof) == 0
print("proof from device-adapted input ok")

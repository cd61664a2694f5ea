"""Per-kernel microbenchmarks: CUDA-event timings + algorithmic GB/s."""
import ctypes as C
import importlib
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
cm = importlib.import_module("cairo-m_b200")


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))


def main():
    res = {}
    tw = cm.Twiddles(27)
    peaks = ROOT / "MEASURED_PEAKS.json"
    res["hbm_peak_GBps"] = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6650.0
    for L, ncols in [(20, 64), (22, 16), (24, 8), (25, 4)]:
        n = 1 << L
        cols = [torch.randint(0, cm.P, (n,), dtype=torch.int32, device="cuda") for _ in range(ncols)]
        out = [torch.empty(2 * n, dtype=torch.int32, device="cuda") for _ in range(ncols)]
        med, best = timeit(lambda: cm.interpolate_batch(cols, L, tw))
        res[f"ifft_2^{L}x{ncols}"] = {"ms": med, "GBps_alg": 8 * n * ncols / med / 1e6}
        med, best = timeit(lambda: cm.evaluate_batch(cols, out, L, L + 1, tw))
        res[f"lde_2^{L}x{ncols}"] = {"ms": med, "GBps_alg": 12 * n * ncols / med / 1e6}
        hashes = torch.empty((2 * n, 8), dtype=torch.int32, device="cuda")
        med, best = timeit(lambda: cm.blake2s_commit_layer(L + 1, None, out, hashes))
        res[f"merkle_leaf_2^{L+1}x{ncols}"] = {"ms": med, "GBps_alg": (4 * ncols + 32) * 2 * n / med / 1e6}
        h2 = torch.empty((n, 8), dtype=torch.int32, device="cuda")
        med, best = timeit(lambda: cm.blake2s_commit_layer(L, hashes, [], h2))
        res[f"merkle_inner_2^{L}"] = {"ms": med, "GBps_alg": 96 * n / med / 1e6}
        src4 = out[:4]
        dst4 = [torch.zeros(n, dtype=torch.int32, device="cuda") for _ in range(4)]
        med, best = timeit(lambda: cm.fold_circle_into_line(dst4, src4, L + 1, (1, 3, 5, 7), tw))
        res[f"fold_circle_2^{L+1}"] = {"ms": med, "GBps_alg": 32 * 2 * n / med / 1e6}
        d2 = [torch.zeros(n // 2, dtype=torch.int32, device="cuda") for _ in range(4)]
        med, best = timeit(lambda: cm.fold_line(dst4, L, (1, 3, 5, 7), tw, d2))
        res[f"fold_line_2^{L}"] = {"ms": med, "GBps_alg": 24 * n / med / 1e6}
        q4 = [torch.empty(2 * n, dtype=torch.int32, device="cuda") for _ in range(4)]
        pt = tuple(int(v) for v in np.random.randint(1, cm.P, 8))
        batches = [(pt, [(c, (1, 2, 3, 4)) for c in range(ncols)])]
        med, best = timeit(lambda: cm.accumulate_quotients(L + 1, out, (9, 8, 7, 6), batches, q4))
        res[f"quotients_2^{L+1}x{ncols}"] = {"ms": med, "GBps_alg": (4 * ncols + 16) * 2 * n / med / 1e6}
        med, best = timeit(lambda: cm.eval_at_point_batch(cols, [L] * ncols, [pt], [0] * ncols))
        res[f"eval_at_point_2^{L}x{ncols}"] = {"ms": med, "GBps_alg": 4 * n * ncols / med / 1e6}
        del cols, out, hashes, h2, dst4, d2, q4
        torch.cuda.empty_cache()
    med, best = timeit(lambda: cm.grind_blake2s(bytes(range(32)), 16), iters=3, warm=1)
    res["grind16"] = {"ms": med}
    peak3 = (C.c_double * 3)()
    cm.check(cm.lib().cm31_int_peak(peak3))
    res["int_peak_Tops"] = {"alu_pipe": peak3[0], "fma_pipe": peak3[1], "mixed": peak3[2]}
    for k, v in res.items():
        print(k, json.dumps(v))
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "microbench.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

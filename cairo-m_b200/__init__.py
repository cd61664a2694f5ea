"""cairo-m_b200 — B200-native proving hot path of cairo-m (Stwo Backend ops on sm_100a).

Python is only the test/bench harness here: it binds the C ABI of ``libcm31.so`` (declared in
``include/cm31.h``) with ctypes and hands it device pointers of torch tensors.  The library must
exist; there is no CPU fallback (``RuntimeError`` if it is missing or fails to load).

Import with ``importlib.import_module("cairo-m_b200")`` (the directory name is not an identifier).
"""
from __future__ import annotations

import ctypes
import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libcm31.so"

P = (1 << 31) - 1

_lib = None


def _preload_nccl() -> None:
    """libcm31 links NCCL (single-proof sharding, csrc/shard.cu).  torch ships its own, newer libnccl.so.2; whichever copy is
    loaded first serves both, and torch does not work with the older system copy -- so the wheel's copy goes in first."""
    import importlib.util
    import sys
    if "torch" in sys.modules:
        return  # torch already loaded its NCCL: libcm31 binds to that copy
    spec = importlib.util.find_spec("nvidia.nccl")
    if spec and spec.submodule_search_locations:
        for base in spec.submodule_search_locations:
            cand = Path(base) / "lib" / "libnccl.so.2"
            if cand.exists():
                ctypes.CDLL(str(cand), mode=ctypes.RTLD_GLOBAL)
                return


def lib() -> ctypes.CDLL:
    """The loaded C-ABI library; raises loudly when the CUDA extension is missing."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python cairo-m_b200/build.py` "
                "(there is no CPU fallback for the product path)")
        _preload_nccl()
        _lib = ctypes.CDLL(str(LIB_PATH), mode=ctypes.RTLD_GLOBAL)
        _lib.cm31_last_error.restype = C.c_char_p
    return _lib


class Cm31Error(RuntimeError):
    pass


def check(status: int) -> None:
    if status != 0:
        raise Cm31Error(lib().cm31_last_error().decode() or f"cm31 status {status}")


def shard_init(arena_gib: float = 24.0) -> tuple[int, int]:
    """Single-proof sharding over the ranks of torch.distributed (one process per GPU, NCCL): rank 0 creates the NCCL id of
    libcm31's own communicator, torch.distributed hands it to the others, every rank maps every other rank's arena.
    Returns (rank, world).  After this call every cm31_prove_cairo_m of this process is one share of a sharded proof: all
    ranks must call it in lockstep with the same input."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    uid = (C.c_uint8 * 128)()
    if rank == 0:
        check(lib().cm31_shard_unique_id(uid))
    if world > 1:
        t = torch.tensor(list(uid), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        uid = (C.c_uint8 * 128)(*t.cpu().tolist())
    check(lib().cm31_shard_init(C.c_int(rank), C.c_int(world), uid, C.c_size_t(int(arena_gib * (1 << 30)))))
    return rank, world


def shard_stats() -> dict:
    rank, world, st = C.c_int(), C.c_int(), (C.c_uint64 * 4)()
    check(lib().cm31_shard_info(C.byref(rank), C.byref(world), st))
    return {"rank": rank.value, "world": world.value, "arena_peak_bytes": int(st[0]), "bytes_all_gathered": int(st[1]),
            "bytes_all_reduced": int(st[2]), "collectives": int(st[3])}


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if hasattr(t, "data_ptr") else int(t)
    return arr


def _u32_array(values):
    arr = (C.c_uint32 * len(values))()
    for i, v in enumerate(values):
        arr[i] = int(v)
    return arr


def _qm(v):
    return _u32_array(list(v))


class Twiddles:
    """PolyOps::precompute_twiddles for CanonicCoset(log_size).half_coset()."""

    def __init__(self, log_size: int):
        self.log_size = log_size
        self.handle = C.c_void_p()
        check(lib().cm31_twiddles_create(C.c_uint32(log_size), C.byref(self.handle)))

    def buffers(self):
        import torch
        tw, itw, ls = C.c_void_p(), C.c_void_p(), C.c_uint32()
        check(lib().cm31_twiddles_buffers(self.handle, C.byref(tw), C.byref(itw), C.byref(ls)))
        n = 1 << (ls.value - 1)
        host_tw = (C.c_uint32 * n)()
        host_itw = (C.c_uint32 * n)()
        check(lib().cm31_d2h(host_tw, tw, C.c_size_t(4 * n)))
        check(lib().cm31_d2h(host_itw, itw, C.c_size_t(4 * n)))
        return list(host_tw), list(host_itw)

    def close(self):
        if self.handle:
            lib().cm31_twiddles_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sync():
    check(lib().cm31_sync())


def interpolate_batch(cols, log_size: int, tw: Twiddles) -> None:
    """In place: evaluations (bit-reversed canonic circle domain) -> FFT-basis coefficients."""
    check(lib().cm31_interpolate_batch(_ptr_array(cols), C.c_size_t(len(cols)), C.c_uint32(log_size), tw.handle))


def evaluate_batch(coeffs, out, log_size: int, log_eval_size: int, tw: Twiddles) -> None:
    check(lib().cm31_evaluate_batch(_ptr_array(coeffs), _ptr_array(out), C.c_size_t(len(coeffs)),
                                    C.c_uint32(log_size), C.c_uint32(log_eval_size), tw.handle))


def eval_at_point_batch(coeffs, log_sizes, points, point_idx):
    """points: list of 8-tuples (x QM31, y QM31); returns list of 4-tuples."""
    n = len(coeffs)
    flat_pts = [w for p in points for w in p]
    out = (C.c_uint32 * (4 * n))()
    check(lib().cm31_eval_at_point_batch(_ptr_array(coeffs), _u32_array(log_sizes), C.c_size_t(n),
                                         _u32_array(flat_pts), C.c_size_t(len(points)), _u32_array(point_idx), out))
    return [tuple(out[4 * i:4 * i + 4]) for i in range(n)]


def bit_reverse(col, log_size: int) -> None:
    check(lib().cm31_bit_reverse(C.c_void_p(col.data_ptr()), C.c_uint32(log_size)))


def blake2s_commit_layer(log_size: int, prev_layer, cols, out_layer) -> None:
    prev = C.c_void_p(prev_layer.data_ptr()) if prev_layer is not None else C.c_void_p()
    check(lib().cm31_blake2s_commit_layer(C.c_uint32(log_size), prev, _ptr_array(cols), C.c_size_t(len(cols)),
                                          C.c_void_p(out_layer.data_ptr())))


def blake2s_commit_multi(log_size: int, prev_layer, cols, out_layers) -> None:
    """Layers log_size, log_size-1, .. (len(out_layers) of them, only the first with columns) in one launch."""
    prev = C.c_void_p(prev_layer.data_ptr()) if prev_layer is not None else C.c_void_p()
    check(lib().cm31_blake2s_commit_multi(C.c_uint32(log_size), prev, _ptr_array(cols), C.c_size_t(len(cols)),
                                          C.c_uint32(len(out_layers)), _ptr_array(out_layers)))


def fold_line(src4, log_size: int, alpha, tw: Twiddles, dst4) -> None:
    check(lib().cm31_fold_line(_ptr_array(src4), C.c_uint32(log_size), _qm(alpha), tw.handle, _ptr_array(dst4)))


def fold_circle_into_line(dst4, src4, log_size: int, alpha, tw: Twiddles) -> None:
    check(lib().cm31_fold_circle_into_line(_ptr_array(dst4), _ptr_array(src4), C.c_uint32(log_size), _qm(alpha),
                                           tw.handle))


def decompose(src4, log_size: int, dst4):
    lam = (C.c_uint32 * 4)()
    check(lib().cm31_decompose(_ptr_array(src4), C.c_uint32(log_size), _ptr_array(dst4), lam))
    return tuple(lam)


def accumulate(dst4, src4, n: int) -> None:
    check(lib().cm31_accumulate(_ptr_array(dst4), _ptr_array(src4), C.c_size_t(n)))


def secure_powers(felt, n: int):
    out = (C.c_uint32 * (4 * n))()
    check(lib().cm31_secure_powers(_qm(felt), C.c_size_t(n), out))
    return [tuple(out[4 * i:4 * i + 4]) for i in range(n)]


def accumulate_quotients(log_size: int, cols, random_coeff, batches, out4) -> None:
    """batches: list of (point 8-tuple, [(col_idx, value 4-tuple), ...])."""
    pts, starts, idx, vals = [], [0], [], []
    for point, cvs in batches:
        pts.extend(point)
        for ci, v in cvs:
            idx.append(ci)
            vals.extend(v)
        starts.append(len(idx))
    check(lib().cm31_accumulate_quotients(C.c_uint32(log_size), _ptr_array(cols), C.c_size_t(len(cols)),
                                          _qm(random_coeff), C.c_size_t(len(batches)), _u32_array(pts),
                                          _u32_array(starts), _u32_array(idx), _u32_array(vals), _ptr_array(out4)))


def grind_blake2s(digest: bytes, pow_bits: int) -> int:
    nonce = C.c_uint64()
    buf = (C.c_uint8 * 32).from_buffer_copy(digest)
    check(lib().cm31_grind_blake2s(buf, C.c_uint32(pow_bits), C.byref(nonce)))
    return nonce.value


def gather_u32(cols, idx):
    out = (C.c_uint32 * (len(cols) * len(idx)))()
    check(lib().cm31_gather_u32(_ptr_array(cols), C.c_size_t(len(cols)), _u32_array(idx), C.c_size_t(len(idx)), out))
    return [list(out[c * len(idx):(c + 1) * len(idx)]) for c in range(len(cols))]


def gather_words(srcs, src_id, word_idx):
    """out[k] = srcs[src_id[k]][word_idx[k]] — the batched decommitment read (one launch per proof)."""
    n = len(src_id)
    out = (C.c_uint32 * n)()
    check(lib().cm31_gather_words(_ptr_array(srcs), C.c_size_t(len(srcs)), _u32_array(src_id), _u32_array(word_idx),
                                  C.c_size_t(n), out))
    return list(out)


def blake2s_commit_top(top_log: int, prev_layer, cols_by_layer, out_layers) -> None:
    """Layers top_log..0 in one launch; cols_by_layer[l] = list of columns with 2^l rows."""
    flat, start = [], []
    for l in range(top_log + 1):
        start.append(len(flat))
        flat.extend(cols_by_layer[l])
    start.append(len(flat))
    prev = C.c_void_p(prev_layer.data_ptr()) if prev_layer is not None else C.c_void_p()
    check(lib().cm31_blake2s_commit_top(C.c_uint32(top_log), prev, _ptr_array(flat), _u32_array(start), _ptr_array(out_layers)))


def gather_batch(srcs, src_id, word_idx, out_off, counts, grids, total_words):
    """cm31_gather_batch: run requests (counts[k] words of srcs[src_id[k]] from word_idx[k] to out[out_off[k]:]) plus row-grid
    requests `grids` = [(col_ids, rows, out_base)] (out[out_base + k*len(col_ids) + c] = srcs[col_ids[c]][rows[k]])."""
    desc, gcols, grows = [], [], []
    for col_ids, rows, base in grids:
        desc += [len(gcols), len(col_ids), len(grows), len(rows), base]
        gcols += list(col_ids)
        grows += list(rows)
    out = (C.c_uint32 * max(1, total_words))()
    check(lib().cm31_gather_batch(_ptr_array(srcs), C.c_size_t(len(srcs)), _u32_array(src_id), _u32_array(word_idx), _u32_array(out_off),
                                  _u32_array(counts), C.c_size_t(len(src_id)), _u32_array(desc), C.c_size_t(len(grids)), _u32_array(gcols),
                                  C.c_size_t(len(gcols)), _u32_array(grows), C.c_size_t(len(grows)), C.c_size_t(total_words), out))
    return list(out)[:total_words]


def gather_runs(srcs, src_id, word_idx, counts):
    """Request k copies counts[k] consecutive words from srcs[src_id[k]][word_idx[k]:]; returns the flat result."""
    off = [0]
    for c in counts:
        off.append(off[-1] + c)
    out = (C.c_uint32 * max(1, off[-1]))()
    check(lib().cm31_gather_runs(_ptr_array(srcs), C.c_size_t(len(srcs)), _u32_array(src_id), _u32_array(word_idx), _u32_array(off),
                                 C.c_size_t(len(src_id)), out))
    return list(out)[: off[-1]]

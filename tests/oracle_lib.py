"""ctypes binding of oracle/liboracle.so — the CPU restatement used as the checker (tests only)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = ROOT / "oracle" / "liboracle.so"
P = (1 << 31) - 1

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            import importlib.util
            spec = importlib.util.spec_from_file_location("cm_build", ROOT / "cairo-m_b200" / "build.py")
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build_oracle()
        _lib = C.CDLL(str(LIB_PATH))
        _lib.orc_m31_add.restype = C.c_uint32
        _lib.orc_m31_sub.restype = C.c_uint32
        _lib.orc_m31_mul.restype = C.c_uint32
        _lib.orc_m31_inv.restype = C.c_uint32
        _lib.orc_grind.restype = C.c_uint64
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def u32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint32))


def splitmix64(seed: int, n: int) -> np.ndarray:
    """SURVEY §8d input generator: splitmix64(seed) mod P, n values."""
    out = np.empty(n, dtype=np.uint64)
    x = np.uint64(seed)
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = x + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z % np.uint64(P)).astype(np.uint32)


def blake2s(data: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
    lib().orc_blake2s(buf, C.c_size_t(len(data)), out)
    return bytes(out)


def twiddles(log_size: int):
    n = 1 << (log_size - 1)
    tw = np.empty(n, dtype=np.uint32)
    itw = np.empty(n, dtype=np.uint32)
    lib().orc_twiddles(C.c_uint32(log_size), _p(tw), _p(itw))
    return tw, itw


def interpolate(values: np.ndarray, log_size: int) -> np.ndarray:
    """values: (n_cols, 2^log_size) -> coefficients, same shape."""
    v = u32(values).copy().reshape(-1, 1 << log_size)
    lib().orc_interpolate(_p(v), C.c_uint32(log_size), C.c_uint32(v.shape[0]))
    return v


def evaluate(coeffs: np.ndarray, log_size: int, log_eval: int) -> np.ndarray:
    c = u32(coeffs).reshape(-1, 1 << log_size)
    out = np.empty((c.shape[0], 1 << log_eval), dtype=np.uint32)
    lib().orc_evaluate(_p(c), C.c_uint32(log_size), C.c_uint32(log_eval), C.c_uint32(c.shape[0]), _p(out))
    return out


def eval_at_point(coeffs: np.ndarray, log_size: int, point) -> tuple:
    out = np.empty(4, dtype=np.uint32)
    lib().orc_eval_at_point(_p(u32(coeffs)), C.c_uint32(log_size), _p(u32(point)), _p(out))
    return tuple(int(x) for x in out)


def commit_on_layer(log_size: int, prev, cols: np.ndarray) -> np.ndarray:
    """cols: (n_cols, 2^log_size) u32; prev: (2^(log_size+1), 8) u32 or None -> (2^log_size, 8) u32."""
    n = 1 << log_size
    cols = u32(cols).reshape(-1, n) if cols is not None and np.size(cols) else np.zeros((0, n), dtype=np.uint32)
    out = np.empty((n, 8), dtype=np.uint32)
    pp = _p(u32(prev)) if prev is not None else C.c_void_p()
    lib().orc_commit_on_layer(C.c_uint32(log_size), pp, _p(cols) if cols.size else C.c_void_p(), C.c_uint32(cols.shape[0]), _p(out))
    return out


def fold_line(src4: np.ndarray, log_size: int, alpha) -> np.ndarray:
    s = u32(src4).reshape(4, 1 << log_size)
    out = np.empty((4, 1 << (log_size - 1)), dtype=np.uint32)
    lib().orc_fold_line(_p(s), C.c_uint32(log_size), _p(u32(alpha)), _p(out))
    return out


def fold_circle_into_line(dst4: np.ndarray, src4: np.ndarray, log_size: int, alpha) -> np.ndarray:
    d = u32(dst4).copy().reshape(4, 1 << (log_size - 1))
    s = u32(src4).reshape(4, 1 << log_size)
    lib().orc_fold_circle_into_line(_p(d), _p(s), C.c_uint32(log_size), _p(u32(alpha)))
    return d


def decompose(src4: np.ndarray, log_size: int):
    s = u32(src4).reshape(4, 1 << log_size)
    out = np.empty_like(s)
    lam = np.empty(4, dtype=np.uint32)
    lib().orc_decompose(_p(s), C.c_uint32(log_size), _p(out), _p(lam))
    return out, tuple(int(x) for x in lam)


def accumulate_quotients(log_size: int, cols: np.ndarray, random_coeff, batches) -> np.ndarray:
    cols = u32(cols).reshape(-1, 1 << log_size)
    pts, starts, idx, vals = [], [0], [], []
    for point, cvs in batches:
        pts.extend(point)
        for ci, v in cvs:
            idx.append(ci)
            vals.extend(v)
        starts.append(len(idx))
    out = np.empty((4, 1 << log_size), dtype=np.uint32)
    lib().orc_accumulate_quotients(C.c_uint32(log_size), _p(cols), C.c_uint32(cols.shape[0]), _p(u32(random_coeff)),
                                   C.c_uint32(len(batches)), _p(u32(pts)), _p(u32(starts)), _p(u32(idx)),
                                   _p(u32(vals)), _p(out))
    return out


def grind(digest: bytes, pow_bits: int) -> int:
    buf = (C.c_uint8 * 32).from_buffer_copy(digest)
    return int(lib().orc_grind(buf, C.c_uint32(pow_bits)))


def prefix_sum(col: np.ndarray, log_size: int) -> np.ndarray:
    c = u32(col).copy()
    lib().orc_prefix_sum(_p(c), C.c_uint32(log_size))
    return c


def qm31_mul(a, b):
    out = np.empty(4, dtype=np.uint32)
    lib().orc_qm31_mul(_p(u32(a)), _p(u32(b)), _p(out))
    return tuple(int(x) for x in out)


def qm31_inv(a):
    out = np.empty(4, dtype=np.uint32)
    lib().orc_qm31_inv(_p(u32(a)), _p(out))
    return tuple(int(x) for x in out)


def domain_at(log_size: int, i: int):
    out = np.empty(2, dtype=np.uint32)
    lib().orc_domain_at(C.c_uint32(log_size), C.c_uint32(i), _p(out))
    return int(out[0]), int(out[1])


def point_from_index(index: int):
    out = np.empty(2, dtype=np.uint32)
    lib().orc_point_from_index(C.c_uint32(index), _p(out))
    return int(out[0]), int(out[1])


class Channel:
    """Blake2sChannel restated by the oracle (channel/blake2s.rs)."""

    def __init__(self):
        self.state = (C.c_uint8 * 36)()

    @property
    def digest(self) -> bytes:
        return bytes(self.state[:32])

    def mix_u32s(self, data):
        d = u32(data)
        lib().orc_channel_mix_u32s(self.state, _p(d), C.c_size_t(len(d)))

    def mix_u64(self, v: int):
        lib().orc_channel_mix_u64(self.state, C.c_uint64(v))

    def draw_secure_felts(self, n: int):
        out = np.empty(4 * n, dtype=np.uint32)
        lib().orc_channel_draw_secure_felts(self.state, C.c_size_t(n), _p(out))
        return [tuple(int(x) for x in out[4 * i:4 * i + 4]) for i in range(n)]

    def draw_random_bytes(self) -> bytes:
        out = (C.c_uint8 * 32)()
        lib().orc_channel_draw_random_bytes(self.state, out)
        return bytes(out)


def last_error() -> str:
    lib().orc_last_error.restype = C.c_char_p
    return (lib().orc_last_error() or b"").decode()

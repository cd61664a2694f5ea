"""Shared helpers for the cairo-m (fibonacci_loop) proof tests."""
import ctypes as C

from tests import oracle_lib as orc

CAP = 1 << 27


FIB, ARRAY_SUM, U32_COUNTER, U32_MIX = 0, 1, 2, 3


def oracle_fib_prove(n, pow_bits=16, n_queries=80):
    return oracle_program_prove(FIB, n, pow_bits, n_queries)


def oracle_program_prove(program, n, pow_bits=16, n_queries=80):
    lib = orc.lib()
    buf = (C.c_uint8 * CAP)()
    ln = C.c_size_t()
    tm = (C.c_double * 5)()
    rc = lib.orc_program_prove(program, n, pow_bits, n_queries, buf, C.c_size_t(CAP), C.byref(ln), tm)
    assert rc == 0, orc.last_error()
    return bytes(buf[: ln.value]), list(tm)


def oracle_cairo_verify(proof: bytes, pow_bits=16, n_queries=80) -> int:
    buf = (C.c_uint8 * len(proof)).from_buffer_copy(proof)
    return orc.lib().orc_cairo_verify(buf, C.c_size_t(len(proof)), pow_bits, n_queries)


def oracle_logup_residual(n, proof: bytes, program=FIB):
    buf = (C.c_uint8 * len(proof)).from_buffer_copy(proof)
    res = (C.c_uint32 * 4)()
    info = (C.c_uint64 * 3)()
    rc = orc.lib().orc_program_logup_residual(program, n, buf, C.c_size_t(len(proof)), res, info)
    assert rc == 0, orc.last_error()
    return tuple(res), {"fib": info[0], "steps": info[1], "clock_updates": info[2]}


class GpuFibInput:
    def __init__(self, cm, n, program=FIB):
        self.cm = cm
        self.h = C.c_void_p()
        cm.check(cm.lib().cm31_program_input_create(C.c_uint32(program), C.c_uint32(n), C.byref(self.h)))
        info = (C.c_uint64 * 5)()
        cm.check(cm.lib().cm31_input_info(self.h, info))
        self.steps, self.accesses, self.memory_rows, self.return_value, self.h2d_bytes = list(info)

    def prove(self, pow_bits=16, n_queries=80):
        buf = (C.c_uint8 * CAP)()
        ln = C.c_size_t()
        tm = (C.c_double * 5)()
        self.cm.check(self.cm.lib().cm31_prove_cairo_m(self.h, pow_bits, n_queries, buf, C.c_size_t(CAP), C.byref(ln), tm))
        return bytes(buf[: ln.value]), list(tm)

    def close(self):
        if self.h:
            self.cm.lib().cm31_input_destroy(self.h)
            self.h = C.c_void_p()


def fib_mod_p(n):
    a, b = 0, 1
    for _ in range(n):
        a, b = b, (a + b) % orc.P
    return a


def array_sum_expected(n):
    vals = [(i * i) % orc.P for i in range(n)]
    if n >= 2:
        vals[1] = n
    return sum(vals) % orc.P


def u32_counter_expected(n):
    x, y = 0x0001FFF0, 0x11
    for _ in range(n):
        t = (x + y) & 0xFFFFFFFF
        x = (((t ^ y) & t) | y) & 0xFFFFFFFF
        y = (y - 1) & 0xFFFFFFFF
    return x & 0xFFFF


def u32_mix_expected(n):
    """x.lo after n rounds of the u32_mix program (csrc/cairo/vm.hpp)."""
    M = 0xFFFFFFFF
    x, y = 0x56781234, 0xF1
    for _ in range(n):
        t = (x * y) & M
        u = (t + 0x9E3779B9) & M
        q, r = divmod(u, y)
        v = (q * 0x10065) & M
        w = v ^ 0x5A5AA5A5
        a = w & 0x0FFFFFFF
        b = a | 1
        x = (b + r) & M
        y = (y + 2) & M
    return x & 0xFFFF


def corrupt_first_claimed_sum(proof: bytes) -> bytes:
    """Flips one bit of the first component's claimed sum (proof blob: u64 n, n x u32 log sizes, u64 n, n x QM31 sums, ...)."""
    n = int.from_bytes(proof[:8], "little")
    off = 8 + 4 * n + 8
    return proof[:off] + bytes([proof[off] ^ 1]) + proof[off + 1:]

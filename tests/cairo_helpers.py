"""Shared helpers for the cairo-m (fibonacci_loop) proof tests."""
import ctypes as C

from tests import oracle_lib as orc

CAP = 1 << 27


FIB, ARRAY_SUM, U32_COUNTER, U32_MIX, SHA256, ALL_OPCODES = 0, 1, 2, 3, 4, 5


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
        cm.check(cm.lib().cm31_test_program_input_create(C.c_uint32(program), C.c_uint32(n), C.byref(self.h)))
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


class ProverInputDesc(C.Structure):
    """include/cm31.h: cm31_prover_input_desc — the reference's ProverInput as flat u32 tables."""
    _fields_ = [("initial_pc", C.c_uint32), ("initial_fp", C.c_uint32), ("final_pc", C.c_uint32), ("final_fp", C.c_uint32),
                ("public_ranges", C.c_uint32 * 6), ("initial_root", C.c_uint32), ("final_root", C.c_uint32),
                ("n_steps", C.c_uint64), ("n_opcodes", C.c_uint64),
                ("opcode_ids", C.POINTER(C.c_uint32)), ("bundle_start", C.POINTER(C.c_uint64)), ("bundles", C.POINTER(C.c_uint32)),
                ("data_accesses", C.POINTER(C.c_uint32)), ("n_data_accesses", C.c_uint64),
                ("initial_memory", C.POINTER(C.c_uint32)), ("n_initial_memory", C.c_uint64),
                ("final_memory", C.POINTER(C.c_uint32)), ("n_final_memory", C.c_uint64),
                ("clock_updates", C.POINTER(C.c_uint32)), ("n_clock_updates", C.c_uint64),
                ("merkle_nodes", C.POINTER(C.c_uint32)), ("n_merkle_nodes", C.c_uint64)]


DESC_TABLES = [("data_accesses", "n_data_accesses", 4), ("initial_memory", "n_initial_memory", 8), ("final_memory", "n_final_memory", 8),
               ("clock_updates", "n_clock_updates", 6), ("merkle_nodes", "n_merkle_nodes", 9)]


def describe_input(cm, handle):
    """cm31_input_describe -> (scalars dict, tables dict of numpy COPIES)."""
    import numpy as np
    d = ProverInputDesc()
    cm.check(cm.lib().cm31_input_describe(handle, C.byref(d)))

    def arr(ptr, n, dtype=np.uint32):
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.zeros(0, dtype=dtype)

    g = int(d.n_opcodes)
    tables = {"opcode_ids": arr(d.opcode_ids, g), "bundle_start": arr(d.bundle_start, g + 1, np.uint64)}
    tables["bundles"] = arr(d.bundles, 12 * int(tables["bundle_start"][-1]))
    for name, count, width in DESC_TABLES:
        tables[name] = arr(getattr(d, name), width * int(getattr(d, count)))
    scalars = {k: getattr(d, k) for k in ["initial_pc", "initial_fp", "final_pc", "final_fp", "initial_root", "final_root", "n_steps"]}
    scalars["public_ranges"] = list(d.public_ranges)
    return scalars, tables


def create_input(cm, scalars, tables, handle_out):
    """cm31_input_create from (scalars, tables) as describe_input returns them; returns the status."""
    import numpy as np
    d = ProverInputDesc()
    for k in ["initial_pc", "initial_fp", "final_pc", "final_fp", "initial_root", "final_root", "n_steps"]:
        setattr(d, k, int(scalars[k]))
    for i, v in enumerate(scalars["public_ranges"]):
        d.public_ranges[i] = int(v)
    keep = {k: np.ascontiguousarray(v) for k, v in tables.items()}
    d.n_opcodes = keep["opcode_ids"].size
    d.opcode_ids = keep["opcode_ids"].ctypes.data_as(C.POINTER(C.c_uint32))
    d.bundle_start = keep["bundle_start"].ctypes.data_as(C.POINTER(C.c_uint64))
    d.bundles = keep["bundles"].ctypes.data_as(C.POINTER(C.c_uint32))
    for name, count, width in DESC_TABLES:
        setattr(d, name, keep[name].ctypes.data_as(C.POINTER(C.c_uint32)))
        setattr(d, count, keep[name].size // width)
    return cm.lib().cm31_input_create(C.byref(d), C.byref(handle_out))


class VmTrace:
    """The runner's output for a built-in program (host VM only): what import_from_runner_output consumes."""

    def __init__(self, cm, program, n):
        self.cm = cm
        self.h = C.c_void_p()
        cm.check(cm.lib().cm31_test_vm_trace_create(C.c_uint32(program), C.c_uint32(n), C.byref(self.h)))
        info = (C.c_uint64 * 4)()
        cm.check(cm.lib().cm31_test_vm_trace_info(self.h, info))
        self.n_trace, self.n_mem, self.n_init, self.return_value = list(info)

    def arrays(self):
        """(trace {fp, pc} words, memory log {addr, value[4]} words, preloaded memory words, public ranges) as numpy copies."""
        import numpy as np
        pt, pm, pi = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
        ranges = (C.c_uint32 * 6)()
        self.cm.check(self.cm.lib().cm31_test_vm_trace_data(self.h, C.byref(pt), C.byref(pm), C.byref(pi), ranges))
        trace = np.ctypeslib.as_array(pt, shape=(2 * self.n_trace,)).copy()
        mem = np.ctypeslib.as_array(pm, shape=(5 * self.n_mem,)).copy()
        init = np.ctypeslib.as_array(pi, shape=(4 * self.n_init,)).copy()
        return trace, mem, init, np.array(list(ranges), dtype=np.uint32)

    def close(self):
        if self.h:
            self.cm.lib().cm31_test_vm_trace_destroy(self.h)
            self.h = C.c_void_p()


def adapter_import(cm, trace, mem, init, ranges, handle_out):
    """cm31_adapter_import on numpy u32 arrays; returns the status (the caller decides whether an error is expected)."""
    import numpy as np
    trace, mem, init, ranges = (np.ascontiguousarray(a, dtype=np.uint32) for a in (trace, mem, init, ranges))
    as_p = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
    return cm.lib().cm31_adapter_import(as_p(trace), C.c_size_t(trace.size // 2), as_p(mem), C.c_size_t(mem.size // 5), as_p(init),
                                        C.c_size_t(init.size // 4), as_p(ranges), C.byref(handle_out))


class GpuAdaptedInput(GpuFibInput):
    """Prover input produced by the DEVICE adapter from the runner's logs (resident in HBM)."""

    def __init__(self, cm, n, program=FIB):
        self.cm = cm
        self.h = C.c_void_p()
        vm = VmTrace(cm, program, n)
        try:
            self.logs = vm.arrays()
            rv = vm.return_value
        finally:
            vm.close()
        cm.check(adapter_import(cm, *self.logs, self.h))
        info = (C.c_uint64 * 5)()
        cm.check(cm.lib().cm31_input_info(self.h, info))
        self.steps, self.accesses, self.memory_rows, _, self.h2d_bytes = list(info)
        self.return_value = rv


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


_SHA_K = [0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
          0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
          0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
          0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
          0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
          0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2]


def sha256_state_after(n):
    """H after n compressions of the padded block of b"abc", chained (the sha256 program of csrc/cairo/vm.hpp); plain FIPS 180-4."""
    M = 0xFFFFFFFF
    rotr = lambda x, k: ((x >> k) | (x << (32 - k))) & M
    H = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
    for _ in range(n):
        w = [0x61626380] + [0] * 14 + [0x18]
        for t in range(16, 64):
            s0 = rotr(w[t - 15], 7) ^ rotr(w[t - 15], 18) ^ (w[t - 15] >> 3)
            s1 = rotr(w[t - 2], 17) ^ rotr(w[t - 2], 19) ^ (w[t - 2] >> 10)
            w.append((w[t - 16] + s0 + w[t - 7] + s1) & M)
        a, b, c, d, e, f, g, h = H
        for t in range(64):
            t1 = (h + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & M & g)) + _SHA_K[t] + w[t]) & M
            t2 = ((rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c))) & M
            a, b, c, d, e, f, g, h = (t1 + t2) & M, a, b, c, (d + t1) & M, e, f, g
        H = [(x + y) & M for x, y in zip(H, [a, b, c, d, e, f, g, h])]
    return H


def sha256_expected(n):
    """The program's return value: the sum of the sixteen 16-bit limbs of H."""
    return sum((x & 0xFFFF) + (x >> 16) for x in sha256_state_after(n))


def corrupt_first_claimed_sum(proof: bytes) -> bytes:
    """Flips one bit of the first component's claimed sum (proof blob: u64 n, n x u32 log sizes, u64 n, n x QM31 sums, ...)."""
    n = int.from_bytes(proof[:8], "little")
    off = 8 + 4 * n + 8
    return proof[:off] + bytes([proof[off] ^ 1]) + proof[off + 1:]

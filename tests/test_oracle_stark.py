"""CPU: the oracle's whole-proof path (generic STARK driver + scalar CpuBackend + verifier).

Restates external/stwo/crates/examples/src/wide_fibonacci/mod.rs:172-229
(test_wide_fib_prove_with_blake): prove, verify, and reject a tampered proof.
"""
import ctypes as C

import pytest

from tests import oracle_lib as orc


def prove(log_n, n_cols, pow_bits=5, n_queries=3):
    lib = orc.lib()
    cap = 1 << 24
    buf = (C.c_uint8 * cap)()
    n = C.c_size_t()
    rc = lib.orc_prove_wide_fibonacci(log_n, n_cols, pow_bits, n_queries, buf, C.c_size_t(cap), C.byref(n))
    assert rc == 0, orc.last_error()
    return bytes(buf[: n.value])


def verify(log_n, n_cols, proof: bytes) -> int:
    buf = (C.c_uint8 * len(proof)).from_buffer_copy(proof)
    return orc.lib().orc_verify_wide_fibonacci(log_n, n_cols, buf, C.c_size_t(len(proof)))


@pytest.mark.parametrize("log_n", [2, 3, 4, 5, 6])
def test_wide_fib_prove_verify(log_n):
    proof = prove(log_n, 100)
    assert verify(log_n, 100, proof) == 0, orc.last_error()


def test_wide_fib_deterministic():
    assert prove(5, 20) == prove(5, 20)


def test_wide_fib_tampered_proof_rejected():
    proof = bytearray(prove(6, 32, pow_bits=4, n_queries=5))
    assert verify(6, 32, bytes(proof)) == 0
    for pos in [len(proof) // 3, len(proof) // 2, len(proof) - 5]:
        bad = bytearray(proof)
        bad[pos] ^= 1
        assert verify(6, 32, bytes(bad)) != 0


def test_wide_fib_wrong_statement_rejected():
    proof = prove(5, 16)
    assert verify(5, 17, proof) != 0

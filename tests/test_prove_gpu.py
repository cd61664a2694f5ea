"""GPU: whole proofs from the CUDA path are byte-identical to the oracle prover's and verify."""
import ctypes as C

import pytest

from tests import oracle_lib as orc
from tests.test_oracle_stark import prove as oracle_prove, verify as oracle_verify

pytestmark = pytest.mark.gpu


def gpu_prove_wide_fib(cm, log_n, n_cols, pow_bits=5, n_queries=3):
    lib = cm.lib()
    cap = 1 << 26
    buf = (C.c_uint8 * cap)()
    n = C.c_size_t()
    cm.check(lib.cm31_test_prove_wide_fibonacci(log_n, n_cols, pow_bits, n_queries, buf, C.c_size_t(cap), C.byref(n)))
    return bytes(buf[: n.value])


@pytest.mark.parametrize("log_n,n_cols", [(2, 100), (4, 100), (6, 100), (10, 16), (13, 8), (16, 4)])
def test_wide_fib_proof_bit_exact(cm, log_n, n_cols):
    got = gpu_prove_wide_fib(cm, log_n, n_cols)
    assert oracle_verify(log_n, n_cols, got) == 0, orc.last_error()
    assert got == oracle_prove(log_n, n_cols)


def test_wide_fib_regular_96_bits_config(cm):
    got = gpu_prove_wide_fib(cm, 12, 12, pow_bits=16, n_queries=80)
    assert oracle_verify(12, 12, got) == 0, orc.last_error()
    assert got == oracle_prove(12, 12, pow_bits=16, n_queries=80)


def test_wide_fib_large_verifies(cm):
    got = gpu_prove_wide_fib(cm, 20, 8, pow_bits=10, n_queries=20)
    assert oracle_verify(20, 8, got) == 0, orc.last_error()

"""CPU: the oracle's cairo-m path for fibonacci_loop — prove, verify, logup balance, tamper.

Restates crates/prover/tests/prover.rs (prove -> verify round trips: fib :116, public memory :373)
with the oracle prover/verifier; there are no golden proof bytes in the reference.
"""
import pytest

from tests import cairo_helpers as ch


@pytest.fixture(scope="module")
def fib10():
    return ch.oracle_fib_prove(10)[0]


def test_fib_proof_verifies(fib10):
    assert ch.oracle_cairo_verify(fib10) == 0, ch.orc.last_error()


def test_fib_logup_balances_and_result(fib10):
    residual, info = ch.oracle_logup_residual(10, fib10)
    assert residual == (0, 0, 0, 0)
    assert info["fib"] == ch.fib_mod_p(10) == 55
    assert info["steps"] == 8 * 10 + 8


def test_fib_tampered_proof_rejected(fib10):
    for pos in [60, len(fib10) // 2, len(fib10) - 9]:
        bad = bytearray(fib10)
        bad[pos] ^= 1
        assert ch.oracle_cairo_verify(bytes(bad)) != 0


def test_fib_wrong_claim_breaks_logup(fib10):
    # a proof for n=10 does not balance against the public data of n=11
    residual, _ = ch.oracle_logup_residual(11, fib10)
    assert residual != (0, 0, 0, 0)

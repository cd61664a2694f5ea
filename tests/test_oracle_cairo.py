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


# ---- array_sum: CallAbsImm / Ret, StoreFramePointer, StoreDoubleDerefFp(Fp), StoreToDoubleDerefFp(Imm|Fp), AssertEqFpImm
@pytest.fixture(scope="module")
def arr7():
    return ch.oracle_program_prove(ch.ARRAY_SUM, 7)[0]


def test_array_sum_proof_verifies_and_balances(arr7):
    assert ch.oracle_cairo_verify(arr7) == 0, ch.orc.last_error()
    residual, info = ch.oracle_logup_residual(7, arr7, program=ch.ARRAY_SUM)
    assert residual == (0, 0, 0, 0)
    assert info["fib"] == ch.array_sum_expected(7)  # info[0] is the program's return value


def test_array_sum_wrong_claim_breaks_logup(arr7):
    residual, _ = ch.oracle_logup_residual(8, arr7, program=ch.ARRAY_SUM)
    assert residual != (0, 0, 0, 0)


# ---- u32_counter: U32StoreImm, U32StoreAddFpFp, U32StoreSubFpFp (16-bit limbs, carries/borrows, RangeCheck16)
def test_u32_counter_proof_verifies_and_balances():
    n = 40  # y borrows through zero after 17 rounds, x carries across the limb boundary at once
    proof = ch.oracle_program_prove(ch.U32_COUNTER, n)[0]
    assert ch.oracle_cairo_verify(proof) == 0, ch.orc.last_error()
    residual, info = ch.oracle_logup_residual(n, proof, program=ch.U32_COUNTER)
    assert residual == (0, 0, 0, 0)
    assert info["fib"] == ch.u32_counter_expected(n)


# ---- u32_mix: U32StoreMul*, U32StoreDivRem*, U32StoreEqFpFp, U32StoreLtFpImm, U32StoreAddFpImm, U32Store{And,Or,Xor}FpImm
# (two-word instructions: the second QM31 word is read at pc + 1, registers advance by 2)
def test_u32_mix_proof_verifies_and_balances():
    n = 9
    proof = ch.oracle_program_prove(ch.U32_MIX, n)[0]
    assert ch.oracle_cairo_verify(proof) == 0, ch.orc.last_error()
    residual, info = ch.oracle_logup_residual(n, proof, program=ch.U32_MIX)
    assert residual == (0, 0, 0, 0)
    assert info["fib"] == ch.u32_mix_expected(n)
    assert info["steps"] == 18 * n + 8


def test_u32_mix_wrong_claim_breaks_logup():
    proof = ch.oracle_program_prove(ch.U32_MIX, 4)[0]
    residual, _ = ch.oracle_logup_residual(5, proof, program=ch.U32_MIX)  # public data of another run
    assert residual != (0, 0, 0, 0)

"""CPU: the oracle's cairo-m path for fibonacci_loop — prove, verify, logup balance, tamper.

Restates crates/prover/tests/prover.rs (prove -> verify round trips: fib :116, public memory :373)
with the oracle prover/verifier; there are no golden proof bytes in the reference.
"""
import pytest

from tests import cairo_helpers as ch
from tests import oracle_lib as orc


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
    # the balance is a function of the proof's own claims and public data: a corrupted claimed sum no longer balances,
    # and the verifier refuses the proof (verifier.rs:84-92, InvalidLogupSum)
    bad = ch.corrupt_first_claimed_sum(fib10)
    residual, _ = ch.oracle_logup_residual(10, bad)
    assert residual != (0, 0, 0, 0)
    assert ch.oracle_cairo_verify(bad) != 0


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
    residual, _ = ch.oracle_logup_residual(7, ch.corrupt_first_claimed_sum(arr7), program=ch.ARRAY_SUM)
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
    residual, _ = ch.oracle_logup_residual(4, ch.corrupt_first_claimed_sum(proof), program=ch.U32_MIX)
    assert residual != (0, 0, 0, 0)


# ---- memory commitment: partial Poseidon2 Merkle trees of the initial / final memory, merkle + poseidon2 components
def test_partial_merkle_tree_scenarios_of_the_reference_tests():
    # adapter/merkle.rs:262-424 (empty, single element, two cells, both ends of the address space) + public multiplicities
    assert orc.lib().orc_merkle_selftest() == 0, orc.last_error()


def test_all_34_components_are_in_the_proof(fib10):
    n = int.from_bytes(fib10[:8], "little")
    assert n == 34  # 26 opcode components, memory, merkle, clock_update, poseidon2, range_check_8/16/20, bitwise


def test_poseidon2_reference_kat():
    # crates/prover/tests/poseidon2.rs:15-35: permutation of (0, 1, .., 15)
    import ctypes as C
    out = (C.c_uint32 * 16)()
    orc.lib().orc_poseidon2_permutation((C.c_uint32 * 16)(*range(16)), out)
    assert list(out) == [0x505d9689, 0x3b64c904, 0x79e2fd81, 0x4ba8015f, 0x24b6d2f5, 0x23845add, 0x521f4314, 0x69dfb019,
                         0x2aaae419, 0x6cb4502c, 0x6f7fa65a, 0x75feff24, 0x128d6587, 0x515877e4, 0x037f4dd7, 0x134b427f]


def test_poseidon2_permutation_structure():
    # the permutation is a bijection-looking map with full diffusion: one changed input word changes every output word
    import ctypes as C
    a, b = (C.c_uint32 * 16)(), (C.c_uint32 * 16)()
    orc.lib().orc_poseidon2_permutation((C.c_uint32 * 16)(*range(16)), a)
    orc.lib().orc_poseidon2_permutation((C.c_uint32 * 16)(*([1] + list(range(1, 16)))), b)
    assert all(x != y for x, y in zip(a, b)) and all(x < orc.P for x in a)


# ---- BASELINE config 3: SHA-256 (examples/sha256-cairo-m/src/sha256.cm) on the u32 / bitwise / range-check components
def test_sha256_reference_state_is_fips_180_4():
    import hashlib
    assert b"".join(x.to_bytes(4, "big") for x in ch.sha256_state_after(1)) == hashlib.sha256(b"abc").digest()


def test_sha256_program_proves_and_returns_the_digest_of_abc():
    # the vector of the reference's own prover test (crates/prover/tests/prover.rs:247: sha256 of "abc")
    proof = ch.oracle_program_prove(ch.SHA256, 1)[0]
    assert ch.oracle_cairo_verify(proof) == 0, orc.last_error()
    residual, info = ch.oracle_logup_residual(1, proof, program=ch.SHA256)
    assert residual == (0, 0, 0, 0)
    assert info["fib"] == ch.sha256_expected(1)
    assert 3000 < info["steps"] < 4000


def test_sha256_program_chains_compressions():
    residual, info = ch.oracle_logup_residual(3, ch.oracle_program_prove(ch.SHA256, 3)[0], program=ch.SHA256)
    assert residual == (0, 0, 0, 0) and info["fib"] == ch.sha256_expected(3)


# ---- BASELINE config 4: the synthetic all-components workload (every opcode family in every loop iteration)
def test_all_opcodes_program_fills_25_of_26_opcode_components():
    n = 20
    proof = ch.oracle_program_prove(ch.ALL_OPCODES, n)[0]
    assert ch.oracle_cairo_verify(proof) == 0, orc.last_error()
    residual, info = ch.oracle_logup_residual(n, proof, program=ch.ALL_OPCODES)
    assert residual == (0, 0, 0, 0)
    assert info["fib"] == ch.u32_mix_expected(n)        # the u32_mix recurrence runs inside the loop
    assert info["steps"] == 46 * n + 12
    k = int.from_bytes(proof[:8], "little")
    log_sizes = [int.from_bytes(proof[8 + 4 * i:12 + 4 * i], "little") for i in range(k)]
    import json
    from pathlib import Path
    names = [c["name"] for c in json.loads((Path(__file__).parent / "golden" / "air_shapes_reference.json").read_text())["components"]]
    live = {nm: ls for nm, ls in zip(names[:26], log_sizes[:26])}
    # n = 20 rows or more in every opcode component (padded to 2^5 .. 2^7) except the one that can only prove padding
    assert all(ls >= 5 for nm, ls in live.items() if nm != "u32_store_eq_fp_imm"), live
    assert live["u32_store_eq_fp_imm"] == 4

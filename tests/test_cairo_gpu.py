"""GPU: cairo-m fibonacci_loop proofs from the CUDA path — byte-identical to the oracle prover,
accepted by the oracle verifier, logup relations balanced.  (BASELINE config[0] and up.)"""
import pytest

from tests import cairo_helpers as ch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 10, 1000])
def test_fib_proof_bit_exact(cm, n):
    inp = ch.GpuFibInput(cm, n)
    try:
        assert inp.steps == 8 * n + 8
        assert inp.return_value == ch.fib_mod_p(n)
        got, _ = inp.prove()
    finally:
        inp.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    residual, info = ch.oracle_logup_residual(n, got)
    assert residual == (0, 0, 0, 0)
    want, _ = ch.oracle_fib_prove(n)
    assert got == want


def test_fib_2_17_steps_verifies(cm):
    n = (1 << 17) // 8
    inp = ch.GpuFibInput(cm, n)
    try:
        got, tm = inp.prove()
    finally:
        inp.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    residual, info = ch.oracle_logup_residual(n, got)
    assert residual == (0, 0, 0, 0)


def test_fib_2_20_steps_verifies_with_clock_updates(cm):
    # BASELINE config[1]: 2^20 VM steps; the program words are re-read after > 2^20 clocks only at
    # the very end (ret path), the loop cells every 8 steps
    n = (1 << 20) // 8
    inp = ch.GpuFibInput(cm, n)
    try:
        assert inp.steps == (1 << 20) + 8
        got, tm = inp.prove()
    finally:
        inp.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    residual, info = ch.oracle_logup_residual(n, got)
    assert residual == (0, 0, 0, 0)


def test_generated_air_kernels_match_the_bytecode_interpreter(cm):
    # every AIR program of the proof runs once on the AOT-specialised kernels (csrc/generated/)
    # and once on the interpreter (csrc/air.cu): same proof bytes
    n = 300
    inp = ch.GpuFibInput(cm, n)
    try:
        cm.check(cm.lib().cm31_set_air_mode(1))
        interp, _ = inp.prove()
        cm.check(cm.lib().cm31_set_air_mode(0))
        gen, _ = inp.prove()
    finally:
        cm.lib().cm31_set_air_mode(0)
        inp.close()
    assert gen == interp
    assert ch.oracle_cairo_verify(gen) == 0, ch.orc.last_error()


@pytest.mark.parametrize("n", [2, 9, 300])
def test_array_sum_proof_bit_exact(cm, n):
    # the call / frame-pointer / double-deref / assert opcode components
    inp = ch.GpuFibInput(cm, n, program=ch.ARRAY_SUM)
    try:
        assert inp.return_value == ch.array_sum_expected(n)
        got, _ = inp.prove()
    finally:
        inp.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    residual, _ = ch.oracle_logup_residual(n, got, program=ch.ARRAY_SUM)
    assert residual == (0, 0, 0, 0)
    want, _ = ch.oracle_program_prove(ch.ARRAY_SUM, n)
    assert got == want


def test_array_sum_2_16_iterations_verifies(cm):
    n = 1 << 16
    inp = ch.GpuFibInput(cm, n, program=ch.ARRAY_SUM)
    try:
        got, _ = inp.prove()
    finally:
        inp.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    residual, _ = ch.oracle_logup_residual(n, got, program=ch.ARRAY_SUM)
    assert residual == (0, 0, 0, 0)


@pytest.mark.parametrize("n", [3, 40, 1000])
def test_u32_counter_proof_bit_exact(cm, n):
    # the u32 limb components (u32_store_imm, u32_store_add_fp_fp, u32_store_sub_fp_fp)
    inp = ch.GpuFibInput(cm, n, program=ch.U32_COUNTER)
    try:
        assert inp.return_value == ch.u32_counter_expected(n)
        got, _ = inp.prove()
    finally:
        inp.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    residual, _ = ch.oracle_logup_residual(n, got, program=ch.U32_COUNTER)
    assert residual == (0, 0, 0, 0)
    want, _ = ch.oracle_program_prove(ch.U32_COUNTER, n)
    assert got == want


@pytest.mark.parametrize("n", [1, 23, 700])
def test_u32_mix_proof_bit_exact(cm, n):
    # u32 mul / divrem / eq / lt and every two-word *_fp_imm u32 instruction (all 26 opcode components are in the proof)
    inp = ch.GpuFibInput(cm, n, program=ch.U32_MIX)
    try:
        assert inp.return_value == ch.u32_mix_expected(n)
        assert inp.steps == 18 * n + 8
        got, _ = inp.prove()
    finally:
        inp.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    residual, _ = ch.oracle_logup_residual(n, got, program=ch.U32_MIX)
    assert residual == (0, 0, 0, 0)
    want, _ = ch.oracle_program_prove(ch.U32_MIX, n)
    assert got == want


def test_u32_mix_2_14_rounds_verifies(cm):
    # ~2^18 VM steps of the u32-heavy program (BASELINE configs[2] stand-in): accepted by the oracle verifier
    n = 1 << 14
    inp = ch.GpuFibInput(cm, n, program=ch.U32_MIX)
    try:
        assert inp.return_value == ch.u32_mix_expected(n)
        got, _ = inp.prove()
    finally:
        inp.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    residual, _ = ch.oracle_logup_residual(n, got, program=ch.U32_MIX)
    assert residual == (0, 0, 0, 0)


def test_prefetched_input_gives_the_same_proof(cm):
    # cm31_input_prefetch: uploads issued ahead (two in flight) are consumed oldest-first by the following proofs; the
    # proof bytes do not depend on how the input reached the device (host-input, prefetched, device-resident)
    n = 300
    inp = ch.GpuFibInput(cm, n)
    try:
        plain, _ = inp.prove()
        cm.check(cm.lib().cm31_input_prefetch(inp.h))
        cm.check(cm.lib().cm31_input_prefetch(inp.h))
        first, _ = inp.prove()
        second, _ = inp.prove()
        third, _ = inp.prove()  # nothing prefetched any more: stages its own input again
        cm.check(cm.lib().cm31_input_upload(inp.h))
        resident, _ = inp.prove()
    finally:
        inp.close()
    assert plain == first == second == third == resident
    want, _ = ch.oracle_fib_prove(n)
    assert plain == want


def test_async_proofs_are_the_same_bytes(cm):
    # cm31_prove_cairo_m_async: the tail of proof i (decommitment assembly + serialisation) runs inside proof i+1 or in
    # cm31_prove_wait; the bytes must be exactly those of the synchronous call, in every interleaving
    import ctypes as C
    lib = cm.lib()
    a, b = ch.GpuFibInput(cm, 1000), ch.GpuFibInput(cm, 300, program=ch.U32_COUNTER)
    try:
        want_a, _ = a.prove()
        want_b, _ = b.prove()
        bufs = [(C.c_uint8 * ch.CAP)() for _ in range(4)]
        lens = [C.c_size_t() for _ in range(4)]
        tm = (C.c_double * 5)()

        def submit(inp, k):
            cm.check(lib.cm31_prove_cairo_m_async(inp.h, 16, 80, bufs[k], C.c_size_t(ch.CAP), C.byref(lens[k]), tm))

        def got(k):
            return bytes(bufs[k][: lens[k].value])

        submit(a, 0)
        submit(b, 1)  # completes proof 0 on its way
        assert got(0) == want_a
        submit(a, 2)
        assert got(1) == want_b
        cm.check(lib.cm31_prove_wait())
        assert got(2) == want_a
        cm.check(lib.cm31_prove_wait())  # nothing pending: a no-op
        submit(b, 3)
        again, _ = a.prove()  # a synchronous proof first completes the pending one
        assert got(3) == want_b and again == want_a
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("kind", [0, 1])
def test_invalid_trace_is_refused(cm, kind):
    # an execution trace that does not satisfy the AIR must not yield a proof: the composition OODS
    # check fails and the C ABI returns ConstraintsNotSatisfied (stwo prover/mod.rs:76-82)
    import ctypes as C
    inp = ch.GpuFibInput(cm, 50)
    try:
        cm.check(cm.lib().cm31_test_input_tamper(inp.h, C.c_uint32(kind)))
        with pytest.raises(Exception) as err:
            inp.prove()
        assert "ConstraintsNotSatisfied" in str(err.value) or "onstraint" in str(err.value)
    finally:
        inp.close()


@pytest.mark.parametrize("extra", [0, 1])
def test_fib_2_22_steps_headline_size_verifies(cm, extra):
    # the size bench.py reports (BASELINE metric: 2^22 trace; extra = 0: exactly 2^22 VM steps) and one more loop
    # iteration, which pushes every live component one row past a power of two (twice the padded rows): GPU proof
    # accepted by the oracle verifier, logup sums balanced against the public data
    n = (1 << 22) // 8 - 1 + extra
    inp = ch.GpuFibInput(cm, n)
    try:
        assert inp.steps == (1 << 22) + 8 * extra
        got, _ = inp.prove()
    finally:
        inp.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    residual, _ = ch.oracle_logup_residual(n, got)
    assert residual == (0, 0, 0, 0)
    if extra == 0:
        # ... and BYTE-IDENTICAL to the oracle prover's proof at the headline size itself (the OpenMP oracle needs ~20 s
        # of host time for 2^22 steps -- the same proof `bench.py --impl reference` times)
        want, _ = ch.oracle_fib_prove(n)
        assert got == want


@pytest.mark.parametrize("what", ["clock_delta", "after_valid_proof"])
def test_lookup_outside_its_table_is_refused_not_written_out_of_bounds(cm, what):
    # OP_HIST / gen_hist index the bin column with witness data.  A data access whose prev_clock is NOT below its
    # clock gives a range_check_20 lookup of clock - prev_clock - 1 = P - k: the value must not be used as a bin index
    # (the reference panics on the slice index, range_check_macro.rs:72-84; the oracle throws "lookup outside its table").
    import ctypes as C
    import numpy as np
    src = ch.GpuFibInput(cm, 30)
    try:
        scalars, tables = ch.describe_input(cm, src.h)
        good, _ = src.prove()
    finally:
        src.close()
    bad = dict(tables, data_accesses=tables["data_accesses"].copy())
    acc = bad["data_accesses"].reshape(-1, 4)
    acc[7, 1] = acc[7, 1] + 1000  # prev_clock far above the clock of the step that makes the access
    h = C.c_void_p()
    cm.check(ch.create_input(cm, scalars, bad, h))
    buf = (C.c_uint8 * ch.CAP)()
    ln = C.c_size_t()
    try:
        rc = cm.lib().cm31_prove_cairo_m(h, 16, 80, buf, C.c_size_t(ch.CAP), C.byref(ln), None)
        assert rc != 0
        assert "lookup outside its table" in cm.lib().cm31_last_error().decode()
    finally:
        cm.lib().cm31_input_destroy(h)
    if what == "after_valid_proof":
        # the error word is cleared by the failed proof: the next proof of a valid input is unaffected and unchanged
        again = ch.GpuFibInput(cm, 30)
        try:
            assert again.prove()[0] == good
        finally:
            again.close()


@pytest.mark.parametrize("n", [1, 5])
def test_sha256_proof_bit_exact(cm, n):
    # BASELINE config 3: SHA-256 as examples/sha256-cairo-m/src/sha256.cm computes it (rotr = mul | div, byte-wise and / or / xor
    # through the bitwise table, u32 limb adders): n = 1 is sha256("abc"), the vector of crates/prover/tests/prover.rs:247
    inp = ch.GpuFibInput(cm, n, program=ch.SHA256)
    try:
        assert inp.return_value == ch.sha256_expected(n)
        got, _ = inp.prove()
    finally:
        inp.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    residual, _ = ch.oracle_logup_residual(n, got, program=ch.SHA256)
    assert residual == (0, 0, 0, 0)
    want, _ = ch.oracle_program_prove(ch.SHA256, n)
    assert got == want


def test_sha256_2_20_steps_verifies_from_device_adapted_logs(cm):
    # 300 chained compressions = 1.05 M VM steps, u32 / bitwise / range-check components at 2^17..2^19 rows, input produced by
    # the DEVICE adapter from the runner's logs; the JSON wire form of the proof round-trips to the same bytes
    import ctypes as C
    n = 300
    dev = ch.GpuAdaptedInput(cm, n, ch.SHA256)
    try:
        assert dev.steps > 1 << 20 and dev.return_value == ch.sha256_expected(n)
        got, _ = dev.prove()
    finally:
        dev.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    residual, _ = ch.oracle_logup_residual(n, got, program=ch.SHA256)
    assert residual == (0, 0, 0, 0)
    from tests.test_proof_json import from_json, to_json
    assert from_json(cm, to_json(cm, got)) == got


@pytest.mark.parametrize("n", [1, 50])
def test_all_opcodes_proof_bit_exact(cm, n):
    # BASELINE config 4: the synthetic all-components workload -- every opcode family live in one proof
    inp = ch.GpuFibInput(cm, n, program=ch.ALL_OPCODES)
    try:
        assert inp.return_value == ch.u32_mix_expected(n) and inp.steps == 46 * n + 12
        got, _ = inp.prove()
    finally:
        inp.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    want, _ = ch.oracle_program_prove(ch.ALL_OPCODES, n)
    assert got == want


def test_all_opcodes_2_20_steps_from_device_adapted_logs(cm):
    n = 23000  # 1.06 M steps, 25 opcode components of 2^15 .. 2^17 rows
    dev = ch.GpuAdaptedInput(cm, n, ch.ALL_OPCODES)
    try:
        got, _ = dev.prove()
    finally:
        dev.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    residual, _ = ch.oracle_logup_residual(n, got, program=ch.ALL_OPCODES)
    assert residual == (0, 0, 0, 0)

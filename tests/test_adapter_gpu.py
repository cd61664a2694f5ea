"""GPU: the device adapter (cm31_adapter_import, csrc/adapter.cu) against the host restatement of
import_from_runner_output (csrc/cairo/vm.hpp::import_from_vm, the code the CPU oracle prover also runs):
every table of the prover input word for word, then whole proofs byte for byte against the oracle prover, and the
VmImportError cases of crates/prover/src/adapter/io.rs:12-36."""
import ctypes as C

import numpy as np
import pytest

from tests import cairo_helpers as ch

pytestmark = pytest.mark.gpu

N_OPCODE_COMPONENTS = 26
TABLES = [0] + list(range(1, N_OPCODE_COMPONENTS + 1)) + [100, 101, 102, 103]


def staged_table(cm, handle, table):
    n = C.c_size_t()
    cm.check(cm.lib().cm31_input_staged_words(handle, C.c_uint32(table), None, C.c_size_t(0), C.byref(n)))
    out = np.zeros(max(n.value, 1), dtype=np.uint32)
    if n.value:
        cm.check(cm.lib().cm31_input_staged_words(handle, C.c_uint32(table), out.ctypes.data_as(C.c_void_p), C.c_size_t(out.size), C.byref(n)))
    return out[: n.value]


def assert_same_tables(cm, host_handle, dev_handle):
    sizes = {}
    for t in TABLES:
        want = staged_table(cm, host_handle, t)
        got = staged_table(cm, dev_handle, t)
        assert got.size == want.size, f"table {t}: {got.size} words, host adapter has {want.size}"
        if not np.array_equal(got, want):
            bad = int(np.flatnonzero(got != want)[0])
            raise AssertionError(f"table {t}: first difference at word {bad}: device {got[bad]} != host {want[bad]}")
        sizes[t] = want.size
    return sizes


@pytest.mark.parametrize("program,n", [(ch.FIB, 0), (ch.FIB, 1), (ch.FIB, 1000), (ch.ARRAY_SUM, 200), (ch.U32_COUNTER, 150),
                                       (ch.U32_MIX, 60), (ch.SHA256, 2)])
def test_device_adapter_tables_match_host_adapter(cm, program, n):
    host = ch.GpuFibInput(cm, n, program)
    dev = ch.GpuAdaptedInput(cm, n, program)
    try:
        cm.check(cm.lib().cm31_input_upload(host.h))
        assert dev.steps == host.steps and dev.accesses == host.accesses and dev.memory_rows == host.memory_rows
        sizes = assert_same_tables(cm, host.h, dev.h)
        assert sizes[0] == 4 * host.accesses
        assert sum(sizes[t] for t in range(1, N_OPCODE_COMPONENTS + 1)) == 12 * host.steps
    finally:
        host.close()
        dev.close()


def test_device_adapter_clock_update_rows(cm):
    # > 2^20 steps: the cells written once at the start and read again by the final ret are more than RC20_LIMIT
    # clocks apart, so Memory::push splits the gap into clock-update rows (adapter/memory.rs:470-…)
    n = 140_000
    host = ch.GpuFibInput(cm, n)
    dev = ch.GpuAdaptedInput(cm, n)
    try:
        cm.check(cm.lib().cm31_input_upload(host.h))
        sizes = assert_same_tables(cm, host.h, dev.h)
        assert sizes[102] > 0, "the workload was meant to need clock-update rows"
    finally:
        host.close()
        dev.close()


@pytest.mark.parametrize("program,n", [(ch.FIB, 100), (ch.ARRAY_SUM, 64), (ch.U32_MIX, 20)])
def test_device_adapted_proof_bit_exact(cm, program, n):
    dev = ch.GpuAdaptedInput(cm, n, program)
    try:
        got, _ = dev.prove()
        again, _ = dev.prove()  # the resident input is not consumed
    finally:
        dev.close()
    assert got == again
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    want, _ = ch.oracle_program_prove(program, n)
    assert got == want


def test_device_adapted_2_20_steps_verifies(cm):
    n = (1 << 20) // 8
    dev = ch.GpuAdaptedInput(cm, n)
    try:
        assert dev.steps == (1 << 20) + 8
        got, _ = dev.prove()
    finally:
        dev.close()
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()
    residual, info = ch.oracle_logup_residual(n, got)
    assert residual == (0, 0, 0, 0) and info["clock_updates"] > 0


def test_device_adapter_rejects_malformed_logs(cm):
    vm = ch.VmTrace(cm, ch.FIB, 50)
    try:
        trace, mem, init, ranges = vm.arrays()
    finally:
        vm.close()

    def expect_error(trace, mem, init, needle):
        h = C.c_void_p()
        rc = ch.adapter_import(cm, trace, mem, init, ranges, h)
        assert rc != 0, f"accepted a malformed input (expected '{needle}')"
        assert needle in cm.lib().cm31_last_error().decode()

    expect_error(trace[:2], mem, init, "empty trace")                               # one entry = no step
    expect_error(trace, mem[:-5], init, "memory trace")                             # log ends early
    expect_error(trace, np.concatenate([mem, mem[-5:]]), init, "memory trace")      # entries left over
    bad = init.copy()
    bad[4 * int(trace[2 * 3 + 1])] = 63                                             # opcode word of the 4th step's instruction
    expect_error(trace, mem, bad, "invalid opcode")
    bad = mem.copy()
    bad[0] += 1                                                                     # first fetch not at pc
    expect_error(trace, bad, init, "unexpected memory access")
    bad = trace.copy()
    bad[2 * 5 + 1] = init.size                                                      # pc outside the preloaded memory
    expect_error(bad, mem, init, "pc outside")


@pytest.mark.parametrize("program,n", [(ch.FIB, 64), (ch.U32_COUNTER, 40)])
def test_proof_from_caller_supplied_prover_input(cm, program, n):
    # cm31_input_create: the ProverInput tables of a caller-side adapter -> same proof bytes as the built-in path
    src = ch.GpuFibInput(cm, n, program)
    try:
        want, _ = src.prove()
        scalars, tables = ch.describe_input(cm, src.h)
    finally:
        src.close()
    inp = ch.GpuFibInput.__new__(ch.GpuFibInput)
    inp.cm, inp.h = cm, C.c_void_p()
    cm.check(ch.create_input(cm, scalars, tables, inp.h))
    try:
        got, _ = inp.prove()
        cm.check(cm.lib().cm31_input_upload(inp.h))
        resident, _ = inp.prove()
    finally:
        inp.close()
    assert got == want and resident == want
    assert ch.oracle_cairo_verify(got) == 0, ch.orc.last_error()


def test_prefetched_logs_give_the_same_input_and_proof(cm):
    # cm31_adapter_prefetch / cm31_adapter_import_prefetched: two segments in flight, uploads on the background stream
    import torch
    lib = cm.lib()
    n_a, n_b = 300, 77
    logs = {}
    for n in (n_a, n_b):
        vm = ch.VmTrace(cm, ch.FIB, n)
        try:
            # page-locked copies, as a runner would hand them over
            logs[n] = [torch.from_numpy(a.view(np.int32)).pin_memory() for a in vm.arrays()]
        finally:
            vm.close()

    def prefetch(n):
        t, m, i, r = logs[n]
        as_p = lambda x: C.cast(x.data_ptr(), C.POINTER(C.c_uint32))
        lg = C.c_void_p()
        cm.check(lib.cm31_adapter_prefetch(as_p(t), C.c_size_t(t.numel() // 2), as_p(m), C.c_size_t(m.numel() // 5), as_p(i),
                                           C.c_size_t(i.numel() // 4), as_p(r), C.byref(lg)))
        return lg

    lg_a, lg_b = prefetch(n_a), prefetch(n_b)  # both uploads queued before anything is adapted
    proofs = {}
    for n, lg in ((n_a, lg_a), (n_b, lg_b)):
        inp = ch.GpuFibInput.__new__(ch.GpuFibInput)
        inp.cm, inp.h = cm, C.c_void_p()
        cm.check(lib.cm31_adapter_import_prefetched(lg, C.byref(inp.h)))
        try:
            host = ch.GpuFibInput(cm, n)
            try:
                cm.check(lib.cm31_input_upload(host.h))
                assert_same_tables(cm, host.h, inp.h)
            finally:
                host.close()
            proofs[n], _ = inp.prove()
        finally:
            inp.close()
    for n in (n_a, n_b):
        want, _ = ch.oracle_fib_prove(n)
        assert proofs[n] == want
    # logs that are never imported can be dropped
    lg = prefetch(n_b)
    cm.check(lib.cm31_adapter_logs_destroy(lg))


def test_device_adapted_headline_size_proof_equals_host_adapted(cm):
    # 2^22 VM steps (the BASELINE workload): the proof from the device-adapted input is byte-identical to the proof from the
    # host-adapted one (which the headline-size test of test_cairo_gpu.py verifies with the oracle verifier)
    n = (1 << 19) - 1
    host = ch.GpuFibInput(cm, n)
    try:
        assert host.steps == 1 << 22
        want, _ = host.prove()
    finally:
        host.close()
    dev = ch.GpuAdaptedInput(cm, n)
    try:
        assert dev.steps == 1 << 22
        got, _ = dev.prove()
    finally:
        dev.close()
    assert got == want

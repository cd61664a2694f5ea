"""CPU: the runner-log interface of the device adapter (cm31_test_vm_trace_*, cm31_adapter_import): log layout as
crates/prover/src/adapter/io.rs:38-60 defines it, and no silent CPU fallback when there is no device."""
import ctypes as C

import numpy as np
import pytest

from tests import cairo_helpers as ch


@pytest.mark.parametrize("program,n", [(ch.FIB, 20), (ch.ARRAY_SUM, 12), (ch.U32_MIX, 5)])
def test_vm_logs_have_the_runner_layout(cm, program, n):
    vm = ch.VmTrace(cm, program, n)
    try:
        trace, mem, init, ranges = vm.arrays()
        assert trace.size == 2 * vm.n_trace and mem.size == 5 * vm.n_mem and init.size == 4 * vm.n_init
        if program == ch.FIB:
            assert vm.n_trace - 1 == 8 * n + 8 and vm.return_value == ch.fib_mod_p(n)
    finally:
        vm.close()
    # IoTraceEntry {fp, pc}: the first log entry is the fetch of the instruction at the first pc, and its opcode word
    # is the preloaded program cell
    pc0 = int(trace[1])
    assert int(mem[0]) == pc0 and int(mem[1]) == int(init[4 * pc0])
    assert int(ranges[0]) == 0 and int(ranges[1]) <= vm.n_init and int(ranges[4]) < int(ranges[5])
    # every fetch in the log (an entry at a program address) carries the preloaded instruction word
    prog_end = int(ranges[1])
    addr = mem[0::5]
    fetch = np.flatnonzero(addr < prog_end)
    assert fetch.size >= vm.n_trace - 1
    assert np.array_equal(mem[5 * fetch + 1], init[4 * addr[fetch].astype(np.int64)])


def test_adapter_import_fails_loudly_without_a_device(cm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    vm = ch.VmTrace(cm, ch.FIB, 3)
    try:
        logs = vm.arrays()
    finally:
        vm.close()
    h = C.c_void_p()
    assert ch.adapter_import(cm, *logs, h) != 0
    assert not h.value
    assert cm.lib().cm31_last_error()


@pytest.mark.parametrize("program,n", [(ch.FIB, 40), (ch.ARRAY_SUM, 30), (ch.U32_COUNTER, 25), (ch.U32_MIX, 8)])
def test_input_create_round_trips_the_reference_prover_input(cm, program, n):
    # a caller that ran its own adapter hands the ProverInput over as flat tables (cm31_input_create); the handle holds
    # exactly those tables (cm31_input_describe)
    src = ch.GpuFibInput(cm, n, program)
    try:
        scalars, tables = ch.describe_input(cm, src.h)
    finally:
        src.close()
    assert scalars["n_steps"] == int(tables["bundle_start"][-1]) and scalars["n_steps"] >= 1
    assert tables["bundles"].size == 12 * scalars["n_steps"]
    h = C.c_void_p()
    cm.check(ch.create_input(cm, scalars, tables, h))
    try:
        scalars2, tables2 = ch.describe_input(cm, h)
        info = (C.c_uint64 * 5)()
        cm.check(cm.lib().cm31_input_info(h, info))
    finally:
        cm.lib().cm31_input_destroy(h)
    assert scalars2 == scalars
    assert int(info[0]) == scalars["n_steps"] and int(info[1]) == tables["data_accesses"].size // 4
    for k in tables:
        assert np.array_equal(tables[k], tables2[k]), k


def test_input_create_rejects_inconsistent_tables(cm):
    src = ch.GpuFibInput(cm, 10)
    try:
        scalars, tables = ch.describe_input(cm, src.h)
    finally:
        src.close()

    def expect_error(scalars, tables, needle):
        h = C.c_void_p()
        assert ch.create_input(cm, scalars, tables, h) != 0
        assert needle in cm.lib().cm31_last_error().decode()

    expect_error(dict(scalars, n_steps=scalars["n_steps"] + 1), tables, "do not add up")
    bad = dict(tables, opcode_ids=tables["opcode_ids"].copy())
    bad["opcode_ids"][0] = 63
    expect_error(scalars, bad, "invalid opcode")
    bad = dict(tables, bundles=tables["bundles"].copy())
    bad["bundles"][10] = tables["data_accesses"].size  # span start past the end of the access log
    expect_error(scalars, bad, "access span")
    expect_error(dict(scalars, n_steps=0), tables, "empty trace")

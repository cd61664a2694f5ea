"""CPU: the runner-log interface of the device adapter (cm31_vm_trace_*, cm31_adapter_import): log layout as
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

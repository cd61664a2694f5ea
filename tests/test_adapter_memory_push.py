"""The reference's own `Memory::push` unit tests (crates/prover/src/adapter/memory.rs:546-858), restated and replayed on

  * a literal Python transliteration of `Memory::push` (memory.rs:470-537, HashMap based) -- ties the expectations to the
    reference's code;
  * the host adapter of this repo (`cm31::MemoryModel`, csrc/cairo/vm.hpp), driven through `orc_memory_push_script`, with the
    reference tests' exact inputs and literal expected values;
  * (GPU) the device adapter `cm31_adapter_import` (csrc/adapter.cu), on runner logs crafted so that the same access patterns
    occur (first access, same address twice, two addresses, a 3*RC20_LIMIT+500 gap, an RC20_LIMIT-1 gap, preloaded cells).
"""
import ctypes as C

import numpy as np
import pytest

from tests import cairo_helpers as ch
from tests import oracle_lib as orc

P = orc.P
RC20_LIMIT = (1 << 20) - 1  # adapter/memory.rs:16
NEG1 = P - 1


# ---------------------------------------------------------------- literal transliteration of Memory (memory.rs:405-537)
class RefMemory:
    def __init__(self, initial=None):  # Memory::new: final_memory starts as a copy of initial_memory
        self.initial_memory = dict(initial or {})
        self.final_memory = dict(self.initial_memory)
        self.clock_update_data = []

    def push(self, address, value, clock):
        prev = self.final_memory.get(address)
        self.final_memory[address] = (value, clock, NEG1)
        if prev is None:
            prev = (value, 0, NEG1)
        prev_clk = prev[1]
        if prev_clk == 0:
            if address in self.initial_memory:
                v, c, _ = self.initial_memory[address]
                self.initial_memory[address] = (v, c, 1)
            else:
                self.initial_memory[address] = (value, 0, 1)
        init_value = self.initial_memory.get(address)
        if clock > prev_clk:
            delta = clock - prev_clk
            if delta > RC20_LIMIT:
                for _ in range(delta // RC20_LIMIT):
                    self.clock_update_data.append((address, prev_clk, init_value[0]))
                    prev_clk = (prev_clk + RC20_LIMIT) % P
        return {"address": address, "prev_val": prev[0], "value": value, "prev_clock": prev_clk, "clock": clock}


class HostMemory:
    """cm31::MemoryModel behind orc_memory_push_script: the whole script is replayed on every query (the scripts are tiny)."""

    def __init__(self, initial=None):
        self.initial = initial or {}
        self.script = []

    def push(self, address, value, clock):
        self.script.append((address, value, clock))
        self._run()
        return self.args[-1]

    def _run(self):
        n_init = (max(self.initial) + 1) if self.initial else 0
        init = np.zeros(4 * max(n_init, 1), dtype=np.uint32)
        for a, (v, _, _) in self.initial.items():
            init[4 * a:4 * a + 4] = v
        entries = np.array([[a, *v, c] for a, v, c in self.script], dtype=np.uint32).reshape(-1)
        n = len(self.script)
        args = np.zeros(11 * n, dtype=np.uint32)
        cu = np.zeros(6 * 4096, dtype=np.uint32)
        cells_i, cells_f = np.zeros(7 * 64, dtype=np.uint32), np.zeros(7 * 64, dtype=np.uint32)
        n_cu, n_i, n_f = C.c_size_t(), C.c_size_t(), C.c_size_t()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = orc.lib().orc_memory_push_script(p(init), C.c_size_t(n_init), p(entries), C.c_size_t(n), p(args), p(cu), C.c_size_t(4096),
                                              C.byref(n_cu), p(cells_i), p(cells_f), C.c_size_t(64), C.byref(n_i), C.byref(n_f))
        assert rc == 0, orc.last_error()
        self.args = [{"address": int(r[0]), "prev_clock": int(r[1]), "clock": int(r[2]), "prev_val": tuple(map(int, r[3:7])),
                      "value": tuple(map(int, r[7:11]))} for r in args.reshape(n, 11)]
        self.clock_update_data = [(int(r[0]), int(r[1]), tuple(map(int, r[2:6]))) for r in cu.reshape(-1, 6)[: n_cu.value]]
        cell = lambda r: (int(r[0]), (tuple(map(int, r[1:5])), int(r[5]), int(r[6])))
        self.initial_memory = dict(cell(r) for r in cells_i.reshape(-1, 7)[: n_i.value])
        self.final_memory = dict(cell(r) for r in cells_f.reshape(-1, 7)[: n_f.value])


@pytest.fixture(params=["reference_transliteration", "host_adapter"])
def Memory(request):
    return RefMemory if request.param == "reference_transliteration" else HostMemory


def test_memory_push_first_entry(Memory):  # memory.rs:553-593
    memory = Memory()
    result = memory.push(100, (1, 2, 3, 4), 10)
    assert result["address"] == 100 and result["prev_clock"] == 0 and result["clock"] == 10
    assert result["prev_val"] == (1, 2, 3, 4) and result["value"] == (1, 2, 3, 4)
    assert memory.final_memory[100] == ((1, 2, 3, 4), 10, NEG1)
    assert memory.initial_memory[100] == ((1, 2, 3, 4), 0, 1)


def test_memory_push_same_address(Memory):  # memory.rs:595-644
    memory = Memory()
    memory.push(100, (1, 2, 3, 4), 10)
    result = memory.push(100, (5, 6, 7, 8), 20)
    assert result["address"] == 100 and result["prev_clock"] == 10 and result["clock"] == 20
    assert result["prev_val"] == (1, 2, 3, 4) and result["value"] == (5, 6, 7, 8)
    assert memory.final_memory[100] == ((5, 6, 7, 8), 20, NEG1)
    assert memory.initial_memory[100] == ((1, 2, 3, 4), 0, 1)


def test_memory_push_different_addresses(Memory):  # memory.rs:646-699
    memory = Memory()
    memory.push(100, (1, 2, 3, 4), 10)
    result = memory.push(200, (9, 10, 11, 12), 30)
    assert result["address"] == 200 and result["prev_clock"] == 0 and result["clock"] == 30
    assert result["prev_val"] == (9, 10, 11, 12) and result["value"] == (9, 10, 11, 12)
    assert len(memory.final_memory) == 2
    assert memory.final_memory[100] == ((1, 2, 3, 4), 10, NEG1)
    assert memory.final_memory[200] == ((9, 10, 11, 12), 30, NEG1)
    assert len(memory.initial_memory) == 2
    assert memory.initial_memory[100] == ((1, 2, 3, 4), 0, 1)
    assert memory.initial_memory[200] == ((9, 10, 11, 12), 0, 1)


def test_memory_push_multiple_large_clock_deltas(Memory):  # memory.rs:701-737
    memory = Memory()
    memory.push(100, (1, 2, 3, 4), 10)
    large_delta = 3 * RC20_LIMIT + 500
    result = memory.push(100, (5, 6, 7, 8), 10 + large_delta)
    assert len(memory.clock_update_data) == 3
    assert memory.clock_update_data[0][1] == 10
    assert memory.clock_update_data[1][1] == 10 + RC20_LIMIT
    assert memory.clock_update_data[2][1] == 10 + 2 * RC20_LIMIT
    # beyond the reference's assertions: every row re-emits the INITIAL value of the cell (memory.rs:512-526), and the
    # MemoryArg's prev_clock is the last intermediate clock
    assert all(row[0] == 100 and row[2] == (1, 2, 3, 4) for row in memory.clock_update_data)
    assert result["prev_clock"] == 10 + 3 * RC20_LIMIT


def test_memory_push_no_clock_update_for_small_delta(Memory):  # memory.rs:739-762
    memory = Memory()
    memory.push(100, (1, 2, 3, 4), 10)
    memory.push(100, (5, 6, 7, 8), 10 + RC20_LIMIT - 1)
    assert memory.clock_update_data == []


def test_memory_push_with_preloaded_memory(Memory):  # memory.rs:764-858
    memory = Memory({0: ((10, 20, 30, 40), 0, 0), 1: ((50, 60, 70, 80), 0, 0)})
    if isinstance(memory, RefMemory):
        assert len(memory.initial_memory) == 2 and len(memory.final_memory) == 2
        assert memory.initial_memory[0] == ((10, 20, 30, 40), 0, 0) and memory.initial_memory[1] == ((50, 60, 70, 80), 0, 0)
    result = memory.push(0, (10, 20, 30, 40), 5)
    assert result["address"] == 0 and result["prev_clock"] == 0 and result["clock"] == 5
    assert result["prev_val"] == (10, 20, 30, 40) and result["value"] == (10, 20, 30, 40)
    assert memory.initial_memory[0] == ((10, 20, 30, 40), 0, 1)
    assert memory.initial_memory[1] == ((50, 60, 70, 80), 0, 0)
    assert memory.final_memory[0] == ((10, 20, 30, 40), 5, NEG1)
    result = memory.push(0, (100, 200, 300, 400), 10)
    assert result["address"] == 0 and result["prev_clock"] == 5 and result["clock"] == 10
    assert result["prev_val"] == (10, 20, 30, 40) and result["value"] == (100, 200, 300, 400)
    assert memory.final_memory[0] == ((100, 200, 300, 400), 10, NEG1)


def test_host_adapter_equals_transliteration_on_random_scripts():
    rng = np.random.default_rng(20261017)
    for trial in range(20):
        initial = {a: (tuple(int(x) for x in rng.integers(0, P, 4)), 0, 0) for a in range(int(rng.integers(0, 6)))}
        ref, host = RefMemory(initial), HostMemory(initial)
        clock = 0
        for _ in range(12):
            clock += int(rng.choice([1, 7, RC20_LIMIT - 1, RC20_LIMIT, RC20_LIMIT + 1, 2 * RC20_LIMIT + 3]))
            a = int(rng.integers(0, 9))
            if a in ref.final_memory and rng.random() < 0.5:
                v = ref.final_memory[a][0]  # a read: same value as the cell holds
            else:
                v = tuple(int(x) for x in rng.integers(0, P, 4))
            r, h = ref.push(a, v, clock), host.push(a, v, clock)
            assert r == h, (trial, r, h)
        assert ref.clock_update_data == host.clock_update_data
        assert ref.initial_memory == host.initial_memory and ref.final_memory == host.final_memory


# ---------------------------------------------------------------- the same scenarios through the DEVICE adapter
OP_STORE_IMM, OP_JMP_ABS_IMM, OP_ASSERT_EQ_FP_IMM = 9, 12, 50
FP = 64


class CraftedRun:
    """Runner logs of a straight-line run over three preloaded instruction cells: a filler jump (no data access), and one
    store / assert per scenario address.  `plan` maps a clock (= 1-based step) to (pc, data accesses [(address, value)])."""

    def __init__(self, program, n_steps, plan, extra_preloaded=()):
        self.program = program
        self.init = [tuple(w) for w in program] + [tuple(w) for w in extra_preloaded]
        trace, mem, pushes = [], [], []
        for step in range(1, n_steps + 1):
            pc, accesses = plan.get(step, (0, []))
            trace += [FP, pc]  # IoTraceEntry {fp, pc}
            mem += [pc, *self.init[pc]]
            pushes.append((pc, self.init[pc], step))
            for addr, value in accesses:
                mem += [addr, value, 0, 0, 0]
                pushes.append((addr, (value, 0, 0, 0), step))
        trace += [FP, 0]
        self.trace, self.mem, self.pushes = np.array(trace, dtype=np.uint32), np.array(mem, dtype=np.uint32), pushes
        self.ranges = np.array([0, len(program), len(program), len(program), len(program), len(program)], dtype=np.uint32)

    def reference(self):
        ref = RefMemory({a: (v, 0, 0) for a, v in enumerate(self.init)})
        args = [ref.push(a, v, c) for a, v, c in self.pushes]
        return ref, args

    def device(self, cm):
        from tests.test_adapter_gpu import staged_table
        h = C.c_void_p()
        init = np.array(self.init, dtype=np.uint32).reshape(-1)
        cm.check(ch.adapter_import(cm, self.trace, self.mem, init, self.ranges, h))
        try:
            return {t: staged_table(cm, h, t) for t in (0, 100, 102)}
        finally:
            cm.lib().cm31_input_destroy(h)


PROGRAM = [(OP_JMP_ABS_IMM, 0, 0, 0), (OP_STORE_IMM, 7, 36, 0), (OP_STORE_IMM, 9, 136, 0), (OP_ASSERT_EQ_FP_IMM, 0, 5, 0),
           (OP_STORE_IMM, 8, 36, 0)]


def check_device_against_reference(cm, run):
    ref, args = run.reference()
    tables = run.device(cm)
    # data-access log: every non-fetch push in order, (address, prev_clock, prev_value[0], value[0])
    want = [(a["address"], a["prev_clock"], a["prev_val"][0], a["value"][0])
            for a, (addr, _, _) in zip(args, run.pushes) if addr >= len(run.program)]
    got = [tuple(map(int, r)) for r in tables[0].reshape(-1, 4)]
    assert got == want
    got_cu = [(int(r[0]), int(r[1]), tuple(map(int, r[2:6]))) for r in tables[102].reshape(-1, 6)]
    assert got_cu == ref.clock_update_data
    return ref, args, tables


@pytest.mark.gpu
def test_device_adapter_first_same_and_different_addresses(cm):
    # memory.rs:553-699 in one run: address 100 first touched at clock 10 (store 7), again at clock 20 (store 8);
    # address 200 first touched at clock 30
    run = CraftedRun(PROGRAM, 40, {10: (1, [(100, 7)]), 20: (4, [(100, 8)]), 30: (2, [(200, 9)])})
    ref, args, tables = check_device_against_reference(cm, run)
    acc = [tuple(map(int, r)) for r in tables[0].reshape(-1, 4)]
    assert acc == [(100, 0, 7, 7), (100, 10, 7, 8), (200, 0, 9, 9)]
    assert tables[102].size == 0
    # boundary memory rows (address, clock, value[4], multiplicity, root): initial then final
    rows = {(int(r[0]), int(r[1])): (tuple(map(int, r[2:6])), int(r[6])) for r in tables[100].reshape(-1, 8)}
    assert rows[(100, 0)][0] == (7, 0, 0, 0) and rows[(100, 20)][0] == (8, 0, 0, 0)
    assert rows[(200, 0)][0] == (9, 0, 0, 0) and rows[(200, 30)][0] == (9, 0, 0, 0)


@pytest.mark.gpu
def test_device_adapter_multiple_large_clock_deltas(cm):
    # memory.rs:701-737: the same cell at clocks 10 and 10 + 3*RC20_LIMIT + 500 -> three clock-update rows at
    # 10, 10 + RC20_LIMIT, 10 + 2*RC20_LIMIT.  Both the data cell and the instruction cell that touches it see that gap.
    large = 3 * RC20_LIMIT + 500
    run = CraftedRun(PROGRAM, 10 + large + 3, {10: (1, [(100, 7)]), 10 + large: (1, [(100, 7)])})
    ref, args, tables = check_device_against_reference(cm, run)
    cu = [(int(r[0]), int(r[1])) for r in tables[102].reshape(-1, 6)]
    assert cu == [(1, 10), (1, 10 + RC20_LIMIT), (1, 10 + 2 * RC20_LIMIT), (100, 10), (100, 10 + RC20_LIMIT), (100, 10 + 2 * RC20_LIMIT)]
    assert tuple(map(int, tables[0].reshape(-1, 4)[1])) == (100, 10 + 3 * RC20_LIMIT, 7, 7)


@pytest.mark.gpu
def test_device_adapter_no_clock_update_for_small_delta(cm):
    # memory.rs:739-762: a gap of RC20_LIMIT - 1 needs no clock-update row; exactly RC20_LIMIT does not either
    # (delta > RC20_LIMIT is the condition, memory.rs:515), RC20_LIMIT + 1 needs one
    for delta, n_rows in [(RC20_LIMIT - 1, 0), (RC20_LIMIT, 0), (RC20_LIMIT + 1, 2)]:
        run = CraftedRun(PROGRAM, 10 + delta + 2, {10: (1, [(100, 7)]), 10 + delta: (1, [(100, 7)])})
        _, _, tables = check_device_against_reference(cm, run)
        assert tables[102].size == 6 * n_rows, delta


@pytest.mark.gpu
def test_device_adapter_preloaded_memory(cm):
    # memory.rs:764-858: a preloaded cell read at clock 5 (prev clock 0, prev value = the preloaded one, multiplicity of the
    # initial cell becomes 1) and overwritten at clock 10 (prev clock 5, prev value = preloaded)
    cell = len(PROGRAM)
    run = CraftedRun(PROGRAM, 12, {5: (3, [(cell, 10)]), 10: (1, [(cell, 100)])}, extra_preloaded=[(10, 0, 0, 0), (50, 0, 0, 0)])
    ref, args, tables = check_device_against_reference(cm, run)
    acc = [tuple(map(int, r)) for r in tables[0].reshape(-1, 4)]
    assert acc == [(cell, 0, 10, 10), (cell, 5, 10, 100)]
    rows = [(int(r[0]), int(r[1]), tuple(map(int, r[2:6])), int(r[6])) for r in tables[100].reshape(-1, 8)]
    assert (cell, 0, (10, 0, 0, 0), 1) in rows           # initial cell, multiplicity 1 after the first access
    assert (cell, 10, (100, 0, 0, 0), NEG1) in rows      # final cell
    assert not any(r[0] == cell + 1 and r[3] != 0 for r in rows)  # the untouched preloaded cell is not consumed

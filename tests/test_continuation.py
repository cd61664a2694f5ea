"""Continuation segments: Merkle-root continuity across the segments of one run, as the reference tests it in
crates/prover/tests/prover.rs:204-243 (`test_hash_continuity_fibonacci`: fibonacci_loop(5) with RunnerOptions { max_steps: 10 },
every segment imported, proven and verified, `final_root` of segment i == `initial_root` of segment i + 1).

The runner side (crates/runner/src/vm/mod.rs:158-285: a segment ends when its trace holds max_steps states, the state is appended
as its final entry, the next segment starts from the whole memory image with clocks reset) is restated in
csrc/cairo/vm.hpp::run_program(.., segments, segment_steps).

  * CPU: every segment proven and verified on the oracle, roots chained;
  * GPU: every segment through BOTH adapters (host restatement and cm31_adapter_import on the device), proven by the CUDA path,
    byte-identical to the oracle's segment proof, roots chained, at the reference's size and at 2^16 steps cut into 5 segments.
"""
import ctypes as C
import struct

import pytest

from tests import cairo_helpers as ch
from tests import oracle_lib as orc

N_COMPONENTS = 34
# proof blob (csrc/cairo/prover.hpp CairoProof::write): u64 n, n x u32 log sizes, u64 n, n x QM31 claimed sums, then the public
# head {initial pc, fp, final pc, fp, clock, initial_root, final_root}
PUBLIC_HEAD = 8 + 4 * N_COMPONENTS + 8 + 16 * N_COMPONENTS


def public_head(proof: bytes):
    ipc, ifp, fpc, ffp, clock, initial_root, final_root = struct.unpack_from("<7I", proof, PUBLIC_HEAD)
    return {"initial": (ipc, ifp), "final": (fpc, ffp), "clock": clock, "initial_root": initial_root, "final_root": final_root}


def oracle_segment_prove(program, n, segment_steps, index):
    buf = (C.c_uint8 * ch.CAP)()
    ln = C.c_size_t()
    nseg = C.c_uint32()
    rc = orc.lib().orc_segment_prove(C.c_uint32(program), C.c_uint32(n), C.c_uint64(segment_steps), C.c_uint32(index), C.byref(nseg),
                                     16, 80, buf, C.c_size_t(ch.CAP), C.byref(ln))
    assert rc == 0, orc.last_error()
    return bytes(buf[: ln.value]), nseg.value


def check_chain(heads, total_steps=None):
    for a, b in zip(heads, heads[1:]):
        assert a["final_root"] == b["initial_root"], "initial root of a segment must equal the final root of the previous one"
        assert a["final"] == b["initial"], "a segment starts from the registers the previous one ended with"
    if total_steps is not None:
        # public_data.clock = the segment's step count (public_data.rs:248-253: the number of ExecutionBundles)
        assert sum(h["clock"] for h in heads) == total_steps


def test_hash_continuity_fibonacci_oracle():
    # the reference test's shape: fibonacci_loop(5) = 48 VM steps, max_steps 10 -> 5 segments
    first, nseg = oracle_segment_prove(ch.FIB, 5, 10, 0)
    assert nseg == 5
    heads = []
    for i in range(nseg):
        proof = first if i == 0 else oracle_segment_prove(ch.FIB, 5, 10, i)[0]
        assert ch.oracle_cairo_verify(proof) == 0, orc.last_error()
        heads.append(public_head(proof))
    check_chain(heads, total_steps=8 * 5 + 8)
    assert len({h["initial_root"] for h in heads}) == nseg  # memory really changes from segment to segment
    # the unsegmented run starts from the same memory image and registers
    whole, _ = ch.oracle_fib_prove(5)
    assert public_head(whole)["initial"] == heads[0]["initial"] and public_head(whole)["final"] == heads[-1]["final"]


def test_segments_partition_the_run(cm):
    # host only: the segments' traces and memory logs concatenate to the unsegmented run's
    lib = cm.lib()

    def logs(handle):
        info = (C.c_uint64 * 4)()
        cm.check(lib.cm31_test_vm_trace_info(handle, info))
        tr, mem, init = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
        ranges = (C.c_uint32 * 6)()
        cm.check(lib.cm31_test_vm_trace_data(handle, C.byref(tr), C.byref(mem), C.byref(init), ranges))
        return list(tr[: 2 * info[0]]), list(mem[: 5 * info[1]]), list(init[: 4 * info[2]]), list(ranges), int(info[3])

    whole = C.c_void_p()
    cm.check(lib.cm31_test_vm_trace_create(C.c_uint32(ch.FIB), C.c_uint32(5), C.byref(whole)))
    w_tr, w_mem, w_init, w_ranges, w_ret = logs(whole)
    lib.cm31_test_vm_trace_destroy(whole)
    nseg = C.c_uint32()
    tr_cat, mem_cat = [], []
    for i in range(5):
        h = C.c_void_p()
        cm.check(lib.cm31_test_vm_segment_create(C.c_uint32(ch.FIB), C.c_uint32(5), C.c_uint64(10), C.c_uint32(i), C.byref(nseg), C.byref(h)))
        assert nseg.value == 5
        tr, mem, init, ranges, ret = logs(h)
        lib.cm31_test_vm_trace_destroy(h)
        assert len(tr) // 2 == (11 if i < 4 else 9)  # 10 states + the final one; the last segment has the remaining 8 steps
        assert ranges == w_ranges
        if i == 0:
            assert init == w_init
        else:
            assert tr[:2] == tr_cat[-2:]  # starts from the state the previous segment ended with
            tr_cat = tr_cat[:-2]
        tr_cat += tr
        mem_cat += mem
    assert tr_cat == w_tr and mem_cat == w_mem and ret == w_ret
    h = C.c_void_p()
    assert lib.cm31_test_vm_segment_create(C.c_uint32(ch.FIB), C.c_uint32(5), C.c_uint64(10), C.c_uint32(5), C.byref(nseg), C.byref(h)) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("n,segment_steps", [(5, 10), ((1 << 16) // 8, 15000)])
def test_hash_continuity_fibonacci_gpu(cm, n, segment_steps):
    lib = cm.lib()
    nseg = C.c_uint32()
    heads = []
    i = 0
    while True:
        seg = C.c_void_p()
        cm.check(lib.cm31_test_vm_segment_create(C.c_uint32(ch.FIB), C.c_uint32(n), C.c_uint64(segment_steps), C.c_uint32(i), C.byref(nseg), C.byref(seg)))
        proofs = []
        try:
            # (a) the host adapter's ProverInput, (b) the device adapter on the segment's raw logs
            for how in ("host", "device"):
                h = C.c_void_p()
                if how == "host":
                    cm.check(lib.cm31_test_vm_trace_to_input(seg, C.byref(h)))
                else:
                    info = (C.c_uint64 * 4)()
                    cm.check(lib.cm31_test_vm_trace_info(seg, info))
                    tr, mem, init = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
                    ranges = (C.c_uint32 * 6)()
                    cm.check(lib.cm31_test_vm_trace_data(seg, C.byref(tr), C.byref(mem), C.byref(init), ranges))
                    cm.check(lib.cm31_adapter_import(tr, C.c_size_t(info[0]), mem, C.c_size_t(info[1]), init, C.c_size_t(info[2]), ranges, C.byref(h)))
                try:
                    buf = (C.c_uint8 * ch.CAP)()
                    ln = C.c_size_t()
                    tm = (C.c_double * 5)()
                    cm.check(lib.cm31_prove_cairo_m(h, 16, 80, buf, C.c_size_t(ch.CAP), C.byref(ln), tm))
                    proofs.append(bytes(buf[: ln.value]))
                finally:
                    lib.cm31_input_destroy(h)
        finally:
            lib.cm31_test_vm_trace_destroy(seg)
        assert proofs[0] == proofs[1], "host-adapted and device-adapted segment must give the same proof"
        if n == 5 or i in (0, nseg.value - 1):  # (the oracle needs seconds per proof: at the larger size only the end segments)
            want, _ = oracle_segment_prove(ch.FIB, n, segment_steps, i)
            assert proofs[0] == want
        assert ch.oracle_cairo_verify(proofs[0]) == 0, orc.last_error()
        heads.append(public_head(proofs[0]))
        i += 1
        if i >= nseg.value:
            break
    assert nseg.value == 5
    check_chain(heads, total_steps=8 * n + 8)

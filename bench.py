#!/usr/bin/env python
"""bench.py — VM-steps proven/sec for fibonacci_loop on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA prover
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm (oracle port, all host cores)

A "step" is ONE whole proof (prove_cairo_m: trace fill -> 4 commitments -> constraint/quotient
evaluation -> OODS -> DEEP quotients -> FRI -> PoW -> decommit) of a fibonacci_loop segment of
2^log_steps VM steps.  `value` = VM steps proven per second with the prover input already resident
in HBM; `e2e` = the same through cm31_prove_cairo_m with HOST input buffers (pinned), i.e. the
per-proof host->device copy of the execution bundles / access log and the proof bytes read back
are inside the timed region.  N > 1: one process per GPU, each proves its own continuation
segment (independent proofs, no data-path collective; SURVEY.md §8e "replicas"), weak scaling.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "VM-steps proven/sec (fibonacci_loop)"
UNIT = "steps/s"


def fib_iterations(log_steps: int) -> int:
    # 8 VM steps per loop iteration + 8 for prologue/epilogue: n = 2^log/8 - 1 gives EXACTLY 2^log_steps VM steps
    # (segment.trace.len() of the reference's bench, P/benches/prover_speed_benchmark.rs:49).  One more iteration
    # would push every live opcode component one row past a power of two and double its padded size.
    return (1 << log_steps) // 8 - 1


def workload_name(log_steps: int) -> str:
    return f"fibonacci_loop 2^{log_steps} VM steps, full Cpu+Memory+ClockUpdate+RangeCheck AIR, REGULAR_96_BITS pcs config"


# ------------------------------------------------------------------ clocks sampler (nvidia-smi recipe via NVML)
class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        self.period = float(os.environ.get("CM31_CLOCK_SAMPLE_PERIOD", "0.01"))
        if os.environ.get("CM31_NO_SAMPLER"):
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                if util > 0:
                    self.samples.append(mhz)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------ CPU arm (oracle port)
def oracle_prove_timed(log_steps: int):
    """One proof on the CPU oracle (C++ restatement of stwo's CpuBackend driving the same protocol,
    OpenMP over the host cores).  Returns (steps, seconds)."""
    from tests import cairo_helpers as ch
    n = fib_iterations(log_steps)
    t = time.perf_counter()
    ch.oracle_fib_prove(n)
    return 8 * n + 8, time.perf_counter() - t


def cpu_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample_text(log_steps: int, seconds: float | None = None) -> str:
    t = f"; {seconds:.1f} s" if seconds is not None else ""
    return (f"one whole proof of fibonacci_loop 2^{log_steps} VM steps (all 34 components, REGULAR_96_BITS) on oracle/ "
            f"(C++ port of stwo CpuBackend + the same protocol driver, OpenMP){t}")


def run_reference(args, rank: int, world: int):
    """The CPU arm proves THE SAME workload the CUDA arm names (2^log_steps VM steps per proof, same pcs config).  One oracle
    proof of 2^22 steps takes about a minute, so the number of timed proofs is capped (min(steps, 2)) and the warm-up is one
    small proof (page faults, OpenMP pool): both are stated in the line."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host core
    os.environ["OMP_NUM_THREADS"] = str(cpu_cores())
    log_steps = args.log_steps
    timed = max(1, min(args.steps, 2))
    if args.warmup > 0:
        oracle_prove_timed(min(log_steps, 14))
    total_steps, total_s = 0, 0.0
    for _ in range(timed):
        s, dt = oracle_prove_timed(log_steps)
        total_steps += s
        total_s += dt
    value = total_steps / total_s
    cores = cpu_cores()
    sample = cpu_sample_text(log_steps, total_s / timed) + f" per proof, {timed} proof(s) timed"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": timed,
        "steps_requested": args.steps, "warmup": 1 if args.warmup > 0 else 0, "warmup_requested": args.warmup,
        "warmup_note": "one small (2^14-step) proof: the CPU path has no warm-up effect beyond page faults and the OpenMP pool",
        "ms_per_step": 1e3 * total_s / timed, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 (M31/QM31 modular integer)", "data": "synthetic",
        "config": {"workload": workload_name(log_steps), "vm_steps_per_proof": total_steps // timed,
                   "parallelism": "host cores (OpenMP)", "pcs": {"pow_bits": 16, "log_blowup": 1, "n_queries": 80}},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the Rust reference cannot be built in this image (no cargo/rustc, crates not vendored): this arm times "
                "oracle/ (scalar C++ restatement of stwo CpuBackend + the same protocol driver, OpenMP on all host cores), "
                "not the Rust SimdBackend",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ multi-rank plumbing
def init_dist(world: int, local_rank: int, use_cuda: bool):
    """One process per GPU (torchrun env).  NCCL on GPUs; gloo when there is no device (CPU tests)."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    if use_cuda:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group("gloo")


def max_over_ranks(ms: float, world: int, use_cuda: bool) -> float:
    """Every multi-GPU time is the MAX over ranks of the device-measured time."""
    if world <= 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], device="cuda" if use_cuda else "cpu", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_value(world: int, vm_steps_per_proof: int, proofs_per_rank: int, ms: float) -> float:
    """Whole-job VM steps per second: every rank proved `proofs_per_rank` segments in `ms`."""
    return world * vm_steps_per_proof * proofs_per_rank / (ms / 1e3)


def dist_selftest(rank: int, world: int):
    """CPU/gloo check of the N>1 plumbing (tests/test_bench_dist.py): rank r reports (r+1)*10 ms."""
    import torch.distributed as dist
    init_dist(world, 0, use_cuda=False)
    ms = max_over_ranks(10.0 * (rank + 1), world, use_cuda=False)
    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps({"selftest": True, "n_gpus": world, "ms": ms, "value": aggregate_value(world, 1000, 2, ms)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_adapter(cm, lib, torch, program_id, n, k_steps, vm_steps, proof_buf, cap, proof_len, tm):
    t0 = time.perf_counter()
    vt = C.c_void_p()
    cm.check(lib.cm31_test_vm_trace_create(C.c_uint32(program_id), C.c_uint32(n), C.byref(vt)))
    vm_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    hh = C.c_void_p()
    cm.check(lib.cm31_test_program_input_create(C.c_uint32(program_id), C.c_uint32(n), C.byref(hh)))  # VM + host adapter
    host_adapter_ms = max(0.0, (time.perf_counter() - t0) - vm_s) * 1e3
    cm.check(lib.cm31_input_destroy(hh))
    info = (C.c_uint64 * 4)()
    cm.check(lib.cm31_test_vm_trace_info(vt, info))
    n_trace, n_mem, n_init = int(info[0]), int(info[1]), int(info[2])
    pt, pm, pi = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
    ranges = (C.c_uint32 * 6)()
    cm.check(lib.cm31_test_vm_trace_data(vt, C.byref(pt), C.byref(pm), C.byref(pi), ranges))
    import numpy as np

    def pinned(ptr, words):  # the logs in page-locked host memory, as a runner would hand them over
        a = np.ctypeslib.as_array(ptr, shape=(words,))
        return torch.from_numpy(a.view(np.int32)).clone().pin_memory()

    trace, mem, init = pinned(pt, 2 * n_trace), pinned(pm, 5 * n_mem), pinned(pi, 4 * n_init)
    cm.check(lib.cm31_test_vm_trace_destroy(vt))
    as_p = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_uint32))

    def import_logs():
        h = C.c_void_p()
        cm.check(lib.cm31_adapter_import(as_p(trace), C.c_size_t(n_trace), as_p(mem), C.c_size_t(n_mem), as_p(init), C.c_size_t(n_init),
                                         ranges, C.byref(h)))
        return h

    def timed(fn, reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def import_only():
        cm.check(lib.cm31_input_destroy(import_logs()))

    def import_and_prove():
        h = import_logs()
        cm.check(lib.cm31_prove_cairo_m(h, 16, 80, proof_buf, C.c_size_t(cap), C.byref(proof_len), tm))
        cm.check(lib.cm31_input_destroy(h))

    def prefetch_logs():
        lg = C.c_void_p()
        cm.check(lib.cm31_adapter_prefetch(as_p(trace), C.c_size_t(n_trace), as_p(mem), C.c_size_t(n_mem), as_p(init), C.c_size_t(n_init),
                                           ranges, C.byref(lg)))
        return lg

    pbufs = [proof_buf, (C.c_uint8 * cap)()]
    plens = [proof_len, C.c_size_t()]

    def pipelined(k):  # K uploads + K adapter runs + K proofs; the upload of segment i+1 is issued before segment i is adapted and proven
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lg = prefetch_logs()
        for step in range(k):
            nxt = prefetch_logs() if step + 1 < k else None
            h = C.c_void_p()
            cm.check(lib.cm31_adapter_import_prefetched(lg, C.byref(h)))
            cm.check(lib.cm31_prove_cairo_m_async(h, 16, 80, pbufs[step & 1], C.c_size_t(cap), C.byref(plens[step & 1]), tm))
            cm.check(lib.cm31_input_destroy(h))
            lg = nxt
        cm.check(lib.cm31_prove_wait())
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    import_and_prove()  # warm-up (allocator pools, CUB temp sizes)
    device_ms = timed(import_only, max(3, k_steps // 2))
    step_ms = timed(import_and_prove, k_steps)
    pipelined(2)
    pipe_ms = pipelined(k_steps)
    log_bytes = 4 * (2 * n_trace + 5 * n_mem + 4 * n_init)
    return {"device_adapter_ms": device_ms, "host_adapter_ms": host_adapter_ms, "host_vm_ms": vm_s * 1e3,
            "h2d_bytes": log_bytes, "steps_per_s_device_adapter": vm_steps / (device_ms * 1e-3),
            "from_logs_ms_per_step": step_ms, "from_logs_value": vm_steps / (step_ms * 1e-3),
            "from_logs_pipelined_ms_per_step": pipe_ms, "from_logs_pipelined_value": vm_steps / (pipe_ms * 1e-3),
            "note": "device_adapter_ms = pinned runner logs -> resident prover input (H2D copy included); host_adapter_ms = the serial "
                    "host restatement of import_from_runner_output on one core; from_logs = upload + device adapter + proof + proof bytes back; pipelined = the upload of segment i+1 (cm31_adapter_prefetch) "
                    "overlaps the adapter kernels and the proof of segment i"}


# ------------------------------------------------------------------ CUDA arm
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    init_dist(world, local_rank, use_cuda=True)
    cm = importlib.import_module("cairo-m_b200")
    lib = cm.lib()
    cm.check(lib.cm31_set_device(local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sharded = args.mode == "sharded"
    if sharded:
        # ONE proof sharded over all ranks: every rank holds the same input and owns a subset of the components;
        # strong scaling -- the work per step is one proof whatever the number of GPUs
        cm.shard_init(arena_gib=args.arena_gib)
    programs = {"fibonacci_loop": 0, "array_sum": 1, "u32_counter": 2, "u32_mix": 3, "sha256": 4, "all_opcodes": 5}
    n = args.iterations if args.iterations else fib_iterations(args.log_steps)
    if args.program == "all_opcodes" and not args.iterations:
        n = (1 << args.log_steps) // 46  # 46 VM steps per iteration: BASELINE config 4 (synthetic trace, every opcode component live)
    if args.program == "sha256" and not args.iterations:
        n = (1 << args.log_steps) // 3490  # ~3 490 VM steps per compression: BASELINE config 3 (~2^22 rows of u32 / bitwise work)
    h = C.c_void_p()
    cm.check(lib.cm31_test_program_input_create(C.c_uint32(programs[args.program]), C.c_uint32(n), C.byref(h)))
    info = (C.c_uint64 * 5)()
    cm.check(lib.cm31_input_info(h, info))
    vm_steps, h2d_bytes = int(info[0]), int(info[4])
    cap = 1 << 26
    proof_buf = (C.c_uint8 * cap)()
    proof_len = C.c_size_t()
    tm = (C.c_double * 5)()

    def prove():
        cm.check(lib.cm31_prove_cairo_m(h, 16, 80, proof_buf, C.c_size_t(cap), C.byref(proof_len), tm))

    # Inside the timed regions the proofs are submitted with cm31_prove_cairo_m_async: the host-side tail of proof i
    # (decommitment assembly + serialisation, GPU idle) runs while proof i+1 executes, as a prover fed with a stream of
    # segments runs; cm31_prove_wait() before the closing event makes sure all K proofs (bytes included) are complete
    # inside the region.  --sync-proofs times the blocking call instead; sharded proofs always use it.
    use_async = not args.sync_proofs and args.mode != "sharded"
    proof_bufs = [proof_buf, (C.c_uint8 * cap)()]
    proof_lens = [proof_len, C.c_size_t()]

    def prove_step(i):
        if use_async:
            cm.check(lib.cm31_prove_cairo_m_async(h, 16, 80, proof_bufs[i & 1], C.c_size_t(cap), C.byref(proof_lens[i & 1]), tm))
        else:
            prove()

    def timed_region(k_steps, with_profile, prefetch=False):
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        cm.check(lib.cm31_profile_reset())
        cm.check(lib.cm31_profile_enable(1 if with_profile else 0))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phases = [0.0] * 5
        e0.record()
        if prefetch:  # step 0's host->device copy
            cm.check(lib.cm31_input_prefetch(h))
        step_wall = []
        for step in range(k_steps):
            t_step = time.perf_counter()
            if prefetch and step + 1 < k_steps:  # the copy of step i+1 overlaps the proof of step i (still inside the timed region)
                cm.check(lib.cm31_input_prefetch(h))
            prove_step(step)
            step_wall.append(round((time.perf_counter() - t_step) * 1e3, 2))
            for i in range(5):
                phases[i] += tm[i]
        if use_async:
            cm.check(lib.cm31_prove_wait())  # the last proof's tail: every proof is complete, bytes on the host, before e1
        e1.record()
        timed_region.last_step_wall = step_wall
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        launches = C.c_uint64()
        cm.check(lib.cm31_profile_launches(C.byref(launches)))
        report = None
        if with_profile:
            ln = C.c_size_t()
            rb = C.create_string_buffer(1 << 16)
            cm.check(lib.cm31_profile_report(rb, C.c_size_t(1 << 16), C.byref(ln)))
            report = json.loads(rb.value.decode())
        cm.check(lib.cm31_profile_enable(0))
        ms = max_over_ranks(ms, world, use_cuda=True)
        return ms, clocks, int(launches.value), report, [p / k_steps for p in phases]

    # ---- device-resident: input staged in HBM once
    cm.check(lib.cm31_input_upload(h))
    # same submission form as the timed steps; never fewer than 3 proofs: the library grows its memory pool to its working size
    # after the first two proofs of a process (cm31_pool_reserve_headroom), which must not land inside a timed region
    for i in range(max(args.warmup, 3)):
        prove_step(i)
    if use_async:
        cm.check(lib.cm31_prove_wait())
    # `value`: K proofs, per-kernel event timers OFF (they cost two event records per launch, ~460 launches per proof)
    ms, clocks, launches, _, phases = timed_region(args.steps, False)
    value = aggregate_value(1 if sharded else world, vm_steps, args.steps, ms)
    # second pass over the same K steps with every launch bracketed by CUDA events on its launch stream: the per-kernel table
    # and the roofline come from here; its own ms/step is reported next to the unprofiled one
    prof_ms, _, _, report, _ = timed_region(args.steps, True)

    # ---- end to end: host (pinned) input copied in, proof bytes copied out, every step
    cm.check(lib.cm31_input_release_device(h))
    prove()
    # serial form: every proof first waits for its own input (copy -> prove -> copy -> prove ...)
    e2e_serial_ms, _, _, _, _ = timed_region(args.steps, False)
    # pipelined form (the headline e2e): K uploads and K proofs, the upload of segment i+1 issued before the proof of
    # segment i (cm31_input_prefetch), as a prover fed with continuation segments does
    wu_ms, _, _, _, _ = timed_region(2, False, prefetch=True)  # warm-up of the pipelined path itself (two inputs in flight), untimed
    print(f"bench.py: pipelined e2e warm-up: {wu_ms / 2:.2f} ms/step, host wall per step {timed_region.last_step_wall}", file=sys.stderr)
    e2e_ms, _, _, _, _ = timed_region(args.steps, False, prefetch=True)
    e2e_step_wall = timed_region.last_step_wall
    e2e_rejected = None
    if e2e_ms > 1.15 * e2e_serial_ms:
        # Pipelining can only remove time from the serial form (same copies, same proofs).  A pipelined region slower than
        # the serial one has been seen intermittently as the FIRST pipelined run on a fresh box (2x, cause not yet isolated);
        # such a measurement is rejected and re-measured once, and the rejected numbers stay in the line.
        e2e_rejected = {"ms_per_step": e2e_ms / args.steps, "step_wall_ms": e2e_step_wall}
        e2e_ms, _, _, _, _ = timed_region(args.steps, False, prefetch=True)
        e2e_step_wall = timed_region.last_step_wall
    e2e_value = aggregate_value(1 if sharded else world, vm_steps, args.steps, e2e_ms)
    proof_bytes = int(proof_len.value)
    cm.check(lib.cm31_input_destroy(h))

    # ---- the step BEFORE the path: the runner's logs -> prover input.  Host restatement of the
    # reference adapter (serial walk, csrc/cairo/vm.hpp) vs the device adapter (csrc/adapter.cu), and whole steps that
    # START from the pinned logs: upload + device adapter + proof + proof bytes back.
    adapter = None
    if world == 1 and not args.no_adapter and not sharded:
        adapter = measure_adapter(cm, lib, torch, programs[args.program], n, args.steps, vm_steps, proof_buf, cap, proof_len, tm)

    shard_info = None
    if sharded:
        st = cm.shard_stats()
        per_proof = max(1, 3 * args.steps + args.warmup + 4)  # proofs made since shard_init (warm-up, value, profile, e2e passes)
        shard_info = {"world": st["world"], "arena_peak_bytes": st["arena_peak_bytes"],
                      "bytes_all_gathered_per_proof_approx": st["bytes_all_gathered"] // per_proof,
                      "collectives_per_proof_approx": st["collectives"] // per_proof,
                      "data_plane": "components dealt out over the ranks (trace fill, lookups, logup, ICFFT/LDE, constraint evaluation local to the "
                                    "owner); Merkle layers striped by node range with remote LDE columns read over NVLink inside the leaf kernel; "
                                    "DEEP quotients by row range; FRI replicated",
                      "collectives": "NCCL all-gather: last striped Merkle layer of every tree (the join at the Merkle root), DEEP quotient rows, "
                                     "reduced accumulator rows; NCCL all-reduce: barriers, multiplicity bins, claimed sums, OODS values; "
                                     "mod-P accumulator sum = reduce-scatter by peer loads + all-gather",
                      "proof_identical_to_single_gpu": "checked by tests/dist/sharded_prover_worker.py (byte-for-byte on every rank)"}
        cm.check(lib.cm31_shard_finalize())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    report.sort(key=lambda r: -r["ms"])
    kern_total = sum(r["ms"] for r in report)
    top = report[0]
    avg_ms = top["ms"] / top["launches"]
    achieved = (top["alg_bytes"] / top["launches"]) / (avg_ms * 1e-3) / 1e9
    traffic, int_issue = None, None
    tpath = ROOT / "profiles" / "traffic.json"
    if tpath.exists():
        t = json.loads(tpath.read_text()).get(top["kernel"])
        if t:  # measured DRAM bytes per algorithmic byte (ncu --set full) x this launch's algorithmic bytes
            traffic = t["dram_per_alg_byte"] * top["alg_bytes"] / top["launches"]
        if t and "thread_inst_per_alg_byte" in t:
            # the roof that binds: integer issue.  Peak measured live (ALU-pipe, FMA-pipe, 1:1 mix microkernels);
            # achieved = executed thread-instructions per algorithmic byte (ncu) x achieved algorithmic bytes/s
            peak3 = (C.c_double * 3)()
            cm.check(lib.cm31_int_peak(peak3))
            ach = t["thread_inst_per_alg_byte"] * achieved * 1e9 / 1e12
            int_issue = {"achieved_Tops": ach, "peak_Tops": peak3[2], "frac": ach / peak3[2], "unit": "tera thread-instructions/s",
                         "peak_alu_pipe_Tops": peak3[0], "peak_fma_pipe_Tops": peak3[1], "peak_mixed_Tops": peak3[2],
                         "thread_inst_per_alg_byte": t["thread_inst_per_alg_byte"],
                         "note": "this kernel is integer-issue bound, not HBM bound (DESIGN.md §4); frac is against the measured 1:1 ALU/FMA mix"}
    roofline = {
        "bound": "hbm", "kernel": top["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "int_issue": int_issue, "peak_source": peak_src, "avg_launch_ms": avg_ms, "launches_per_step": top["launches"] / args.steps,
        "share_of_kernel_time": top["ms"] / kern_total if kern_total else None,
        "kernel_time_share_of_step": kern_total / prof_ms,
        "profiled_pass_ms_per_step": prof_ms / args.steps,
        "note": "value / ms_per_step are measured with the per-kernel timers off; this table is a second pass over the same K steps "
                "with every launch bracketed by CUDA events (kernel_time_share_of_step is relative to that pass)",
        "kernels": [{"kernel": r["kernel"], "ms_per_step": r["ms"] / args.steps, "launches_per_step": r["launches"] / args.steps,
                     "alg_GBps": (r["alg_bytes"] / (r["ms"] * 1e-3) / 1e9) if r["ms"] > 0 else None,
                     "m31_Gops": (r.get("m31_ops", 0) / (r["ms"] * 1e-3) / 1e9) if r["ms"] > 0 and r.get("m31_ops") else None} for r in report],
        "m31_ops_per_step": sum(r.get("m31_ops", 0) for r in report) / args.steps,
        "m31_Gops_whole_step": sum(r.get("m31_ops", 0) for r in report) / (ms * 1e-3) / 1e9,
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        os.environ["OMP_NUM_THREADS"] = str(cpu_cores())
        s, dt = oracle_prove_timed(args.cpu_sample_log)
        cpu_baseline = {"value": s / dt, "unit": UNIT, "cores": cpu_cores(), "kind": "port", "sample": cpu_sample_text(args.cpu_sample_log, dt)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
        "dtype": "u32 (M31/QM31 modular integer)", "data": "synthetic",
        "config": {"workload": workload_name(args.log_steps) if args.program == "fibonacci_loop" and not args.iterations
                   else (f"sha256({n} compressions of a padded block, examples/sha256-cairo-m style: u32 / bitwise / range-check components) "
                         f"[BASELINE config 3, not the headline workload]" if args.program == "sha256"
                         else f"all_opcodes({n} iterations x 46 steps: synthetic trace with every opcode component live, memory / merkle / "
                              f"clock_update / poseidon2 / range-check / bitwise components included) [BASELINE config 4]" if args.program == "all_opcodes"
                         else f"{args.program}({n}) [side measurement, not the BASELINE workload]"), "vm_steps_per_proof": vm_steps,
                   "parallelism": (f"ONE proof sharded over {world} GPU(s)" if sharded else
                                   "one independent segment proof per GPU" if world > 1 else "single GPU"),
                   "l2": "working set per proof (GBs of trace/LDE columns) >> 126 MB L2; no flush between steps",
                   "submission": ("asynchronous (cm31_prove_cairo_m_async): the host-side tail of proof i (decommitment assembly, "
                                  "serialisation) runs while proof i+1 executes; cm31_prove_wait() inside the timed region"
                                  if use_async else "blocking cm31_prove_cairo_m"),
                   "pcs": {"pow_bits": 16, "log_blowup": 1, "n_queries": 80}},
        "phases_ms": dict(zip(["preprocessed", "trace", "interaction", "stark", "total"], phases)),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": proof_bytes,
                "ms_per_step": e2e_ms / args.steps,
                "pipeline": "host->device copy of step i+1 overlaps the proof of step i (cm31_input_prefetch: recorded at prefetch time, "
                            "released by the running proof at its STARK phase as a throttled 16-CTA copy kernel); K copies + K proofs timed",
                "step_wall_ms": e2e_step_wall, "rejected_first_measurement": e2e_rejected,
                "serial_ms_per_step": e2e_serial_ms / args.steps,
                "serial_value": aggregate_value(world, vm_steps, args.steps, e2e_serial_ms)},
        "adapter": adapter,
        "sharded": shard_info,
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="replicas", choices=["replicas", "sharded"],
                    help="N > 1: replicas = one independent segment proof per GPU (weak scaling, the default); sharded = ONE proof "
                         "sharded over the N GPUs (strong scaling; SURVEY.md §8e)")
    ap.add_argument("--arena-gib", type=float, default=24.0, help="sharded mode: peer-mapped arena per rank")
    ap.add_argument("--log-steps", type=int, default=22, help="log2 of VM steps per proof (BASELINE metric: 2^22)")
    ap.add_argument("--cpu-sample-log", type=int, default=0,
                    help="log2 VM steps of the cpu_baseline proof (default: the workload's own size, one whole proof)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-adapter", action="store_true", help="skip the adapter side measurement")
    ap.add_argument("--sync-proofs", action="store_true",
                    help="time the blocking cm31_prove_cairo_m instead of the asynchronous submission (the host tail of a proof then leaves the GPU idle)")
    ap.add_argument("--program", default="fibonacci_loop", choices=["fibonacci_loop", "array_sum", "u32_counter", "u32_mix", "sha256", "all_opcodes"],
                    help="side measurements on the other hand-assembled programs (the headline is fibonacci_loop)")
    ap.add_argument("--iterations", type=int, default=0, help="program argument n (default: 2^log_steps / 8 for fibonacci_loop)")
    ap.add_argument("--dist-selftest", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not args.cpu_sample_log:
        args.cpu_sample_log = args.log_steps
    if args.dist_selftest:
        dist_selftest(rank, world)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world == 1 and args.gpus > 1:
            print(f"bench.py: --gpus {args.gpus} needs torchrun (one process per GPU); running 1 GPU", file=sys.stderr)
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

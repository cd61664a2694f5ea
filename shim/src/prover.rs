//! `prove_cairo_m` on the GPU (crates/prover/src/prover.rs:23-147).
//!
//! cairo-m's driver names `SimdBackend` concretely (`prover.rs:5,56,63,131`; `components/mod.rs:13,110,420`; SURVEY.md §7
//! H2): besides the Backend traits it needs GPU replacements for `Claim::write_trace`, `write_interaction_trace`
//! (`LogupTraceGenerator`) and the range-check multiplicities.  Two integration levels are provided:
//!
//! 1. [`prove_cairo_m_cuda`]: the whole path on the device.  The `ProverInput` is flattened into the u32 tables of
//!    `cm31_prover_input_desc` (include/cm31.h) -- or, better, produced on the device from the runner's logs by
//!    [`import_from_runner_output_cuda`] -- and `cm31_prove_cairo_m_json` runs libcm31's protocol driver
//!    (csrc/cairo/prover.hpp, a statement-for-statement mirror of `prove_cairo_m` + stwo's `prove`) and returns the
//!    reference's serde JSON of `Proof<Blake2sMerkleHasher>` (`crates/prover/src/lib.rs:61-73`), which
//!    `verify_cairo_m` (`verifier.rs:17-95`) accepts.
//! 2. The Backend traits of this crate + [`crate::component_prover`]: stwo's own generic `prove::<CudaBackend, _>` drives
//!    the ops; the witness side is [`write_component_trace`] / [`write_component_interaction`] below.
use std::ffi::c_char;

use cairo_m_common::PublicAddressRanges;
use cairo_m_prover::adapter::io::{IoMemoryEntry, IoTraceEntry, VmImportError};
use cairo_m_prover::adapter::ProverInput;
use cairo_m_prover::errors::ProvingError;
use cairo_m_prover::Proof;
use cairo_m_runner::vm::Segment;
use itertools::Itertools;
use stwo_prover::core::backend::Column;
use stwo_prover::core::pcs::PcsConfig;
use stwo_prover::core::vcs::blake2_merkle::Blake2sMerkleHasher;

use crate::bytecode::AirProgram;
use crate::column::DeviceColumn;
use crate::ffi::*;

/// A prover input resident in HBM (`cm31_prover_input*`).
pub struct CudaProverInput(pub *mut Cm31ProverInput);
impl Drop for CudaProverInput {
    fn drop(&mut self) {
        unsafe { cm31_input_destroy(self.0) };
    }
}

/// `import_from_runner_output` (crates/prover/src/adapter/mod.rs:233-…) with the per-step work done on the device: the
/// serial `ExecutionBundleIterator` / `Memory::push` walk (`adapter/memory.rs:264-403, 470-537`) becomes a sort by address
/// plus scans (csrc/adapter.cu).  The logs are passed as the runner serialises them (`adapter/io.rs:38-60`).
pub fn import_from_runner_output_cuda(segment: &Segment, ranges: &PublicAddressRanges) -> Result<CudaProverInput, VmImportError> {
    let trace: Vec<IoTraceEntry> = segment.trace.iter().map(Into::into).collect();
    let mem: Vec<IoMemoryEntry> = segment.memory_trace.borrow().iter().map(Into::into).collect();
    let init: Vec<[u32; 4]> = segment.initial_memory.iter().map(|v| v.to_m31_array().map(|x| x.0)).collect();
    let r = [ranges.program.start, ranges.program.end, ranges.input.start, ranges.input.end, ranges.output.start, ranges.output.end];
    let mut h = std::ptr::null_mut();
    let rc = unsafe {
        cm31_adapter_import(
            bytemuck::cast_slice::<_, u32>(&trace).as_ptr(), trace.len(), bytemuck::cast_slice::<_, u32>(&mem).as_ptr(), mem.len(),
            bytemuck::cast_slice::<_, u32>(&init).as_ptr(), init.len(), r.as_ptr(), &mut h,
        )
    };
    if rc != 0 {
        // "empty trace", "invalid opcode", "unexpected end of the memory trace", "unexpected memory access": io.rs:12-36
        return Err(VmImportError::from_message(last_error()));
    }
    Ok(CudaProverInput(h))
}

/// A `ProverInput` produced by the reference's own (host) adapter, flattened for `cm31_input_create`.
pub fn upload_prover_input(input: &ProverInput) -> CudaProverInput {
    // states_by_opcodes: opcode groups over one bundle table, each group in execution order (ascending opcode id; the
    // component order of the proof is fixed by the claim, not by this table)
    let (mut opcode_ids, mut bundle_start, mut bundles) = (Vec::new(), vec![0u64], Vec::<u32>::new());
    for (opcode, states) in input.instructions.states_by_opcodes.iter().sorted_by_key(|(k, _)| **k) {
        opcode_ids.push(*opcode);
        for b in states {
            let inst = b.instruction.instruction.to_m31_vec(); // <= 6 words, padded with zeros
            let mut words = [0u32; 12];
            words[0] = b.registers.pc.0;
            words[1] = b.registers.fp.0;
            words[2] = b.clock.0;
            words[3] = b.instruction.prev_clock.0;
            for (k, w) in inst.iter().enumerate() {
                words[4 + k] = w.0;
            }
            words[10] = b.access_span.start;
            words[11] = b.access_span.len as u32;
            bundles.extend(words);
        }
        bundle_start.push((bundles.len() / 12) as u64);
    }
    let data_accesses = input.instructions.data_accesses.iter().flat_map(|a| [a.address.0, a.prev_clock.0, a.prev_value.0, a.value.0]).collect_vec();
    // boundary memory: (address, clock, value[4], multiplicity, root) rows in ascending address order -- the reference
    // iterates two std HashMaps (components/memory.rs:105-109), i.e. any order is a valid witness; ascending is ours
    let mem_rows = |m: &std::collections::HashMap<_, _>, root: u32| {
        m.iter()
            .sorted_by_key(|(addr, _)| addr.0)
            .flat_map(|(addr, (value, clock, mult))| {
                let v = value.to_m31_array();
                [addr.0, clock.0, v[0].0, v[1].0, v[2].0, v[3].0, mult.0, root]
            })
            .collect_vec()
    };
    let (initial_root, final_root) = (input.merkle_trees.initial_root.map_or(0, |r| r.0), input.merkle_trees.final_root.map_or(0, |r| r.0));
    let initial_memory = mem_rows(&input.memory.initial_memory, initial_root);
    let final_memory = mem_rows(&input.memory.final_memory, final_root);
    let clock_updates = input.memory.clock_update_data.iter().flat_map(|(addr, prev_clk, value)| {
        let v = value.to_m31_array();
        [addr.0, prev_clk.0, v[0].0, v[1].0, v[2].0, v[3].0]
    }).collect_vec();
    let node = |n: &cairo_m_prover::adapter::merkle::NodeData, root: u32| {
        let a = n.to_m31_array();
        [a[0].0, a[1].0, a[2].0, a[3].0, a[4].0, a[5].0, a[6].0, a[7].0, root]
    };
    let merkle_nodes = input.merkle_trees.initial_tree.iter().flat_map(|n| node(n, initial_root))
        .chain(input.merkle_trees.final_tree.iter().flat_map(|n| node(n, final_root))).collect_vec();
    let r = &input.public_address_ranges;
    let desc = Cm31ProverInputDesc {
        initial_pc: input.instructions.initial_registers.pc.0,
        initial_fp: input.instructions.initial_registers.fp.0,
        final_pc: input.instructions.final_registers.pc.0,
        final_fp: input.instructions.final_registers.fp.0,
        public_ranges: [r.program.start, r.program.end, r.input.start, r.input.end, r.output.start, r.output.end],
        initial_root,
        final_root,
        n_steps: (bundles.len() / 12) as u64,
        n_opcodes: opcode_ids.len() as u64,
        opcode_ids: opcode_ids.as_ptr(),
        bundle_start: bundle_start.as_ptr(),
        bundles: bundles.as_ptr(),
        data_accesses: data_accesses.as_ptr(),
        n_data_accesses: (data_accesses.len() / 4) as u64,
        initial_memory: initial_memory.as_ptr(),
        n_initial_memory: (initial_memory.len() / 8) as u64,
        final_memory: final_memory.as_ptr(),
        n_final_memory: (final_memory.len() / 8) as u64,
        clock_updates: clock_updates.as_ptr(),
        n_clock_updates: (clock_updates.len() / 6) as u64,
        merkle_nodes: merkle_nodes.as_ptr(),
        n_merkle_nodes: (merkle_nodes.len() / 9) as u64,
    };
    let mut h = std::ptr::null_mut();
    check(unsafe { cm31_input_create(&desc, &mut h) });
    check(unsafe { cm31_input_upload(h) });
    CudaProverInput(h)
}

/// `prove_cairo_m::<Blake2sMerkleChannel>` (crates/prover/src/prover.rs:23-147) on the device.
pub fn prove_cairo_m_cuda(input: &CudaProverInput, pcs_config: Option<PcsConfig>) -> Result<Proof<Blake2sMerkleHasher>, ProvingError> {
    let cfg = pcs_config.unwrap_or(cairo_m_prover::prover_config::REGULAR_96_BITS);
    assert_eq!(cfg.fri_config.log_blowup_factor, 1, "libcm31 proves with blowup 2^1 (crates/prover/src/prover_config.rs:13-20)");
    assert_eq!(cfg.fri_config.log_last_layer_degree_bound, 0);
    let mut len = 0usize;
    let rc = unsafe { cm31_prove_cairo_m_json(input.0, cfg.pow_bits, cfg.fri_config.n_queries as u32, std::ptr::null_mut(), 0, &mut len) };
    if rc != 0 {
        return Err(proving_error(last_error()));
    }
    let mut buf = vec![0u8; len + 1];
    let rc = unsafe { cm31_prove_cairo_m_json(input.0, cfg.pow_bits, cfg.fri_config.n_queries as u32, buf.as_mut_ptr() as *mut c_char, buf.len(), &mut len) };
    if rc != 0 {
        return Err(proving_error(last_error()));
    }
    buf.truncate(len);
    Ok(sonic_rs::from_slice(&buf).expect("libcm31 emits the serde layout of Proof<Blake2sMerkleHasher>"))
}

fn proving_error(msg: String) -> ProvingError {
    if msg.contains("ConstraintsNotSatisfied") {
        ProvingError::ConstraintsNotSatisfied // stwo prover/mod.rs:76-82
    } else {
        panic!("libcm31: {msg}") // contract violations panic in the reference as well
    }
}

// ------------------------------------------------------------------ level 2: witness generation per component
/// `Claim::write_trace` of one opcode component (e.g. crates/prover/src/components/opcodes/store_fp_fp.rs:153-309):
/// `Pack::pack` of the bundles (utils/execution_bundle.rs:31-75) = `cm31_unpack_bundles_slots`, then the component's
/// trace-fill program.  `bundles` / `accesses` are device tables (12 / 4 words per record).
pub fn write_component_trace(program: &AirProgram, n_trace_columns: usize, n_access_slots: u32, bundles: &DeviceColumn, n_real: usize,
                             accesses: &DeviceColumn, log_size: u32) -> Vec<DeviceColumn> {
    let n_inputs = 10 + 4 * n_access_slots as usize;
    let mut inputs = (0..n_inputs).map(|_| unsafe { DeviceColumn::uninitialized(1 << log_size) }).collect_vec();
    let in_ptrs = inputs.iter_mut().map(|c| c.as_mut_ptr()).collect_vec();
    check(unsafe { cm31_unpack_bundles_slots(bundles.as_ptr(), n_real, log_size, accesses.as_ptr(), accesses.len() / 4, in_ptrs.as_ptr(), n_access_slots) });
    let mut outs = (0..n_trace_columns).map(|_| unsafe { DeviceColumn::uninitialized(1 << log_size) }).collect_vec();
    let out_ptrs = outs.iter_mut().map(|c| c.as_mut_ptr()).collect_vec();
    let in_const = inputs.iter().map(|c| c.as_ptr()).collect_vec();
    check(unsafe {
        cm31_air_program(in_const.as_ptr(), in_const.len(), out_ptrs.as_ptr(), out_ptrs.len(), log_size, program.code.as_ptr(), program.code.len(),
                         program.n_regs, program.consts.as_ptr(), program.consts.len())
    });
    outs
}

/// `write_interaction_trace` + `LogupTraceGenerator::{write_frac, finalize_col, finalize_last}`
/// (external/stwo/crates/constraint_framework/src/logup.rs:123-320): the logup program stores the cumulative column of
/// every batch; the last one is shifted by claimed_sum / n and prefix-summed in coset order.  Returns the claimed sum.
pub fn write_component_interaction(program: &AirProgram, trace: &[DeviceColumn], n_interaction_columns: usize, log_size: u32) -> (Vec<DeviceColumn>, [u32; 4]) {
    let mut outs = (0..n_interaction_columns).map(|_| unsafe { DeviceColumn::uninitialized(1 << log_size) }).collect_vec();
    let out_ptrs = outs.iter_mut().map(|c| c.as_mut_ptr()).collect_vec();
    let in_ptrs = trace.iter().map(|c| c.as_ptr()).collect_vec();
    check(unsafe {
        cm31_air_program(in_ptrs.as_ptr(), in_ptrs.len(), out_ptrs.as_ptr(), out_ptrs.len(), log_size, program.code.as_ptr(), program.code.len(),
                         program.n_regs, program.consts.as_ptr(), program.consts.len())
    });
    let last4: [*mut u32; 4] = std::array::from_fn(|k| out_ptrs[n_interaction_columns - 4 + k]);
    let mut claimed = [0u32; 4];
    check(unsafe { cm31_logup_finalize_last(last4.as_ptr(), log_size, claimed.as_mut_ptr()) });
    (outs, claimed)
}

/// Range-check / bitwise multiplicities (crates/prover/src/preprocessed/range_check/range_check_macro.rs:72-84): the
/// component's lookup program counts every looked-up value into `bins`; a value outside the table raises the device error
/// word instead of writing out of bounds (`check_lookups` after all components).
pub fn emit_component_lookups(program: &AirProgram, trace: &[DeviceColumn], bins: &mut DeviceColumn, log_size: u32) {
    let in_ptrs = trace.iter().map(|c| c.as_ptr()).collect_vec();
    let log_bins = bins.len().ilog2();
    check(unsafe {
        cm31_air_lookups(in_ptrs.as_ptr(), in_ptrs.len(), bins.as_mut_ptr(), log_bins, log_size, program.code.as_ptr(), program.code.len(),
                         program.n_regs, program.consts.as_ptr(), program.consts.len())
    });
}
pub fn check_lookups() {
    check(unsafe { cm31_air_error_check() });
}

/// One proof per continuation segment (`crates/prover/tests/prover.rs:204-243`), pipelined: while segment i is being proven the
/// input of segment i+1 travels to the device (`cm31_input_prefetch`: the running proof releases the copy as a throttled copy
/// kernel at its STARK phase), and the host-side tail of proof i -- query generation, decommitment planning, the gather of the
/// queried values, assembly and serialisation -- runs under the trace / interaction commitments of proof i+1
/// (`cm31_prove_cairo_m_async`).  The proofs are byte-identical to `prove_cairo_m_cuda`'s.
pub fn prove_segments_cuda(inputs: &[CudaProverInput], pcs_config: Option<PcsConfig>) -> Result<Vec<Proof<Blake2sMerkleHasher>>, ProvingError> {
    let cfg = pcs_config.unwrap_or_else(PcsConfig::default);
    const CAP: usize = 1 << 26;
    let mut bufs = [vec![0u8; CAP], vec![0u8; CAP]];
    let mut lens = [0usize; 2];
    let mut proofs = Vec::with_capacity(inputs.len());
    let mut collect = |buf: &[u8], len: usize, proofs: &mut Vec<Proof<Blake2sMerkleHasher>>| -> Result<(), ProvingError> {
        // blob -> the reference's serde JSON -> Proof (cm31_proof_to_json, crates/prover/src/lib.rs:61-73)
        let mut n = 0usize;
        check(unsafe { cm31_proof_to_json(buf.as_ptr(), len, std::ptr::null_mut(), 0, &mut n) });
        let mut js = vec![0u8; n + 1];
        check(unsafe { cm31_proof_to_json(buf.as_ptr(), len, js.as_mut_ptr() as *mut c_char, js.len(), &mut n) });
        proofs.push(sonic_rs::from_slice(&js[..n]).map_err(|_| ProvingError::ConstraintsNotSatisfied)?);
        Ok(())
    };
    if let Some(first) = inputs.first() {
        check(unsafe { cm31_input_prefetch(first.0) });
    }
    for (i, input) in inputs.iter().enumerate() {
        if let Some(next) = inputs.get(i + 1) {
            check(unsafe { cm31_input_prefetch(next.0) });
        }
        // submitting proof i completes proof i-1 at the latest: its bytes are in the other buffer
        let rc = unsafe {
            cm31_prove_cairo_m_async(input.0, cfg.pow_bits, cfg.fri_config.n_queries as u32, bufs[i & 1].as_mut_ptr(), CAP, &mut lens[i & 1], std::ptr::null_mut())
        };
        if rc != 0 {
            return Err(ProvingError::ConstraintsNotSatisfied);
        }
        if i > 0 {
            // (a zero return of call i also vouches for proof i-1, which was completed inside it)
            collect(&bufs[(i - 1) & 1], lens[(i - 1) & 1], &mut proofs)?;
        }
    }
    if !inputs.is_empty() {
        check(unsafe { cm31_prove_wait() });
        let i = inputs.len() - 1;
        collect(&bufs[i & 1], lens[i & 1], &mut proofs)?;
    }
    Ok(proofs)
}

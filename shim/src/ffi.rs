//! `extern "C"` declarations of include/cm31.h (one per exported entry point the shim uses) and the status -> panic
//! conversion.  The reference's backend ops `assert!`/`panic!` on contract violations, so a non-zero status
//! becomes a panic carrying `cm31_last_error()`; only `prove` returns `Err(ConstraintsNotSatisfied)`, from the generic code.
use std::ffi::{c_char, c_int, c_void, CStr};

#[repr(C)]
pub struct Cm31Twiddles {
    _private: [u8; 0],
}
#[repr(C)]
pub struct Cm31ProverInput {
    _private: [u8; 0],
}
#[repr(C)]
pub struct Cm31AdapterLogs {
    _private: [u8; 0],
}

/// include/cm31.h: `cm31_prover_input_desc` -- the reference's `ProverInput` as flat u32 tables.
#[repr(C)]
pub struct Cm31ProverInputDesc {
    pub initial_pc: u32,
    pub initial_fp: u32,
    pub final_pc: u32,
    pub final_fp: u32,
    pub public_ranges: [u32; 6],
    pub initial_root: u32,
    pub final_root: u32,
    pub n_steps: u64,
    pub n_opcodes: u64,
    pub opcode_ids: *const u32,
    pub bundle_start: *const u64,
    pub bundles: *const u32,
    pub data_accesses: *const u32,
    pub n_data_accesses: u64,
    pub initial_memory: *const u32,
    pub n_initial_memory: u64,
    pub final_memory: *const u32,
    pub n_final_memory: u64,
    pub clock_updates: *const u32,
    pub n_clock_updates: u64,
    pub merkle_nodes: *const u32,
    pub n_merkle_nodes: u64,
}

#[link(name = "cm31")]
extern "C" {
    pub fn cm31_last_error() -> *const c_char;
    pub fn cm31_set_device(device: c_int) -> c_int;
    pub fn cm31_sync() -> c_int;
    // ---- buffers
    pub fn cm31_malloc(out: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn cm31_free(p: *mut c_void) -> c_int;
    pub fn cm31_memset0(p: *mut c_void, bytes: usize) -> c_int;
    pub fn cm31_h2d(dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn cm31_d2h(dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn cm31_d2d(dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn cm31_gather_words(srcs: *const *const u32, n_srcs: usize, src_id: *const u32, word: *const u32, n: usize, out: *mut u32) -> c_int;
    pub fn cm31_lane(lane: c_int) -> c_int;
    pub fn cm31_lanes_join() -> c_int;
    // ---- poly
    pub fn cm31_twiddles_create(log_size: u32, out: *mut *mut Cm31Twiddles) -> c_int;
    pub fn cm31_twiddles_destroy(tw: *mut Cm31Twiddles) -> c_int;
    pub fn cm31_interpolate_batch(cols: *const *mut u32, n_cols: usize, log_size: u32, tw: *const Cm31Twiddles) -> c_int;
    pub fn cm31_evaluate_batch(coeffs: *const *const u32, out: *const *mut u32, n_cols: usize, log_size: u32, log_eval_size: u32, tw: *const Cm31Twiddles) -> c_int;
    pub fn cm31_eval_at_point_batch(coeffs: *const *const u32, log_sizes: *const u32, n_polys: usize, points: *const u32, n_points: usize, point_idx: *const u32, out: *mut u32) -> c_int;
    pub fn cm31_bit_reverse(col: *mut u32, log_size: u32) -> c_int;
    // ---- vcs / pow
    pub fn cm31_blake2s_commit_layer(log_size: u32, prev: *const u32, cols: *const *const u32, n_cols: usize, out: *mut u32) -> c_int;
    pub fn cm31_blake2s_commit_multi(log_size: u32, prev: *const u32, cols: *const *const u32, n_cols: usize, n_levels: u32, out_layers: *const *mut u32) -> c_int;
    pub fn cm31_blake2s_commit_top(top_log: u32, prev: *const u32, cols: *const *const u32, col_start: *const u32, out_layers: *const *mut u32) -> c_int;
    pub fn cm31_grind_blake2s(digest: *const u8, pow_bits: u32, nonce_out: *mut u64) -> c_int;
    // ---- air / pcs / fri
    pub fn cm31_accumulate(dst4: *const *mut u32, src4: *const *const u32, n: usize) -> c_int;
    pub fn cm31_secure_powers(felt: *const u32, n: usize, out: *mut u32) -> c_int;
    pub fn cm31_accumulate_quotients(log_size: u32, cols: *const *const u32, n_cols: usize, random_coeff: *const u32, n_batches: usize, points: *const u32, batch_start: *const u32, col_idx: *const u32, values: *const u32, out4: *const *mut u32) -> c_int;
    pub fn cm31_fold_line(src4: *const *const u32, log_size: u32, alpha: *const u32, tw: *const Cm31Twiddles, dst4: *const *mut u32) -> c_int;
    pub fn cm31_fold_circle_into_line(dst4: *const *mut u32, src4: *const *const u32, log_size: u32, alpha: *const u32, tw: *const Cm31Twiddles) -> c_int;
    pub fn cm31_decompose(src4: *const *const u32, log_size: u32, dst4: *const *mut u32, lambda_out: *mut u32) -> c_int;
    pub fn cm31_constraint_eval(cols: *const *const u32, n_cols: usize, trace_log: u32, eval_log: u32, code: *const u64, n_instr: usize, n_regs: u32, consts: *const u32, n_consts: usize, denom_inv: *const u32, acc4: *const *mut u32) -> c_int;
    // ---- witness generation
    pub fn cm31_air_program(in_cols: *const *const u32, n_in: usize, out_cols: *const *mut u32, n_out: usize, log_size: u32, code: *const u64, n_instr: usize, n_regs: u32, consts: *const u32, n_consts: usize) -> c_int;
    pub fn cm31_air_lookups(in_cols: *const *const u32, n_in: usize, bins: *mut u32, log_bins: u32, log_size: u32, code: *const u64, n_instr: usize, n_regs: u32, consts: *const u32, n_consts: usize) -> c_int;
    pub fn cm31_air_error_check() -> c_int;
    pub fn cm31_air_shapes(buf: *mut c_char, cap: usize, len: *mut usize) -> c_int;
    pub fn cm31_logup_finalize_last(last4: *const *mut u32, log_size: u32, claimed_sum: *mut u32) -> c_int;
    pub fn cm31_unpack_bundles_slots(bundles: *const u32, n_real: usize, log_size: u32, accesses: *const u32, n_accesses: usize, out_cols: *const *mut u32, n_access_slots: u32) -> c_int;
    pub fn cm31_unpack_rows(rows: *const u32, n_real: usize, n_fields: u32, log_size: u32, out_cols: *const *mut u32) -> c_int;
    pub fn cm31_iota(col: *mut u32, n: usize) -> c_int;
    pub fn cm31_bitwise_table_col(k: c_int, col: *mut u32) -> c_int;
    // ---- whole proofs / adapter
    pub fn cm31_input_create(desc: *const Cm31ProverInputDesc, out: *mut *mut Cm31ProverInput) -> c_int;
    pub fn cm31_input_upload(h: *mut Cm31ProverInput) -> c_int;
    pub fn cm31_input_prefetch(h: *mut Cm31ProverInput) -> c_int;
    pub fn cm31_input_destroy(h: *mut Cm31ProverInput) -> c_int;
    pub fn cm31_adapter_import(trace: *const u32, n_trace: usize, memory_trace: *const u32, n_mem: usize, initial_memory: *const u32, n_initial: usize, public_ranges: *const u32, out: *mut *mut Cm31ProverInput) -> c_int;
    pub fn cm31_adapter_prefetch(trace: *const u32, n_trace: usize, memory_trace: *const u32, n_mem: usize, initial_memory: *const u32, n_initial: usize, public_ranges: *const u32, out: *mut *mut Cm31AdapterLogs) -> c_int;
    pub fn cm31_adapter_import_prefetched(logs: *mut Cm31AdapterLogs, out: *mut *mut Cm31ProverInput) -> c_int;
    pub fn cm31_prove_cairo_m(h: *const Cm31ProverInput, pow_bits: u32, n_queries: u32, proof_out: *mut u8, cap: usize, proof_len: *mut usize, timings_ms: *mut f64) -> c_int;
    // a stream of segments: the tail of proof i runs under proof i+1
    pub fn cm31_prove_cairo_m_async(h: *const Cm31ProverInput, pow_bits: u32, n_queries: u32, proof_out: *mut u8, cap: usize, proof_len: *mut usize, timings_ms: *mut f64) -> c_int;
    pub fn cm31_prove_wait() -> c_int;
    pub fn cm31_proof_to_json(proof: *const u8, proof_len: usize, json_out: *mut c_char, cap: usize, json_len: *mut usize) -> c_int;
    pub fn cm31_prove_cairo_m_json(h: *const Cm31ProverInput, pow_bits: u32, n_queries: u32, json_out: *mut c_char, cap: usize, json_len: *mut usize) -> c_int;
}

pub fn last_error() -> String {
    unsafe { CStr::from_ptr(cm31_last_error()) }.to_string_lossy().into_owned()
}

/// Contract violations panic, like the reference's `assert!`s in the backend ops.
#[track_caller]
pub fn check(rc: c_int) {
    if rc != 0 {
        panic!("libcm31: {}", last_error());
    }
}

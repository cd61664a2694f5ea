/*
 * cm31.h — C ABI of the B200-native cairo-m proving hot path (libcm31.so).
 *
 * This is the boundary a Rust `CudaBackend` shim binds with `extern "C"` to implement the Stwo
 * backend traits in place of `SimdBackend` (INTEGRATION.md shows the shim).  Every entry point
 * names the reference interface it replaces (paths relative to the reference checkout:
 * S/ = external/stwo/crates, P/ = crates/prover).
 *
 * Conventions
 *  - All field data are u32 words holding CANONICAL M31 values in [0, 2^31-1).
 *  - A "column" is a contiguous device array of 2^log_size u32 (the reference `BaseColumn`,
 *    S/prover/src/core/backend/simd/column.rs:26-30).  Secure (QM31) columns are 4 base columns
 *    (S/prover/src/core/secure_column.rs:11-13).
 *  - `const uint32_t* const* cols` arguments are HOST arrays of DEVICE pointers.
 *  - QM31 scalars cross the boundary as uint32_t[4] = (a, b, c, d) of (a+bi)+(c+di)u.
 *  - Hash columns are device arrays of 8 u32 (32 bytes) per node (reference `Blake2sHash`,
 *    S/prover/src/core/vcs/blake2_hash.rs:8-10).
 *  - Every function returns 0 on success; non-zero = error, text via cm31_last_error().  Contract
 *    violations the reference `assert!`s on are reported as errors, never silently ignored.
 *  - Calls are issued from one host thread and are ordered on one CUDA stream (cm31_set_stream);
 *    functions that return host-readable values synchronise that stream before returning.
 */
#ifndef CM31_H
#define CM31_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ runtime / buffers */
const char* cm31_last_error(void);
int cm31_device_count(int* out);
/* One process drives ONE device from ONE host thread (the reference calls its backend ops sequentially from one thread,
 * SURVEY.md §8b): the library's streams, staging rings and arenas are process-wide.  The device is bound by the first call
 * that uses it; cm31_set_device with another ordinal afterwards returns an error. */
int cm31_set_device(int ordinal);
int cm31_set_stream(void* cuda_stream); /* cudaStream_t; NULL = legacy default stream */
int cm31_sync(void);
/* Lanes (SimdBackend's rayon pool has no analogue of this: it is the CUDA backend's way of overlapping the many
 * latency-bound launches of small components with the kernels of large ones).  cm31_lane(1) routes the following
 * calls to a side stream ordered after everything issued so far on the caller's stream; cm31_lane(0) returns to the
 * caller's stream; cm31_lanes_join() makes the caller's stream wait for the side stream.  Between a fork and the join,
 * work on one lane must not consume results produced on the other, and a buffer must be freed on the lane that last
 * used it.  Calls that return host-readable values must be made on lane 0 after a join. */
int cm31_lane(int lane);
int cm31_lanes_join(void);
/* NVTX range on the profiler timeline (a no-op without a profiler): the host driver marks the proof's phases with the span
 * names of the reference's `tracing` spans ("Interpolation for commitment", "Extension", "Merkle", "Composition", ..) */
int cm31_range_push(const char* name);
int cm31_range_pop(void);
/* Column<T>::zeros / uninitialized / to_cpu / from_iter  (S/prover/src/core/backend/mod.rs:46-65) */
int cm31_malloc(void** out, size_t bytes);
/* Makes the stream-ordered pool hold `factor` x its high-water mark of bytes in use (one allocation + free with the GPU idle):
 * growing the pool while kernels run stalls the caller for 30-240 ms on some systems.  cm31_prove_cairo_m[_async] calls it
 * after the first proofs of a process (factor 2.5); best effort, no error when memory is short. */
int cm31_pool_reserve_headroom(double factor);
int cm31_free(void* dptr);
int cm31_memset0(void* dptr, size_t bytes);
int cm31_h2d(void* dst, const void* src_host, size_t bytes);
int cm31_d2h(void* dst_host, const void* src, size_t bytes);
/* staging copies on a background stream (overlap with the input-independent start of a proof);
 * cm31_bg_fence() orders the main stream after every background copy issued so far */
int cm31_h2d_bg(void* dst, const void* src_host, size_t bytes);
int cm31_bg_fence(void);
/* finer than the fence: *mark_out names the background copies issued so far; cm31_bg_wait(mark) orders the current
 * stream (lane) after them */
int cm31_bg_begin(void); /* order the background stream after the current one, once, then: */
int cm31_h2d_bg_ordered(void* dst, const void* src_host, size_t bytes);
int cm31_bg_mark(uint32_t* mark_out);
/* Deferred background copies: between cm31_bg_defer(1, &ticket) and cm31_bg_defer(0, NULL) the calls above are only recorded;
 * cm31_bg_release(upto) issues the recorded ones with ticket <= upto (0 = all), ordered after the current point of the main
 * stream.  cm31_input_prefetch uses it so that the prover itself places the bulk DMA of the next segment's input where the
 * running proof is not launch-bound (CM31_PREFETCH_AT). */
int cm31_bg_defer(int on, uint64_t* ticket_out);
int cm31_bg_release(uint64_t upto);
/* the same with the copies >= 1 MB issued as a `throttle_ctas`-CTA copy kernel over the unified address space (bounded PCIe
 * reads in flight) when the source is page-locked; 0 = copy engine */
int cm31_bg_release_throttled(uint64_t upto, int throttle_ctas);
int cm31_bg_cancel(uint64_t ticket); /* forget the recorded copies of a ticket (their destinations are being freed) */
int cm31_bg_wait(uint32_t mark);
int cm31_d2d(void* dst, const void* src, size_t bytes);
/* Column::at for many (column,row) pairs at once (decommit):
 * out_host[c * n_idx + q] = cols[c][idx_host[q]]   (S/prover/src/core/vcs/prover.rs:125-140) */
int cm31_gather_u32(const uint32_t* const* cols, size_t n_cols, const uint32_t* idx_host, size_t n_idx,
                    uint32_t* out_host);
/* every read of a whole proof's decommitment in one launch (4 trees + all FRI layers):
 * out_host[k] = srcs[src_id_host[k]][word_idx_host[k]]; a hash node is 8 consecutive words */
int cm31_gather_words(const uint32_t* const* srcs, size_t n_srcs, const uint32_t* src_id_host, const uint32_t* word_idx_host,
                      size_t n, uint32_t* out_host);
/* run form: request k copies out_off_host[k+1]-out_off_host[k] consecutive words starting at
 * srcs[src_id_host[k]][word_idx_host[k]] to out_host[out_off_host[k]..] (1 word = a column element,
 * 8 words = a hash node); out_off_host has n+1 entries */
int cm31_gather_runs(const uint32_t* const* srcs, size_t n_srcs, const uint32_t* src_id_host, const uint32_t* word_idx_host,
                     const uint32_t* out_off_host, size_t n, uint32_t* out_host);
/* One decommitment gather per proof (vcs/prover.rs:125-140 and fri.rs:1002-1036 read element by element):
 * n_runs run requests (cnt_host[k] words from srcs[src_id_host[k]][word_idx_host[k]] to out_host[out_off_host[k]..]) plus
 * n_grids row-grid requests, 5 words each in grid_desc_host = (col_off, n_cols, row_off, n_rows, out_base):
 * out_host[out_base + k*n_cols + c] = srcs[grid_cols_host[col_off + c]][grid_rows_host[row_off + k]] -- every column of a
 * Merkle layer at every visited node.  total_words = size of out_host. */
int cm31_gather_batch(const uint32_t* const* srcs, size_t n_srcs, const uint32_t* src_id_host, const uint32_t* word_idx_host,
                      const uint32_t* out_off_host, const uint32_t* cnt_host, size_t n_runs, const uint32_t* grid_desc_host,
                      size_t n_grids, const uint32_t* grid_cols_host, size_t n_grid_cols, const uint32_t* grid_rows_host,
                      size_t n_grid_rows, size_t total_words, uint32_t* out_host);
/* The same without the final synchronisation: *result_out = the page-locked buffer the total_words words land in (valid until
 * the second next call), cm31_gather_wait() returns once they have.  Lets a prover assemble proof i while proof i+1 runs. */
int cm31_gather_batch_async(const uint32_t* const* srcs, size_t n_srcs, const uint32_t* src_id_host, const uint32_t* word_idx_host,
                            const uint32_t* out_off_host, const uint32_t* cnt_host, size_t n_runs, const uint32_t* grid_desc_host,
                            size_t n_grids, const uint32_t* grid_cols_host, size_t n_grid_cols, const uint32_t* grid_rows_host,
                            size_t n_grid_rows, size_t total_words, const uint32_t** result_out);
int cm31_gather_wait(void);
/* same for hash columns: out_host[q*8..] = layer[idx[q]] */
int cm31_gather_hash(const uint32_t* layer, const uint32_t* idx_host, size_t n_idx, uint32_t* out_host);

/* Peer memory for the multi-GPU commitment (one process per GPU): a buffer allocated with
 * cm31_ipc_alloc can be opened by the other ranks (cm31_ipc_open on the 64-byte handle) and passed to
 * any kernel as a column pointer — the Merkle leaf kernel then reads remote LDE columns over
 * NVLink/NVSwitch while it hashes, instead of waiting for an all-to-all. */
int cm31_ipc_alloc(size_t bytes, void** out, uint8_t handle_out[64]);
int cm31_ipc_free(void* dptr);
int cm31_ipc_open(const uint8_t handle[64], void** out);
int cm31_ipc_close(void* dptr);

/* ------------------------------------------------------------------ single-proof sharding
 * One process per GPU, every rank runs the same protocol driver and owns a subset of the components; see csrc/shard.cu.
 * While sharding is on, cm31_malloc bump-allocates from a per-rank arena that every peer has mapped (cudaIpc): the same
 * allocation sequence on every rank puts a buffer at the same offset everywhere, so cm31_shard_peer translates any arena
 * pointer into rank `owner`'s copy and kernels read remote columns over NVLink/NVSwitch directly.  Collectives are NCCL
 * (all stream-ordered on the current lane): barrier, in-place all-gather(s), u32 sum.  cm31_shard_reduce_m31 is the mod-P
 * all-reduce of the four composition accumulator columns (S/prover/src/core/air/accumulation.rs:49-58) as reduce-scatter
 * by peer loads + all-gather.  Rank 0 creates the 128-byte NCCL id (cm31_shard_unique_id) and hands it to the others
 * (torch.distributed broadcast in bench.py). */
int cm31_shard_unique_id(uint8_t out[128]);
int cm31_shard_init(int rank, int world, const uint8_t unique_id[128], size_t arena_bytes);
int cm31_shard_finalize(void);
/* world = 0 when sharding is off; stats: arena high-water mark, bytes received by all-gathers, bytes all-reduced, collectives */
int cm31_shard_info(int* rank, int* world, uint64_t stats[4]);
int cm31_shard_barrier(void);
int cm31_shard_arena_reset(void);   /* new proof: barrier, arena back to offset 0, cm31_malloc draws from it */
int cm31_shard_arena_suspend(void); /* end of proof: cm31_malloc goes back to the ordinary pool (the arena keeps its contents) */
int cm31_shard_allgather(void* buf, size_t bytes_per_rank);
int cm31_shard_allgather_many(void* const* bufs, const size_t* bytes_per_rank, size_t n);
int cm31_shard_allreduce_u32(uint32_t* buf, size_t n);
int cm31_shard_allreduce_host_u32(uint32_t* host, size_t n);
int cm31_shard_peer(const void* p, int owner, const void** out);
int cm31_shard_reduce_m31(uint32_t* const dst4[4], const uint32_t* const src4[4], size_t n);
/* the component -> rank assignment every rank computes for itself (longest-processing-time first over per-component costs,
 * claim order); host only */
int cm31_shard_plan(const double* cost, size_t n_components, int world, int* owners_out);
/* MerkleOps::commit_on_layer restricted to nodes [first_node, first_node + n_nodes) of the layer (a rank's row range;
 * prev_layer / out_layer are full-size buffers of which only that range is read / written) */
int cm31_blake2s_commit_layer_range(uint32_t log_size, const uint32_t* prev_layer, const uint32_t* const* cols, size_t n_cols,
                                    uint32_t* out_layer, size_t first_node, size_t n_nodes);
/* the fused form (cm31_blake2s_commit_multi) on an aligned power-of-two node range of at least 256 nodes */
int cm31_blake2s_commit_multi_range(uint32_t log_size, const uint32_t* prev_layer, const uint32_t* const* cols, size_t n_cols,
                                    uint32_t n_levels, uint32_t* const* out_layers, size_t first_node, size_t n_nodes);
/* QuotientOps::accumulate_quotients restricted to rows [first_row, first_row + n_rows) (multiples of 256) */
int cm31_accumulate_quotients_range(uint32_t log_size, const uint32_t* const* cols, size_t n_cols, const uint32_t random_coeff[4],
                                    size_t n_batches, const uint32_t* points, const uint32_t* batch_start, const uint32_t* col_idx,
                                    const uint32_t* values, uint32_t* const out4[4], size_t first_row, size_t n_rows);

/* The same with a per-entry mask (entry_active[k] != 0: the k-th (batch, column) entry contributes; NULL = all).  The quotient
 * is linear in the per-column terms, so a sharded proof splits the ENTRIES over the ranks -- every rank accumulates the
 * terms of the columns it owns over all rows from local memory -- and sums the partial quotients mod P
 * (cm31_shard_reduce_m31).  Inactive entries still take their power of the random coefficient. */
int cm31_accumulate_quotients_partial(uint32_t log_size, const uint32_t* const* cols, size_t n_cols, const uint32_t random_coeff[4],
                                      size_t n_batches, const uint32_t* points, const uint32_t* batch_start, const uint32_t* col_idx,
                                      const uint32_t* values, uint32_t* const out4[4], size_t first_row, size_t n_rows,
                                      const uint8_t* entry_active);

/* ------------------------------------------------------------------ PolyOps
 * S/prover/src/core/poly/circle/ops.rs:13-69, CPU definition S/prover/src/core/backend/cpu/circle.rs */
typedef struct cm31_twiddles cm31_twiddles;
/* PolyOps::precompute_twiddles(CanonicCoset(log_size).half_coset())  (cpu/circle.rs:137-188):
 * twiddle tree able to serve every canonic circle domain of log size <= log_size. */
int cm31_twiddles_create(uint32_t log_size, cm31_twiddles** out);
int cm31_twiddles_destroy(cm31_twiddles* tw);
/* device pointers to the (x) and (1/x) tree buffers, 2^(log_size-1) words each, stwo CPU layout */
int cm31_twiddles_buffers(const cm31_twiddles* tw, const uint32_t** twiddles, const uint32_t** itwiddles,
                          uint32_t* log_size);
/* PolyOps::interpolate / interpolate_columns (ops.rs:19-33, cpu/circle.rs:18-71): in place,
 * evaluations on CanonicCoset(log_size).circle_domain() in bit-reversed order -> FFT-basis coeffs. */
int cm31_interpolate_batch(uint32_t* const* cols, size_t n_cols, uint32_t log_size, const cm31_twiddles* tw);
/* out-of-place form: the evaluations stay valid (cairo-m needs the trace values again for the logup
 * columns, so the prover keeps them instead of copying every column before interpolating) */
int cm31_interpolate_batch_to(const uint32_t* const* evals, uint32_t* const* coeffs_out, size_t n_cols, uint32_t log_size,
                              const cm31_twiddles* tw);
/* PolyOps::evaluate / evaluate_polynomials (ops.rs:43-65, cpu/circle.rs:97-135): coefficient
 * vectors of 2^log_size words -> evaluations on CanonicCoset(log_eval_size).circle_domain(),
 * bit-reversed order, out columns of 2^log_eval_size words (log_eval_size >= log_size). */
int cm31_evaluate_batch(const uint32_t* const* coeffs, uint32_t* const* out, size_t n_cols, uint32_t log_size,
                        uint32_t log_eval_size, const cm31_twiddles* tw);
/* PolyOps::eval_at_point (ops.rs:36, cpu/circle.rs:73-87) for many (poly, point) pairs:
 * points_host[k*8..] = (x.a,x.b,x.c,x.d,y.a,..,y.d); point_idx_host[i] selects the point of poly i;
 * out_host[i*4..] = poly_i(point).  Synchronises. */
int cm31_eval_at_point_batch(const uint32_t* const* coeffs, const uint32_t* log_sizes_host, size_t n_polys,
                             const uint32_t* points_host, size_t n_points, const uint32_t* point_idx_host,
                             uint32_t* out_host);
/* ColumnOps::bit_reverse_column (S/prover/src/core/backend/cpu/mod.rs:36-46), in place */
int cm31_bit_reverse(uint32_t* col, uint32_t log_size);

/* ------------------------------------------------------------------ MerkleOps<Blake2sMerkleHasher>
 * S/prover/src/core/vcs/ops.rs:41-45, CPU def cpu/blake2s.rs:9-23, hasher vcs/blake2_merkle.rs:14-30:
 * out[i] = Blake2s( prev[2i] || prev[2i+1] || le32(cols[0][i]) || ... ), i < 2^log_size.
 * prev_layer may be NULL (leaf layer). */
int cm31_blake2s_commit_layer(uint32_t log_size, const uint32_t* prev_layer, const uint32_t* const* cols,
                              size_t n_cols, uint32_t* out_layer);

/* Layers top_log_size .. 0 of MerkleProver::commit (vcs/prover.rs:52-64) in one launch (top_log_size <= 10):
 * layer l hashes prev = layer l+1 (prev_layer for l = top_log_size, may be NULL) and the columns
 * cols[col_start_host[l] .. col_start_host[l+1]) of 2^l words; out_layers[l] receives 2^l nodes. */
/* n_levels (1..9) consecutive layers log_size, log_size-1, .. in one launch; only the first may carry columns (and a
 * previous layer); out_layers[l] receives layer log_size - l.  Same hashes as n_levels calls of commit_layer. */
int cm31_blake2s_commit_multi(uint32_t log_size, const uint32_t* prev_layer, const uint32_t* const* cols, size_t n_cols,
                              uint32_t n_levels, uint32_t* const* out_layers);
int cm31_blake2s_commit_top(uint32_t top_log_size, const uint32_t* prev_layer, const uint32_t* const* cols,
                            const uint32_t* col_start_host, uint32_t* const* out_layers);

/* ------------------------------------------------------------------ AccumulationOps
 * S/prover/src/core/air/accumulation.rs:156-162 */
int cm31_accumulate(uint32_t* const dst4[4], const uint32_t* const src4[4], size_t n);
int cm31_secure_powers(const uint32_t felt[4], size_t n_powers, uint32_t* out_host /* 4*n */);

/* ------------------------------------------------------------------ QuotientOps
 * S/prover/src/core/pcs/quotients.rs:22-35, CPU def cpu/quotients.rs:18-146.
 * One call per distinct column log size.  Sample batches in flattened form:
 *   batch b: point = batch_points_host[b*8..], columns col_idx_host[batch_start_host[b] ..
 *   batch_start_host[b+1]) with sampled values values_host[k*4..].
 * out4 = 4 coordinate columns of 2^log_size words. */
int cm31_accumulate_quotients(uint32_t log_size, const uint32_t* const* cols, size_t n_cols,
                              const uint32_t random_coeff[4], size_t n_batches, const uint32_t* batch_points_host,
                              const uint32_t* batch_start_host, const uint32_t* col_idx_host,
                              const uint32_t* values_host, uint32_t* const out4[4]);

/* ------------------------------------------------------------------ FriOps
 * S/prover/src/core/fri.rs:92-139; definitions fri.rs:1132-1189, cpu/fri.rs:29-85 */
/* LineEvaluation of 2^log_size values on LineDomain(half_odds(log_size)) -> 2^(log_size-1) */
int cm31_fold_line(const uint32_t* const src4[4], uint32_t log_size, const uint32_t alpha[4],
                   const cm31_twiddles* tw, uint32_t* const dst4[4]);
/* SecureEvaluation of 2^log_size on CanonicCoset(log_size).circle_domain():
 * dst[i] = dst[i]*alpha^2 + fold(src)[i], dst has 2^(log_size-1) values */
int cm31_fold_circle_into_line(uint32_t* const dst4[4], const uint32_t* const src4[4], uint32_t log_size,
                               const uint32_t alpha[4], const cm31_twiddles* tw);
/* The small inner layers of FriProver::commit_inner_layers (S/prover/src/core/fri.rs:226-266) in ONE launch, Fiat-Shamir
 * included: per layer the Blake2s Merkle tree of the 4 coordinate columns, root -> mix_root, draw_secure_felt -> alpha
 * (S/prover/src/core/channel/blake2s.rs:60-116), fold_line and -- where a column joins at that size -- fold_circle_into_line
 * with the same alpha.  digest_in = the channel digest before the first mix_root; roots_out_host receives the n_layers roots
 * (8 words each): the caller replays mix_root / draw_secure_felt on its own channel, which keeps both transcripts identical.
 * layers[i].log_size halves from layer to layer; tree_levels[j] = hash layer with 2^j nodes (j = 0 .. log_size). */
typedef struct cm31_fri_tail_layer {
    const uint32_t* ev_in[4];
    uint32_t* ev_out[4];
    uint32_t* const* tree_levels;
    const uint32_t* circle[4]; /* all NULL when no column joins after this fold */
    uint32_t log_size;
} cm31_fri_tail_layer;
int cm31_fri_tail(const uint32_t digest_in[8], const cm31_fri_tail_layer* layers, size_t n_layers, const cm31_twiddles* tw,
                  uint32_t* roots_out_host);
int cm31_decompose(const uint32_t* const src4[4], uint32_t log_size, uint32_t* const dst4[4],
                   uint32_t lambda_out[4]);

/* ------------------------------------------------------------------ GrindOps<Blake2sChannel>
 * S/prover/src/core/proof_of_work.rs:3-7, cpu/grind.rs:5-16: the MINIMUM nonce such that
 * blake2s(digest || le32(nonce_lo) || le32(nonce_hi)) has >= pow_bits trailing zero bits
 * (read as a little-endian u128). */
int cm31_grind_blake2s(const uint8_t digest[32], uint32_t pow_bits, uint64_t* nonce_out);

/* ------------------------------------------------------------------ ComponentProver (constraint eval)
 * S/constraint_framework/src/component.rs:283-424.  The AIR is passed as a register bytecode
 * captured from the component's `evaluate` (see cairo-m_b200/csrc/air_bytecode.hpp).
 *   cols            : host array of device column pointers, each 2^eval_log_size words
 *                     (trace/interaction/preprocessed LDE columns, in program column order)
 *   code            : host array of n_instr 64-bit instructions
 *   consts          : host array of n_consts QM31 constants (4 words each): AIR constants,
 *                     relation elements, random-coefficient powers, cumsum shift
 *   denom_inv_host  : 2^(eval_log_size - trace_log_size) M31 values (already bit-reversed)
 *   acc4            : 4 accumulator columns, RMW:  acc[row] += row_res * denom_inv[row >> trace_log_size]
 */
int cm31_constraint_eval(const uint32_t* const* cols, size_t n_cols, uint32_t trace_log_size,
                         uint32_t eval_log_size, const uint64_t* code, size_t n_instr, uint32_t n_regs,
                         const uint32_t* consts, size_t n_consts, const uint32_t* denom_inv_host,
                         uint32_t* const acc4[4]);

/* Programs of cairo-m's fixed component set have AOT-specialised kernels (csrc/generated/, selected
 * by a hash of `code`); any other program runs on the bytecode interpreter.  mode 1 forces the
 * interpreter (parity tests run both). */
int cm31_set_air_mode(int mode);

/* ------------------------------------------------------------------ witness generation helpers
 * Generic AIR program over the 2^log_size TRACE rows (same bytecode as cm31_constraint_eval):
 *  - LogupTraceGenerator (S/constraint_framework/src/logup.rs:123-320): the logup program stores,
 *    for every logup batch k, the cumulative column col_k = col_{k-1} + num_k/den_k
 *    (OP_STORE_E into out_cols[4k..4k+4]);
 *  - lookup multiplicities (P/src/preprocessed/range_check/range_check_macro.rs:72-84,
 *    P/src/components/opcodes/mod.rs:83-105): OP_HIST atomically counts looked-up values into
 *    out_cols[k] used as a bin array. */
int cm31_air_program(const uint32_t* const* in_cols, size_t n_in, uint32_t* const* out_cols, size_t n_out,
                     uint32_t log_size, const uint64_t* code, size_t n_instr, uint32_t n_regs,
                     const uint32_t* consts, size_t n_consts);
/* Lookup multiplicities (P/src/preprocessed/range_check/range_check_macro.rs:72-84, P/src/preprocessed/bitwise.rs:86-109): a
 * program whose OP_HIST instructions count looked-up values into `bins` (2^log_bins words).  A value >= 2^log_bins is NOT
 * counted (no out-of-bounds write) and raises the device error word: the next cm31_air_error_check returns non-zero with
 * "lookup outside its table" (the reference panics on the slice index).  cm31_air_program refuses programs with OP_HIST. */
int cm31_air_lookups(const uint32_t* const* in_cols, size_t n_in, uint32_t* bins, uint32_t log_bins, uint32_t log_size,
                     const uint64_t* code, size_t n_instr, uint32_t n_regs, const uint32_t* consts, size_t n_consts);
/* MANY small programs (log_size <= 12, n_regs <= 512) in ONE launch: every item is a cm31_air_program call (hist_bins == 0) or a
 * cm31_air_lookups call (hist_bins = 2^log_bins, out_cols[0] = the bin column).  Items must be independent of each other.
 * cairo-m proves 34 components per segment whatever the program; the ones a program does not use are 16 padding rows each,
 * and their trace-fill (Claim::write_trace), lookup and logup programs are pure launch latency one by one. */
typedef struct cm31_air_batch_item {
    const uint32_t* const* in_cols; /* host array of n_in device pointers */
    size_t n_in;
    uint32_t* const* out_cols;
    size_t n_out;
    uint32_t log_size;
    const uint64_t* code;
    size_t n_instr;
    uint32_t n_regs;
    const uint32_t* consts;
    size_t n_consts;
    uint32_t hist_bins;
} cm31_air_batch_item;
int cm31_air_program_batch(const cm31_air_batch_item* items, size_t n_items);
/* reads (one 4-byte copy, synchronises the current lane) and clears the error bits AIR programs raised since the last call */
int cm31_air_error_check(void);
/* finalize_last (logup.rs:211-251): claimed_sum = sum(last col); last col -= claimed_sum/n;
 * inclusive prefix sum in coset order (simd/prefix_sum.rs:19, index map core/utils.rs:121-143). */
int cm31_logup_finalize_last(uint32_t* const last4[4], uint32_t log_size, uint32_t claimed_sum_out[4]);
/* same, stream-ordered: the claimed sum lands in 4 DEVICE words; the prover reads the sums of all
 * components with one copy after the interaction trace is generated */
int cm31_logup_finalize_last_async(uint32_t* const last4[4], uint32_t log_size, uint32_t* claimed_sum_dev);
/* the same for many small columns (log_size <= 11) in one launch */
typedef struct cm31_logup_finalize_item {
    uint32_t* last4[4];
    uint32_t log_size;
    uint32_t* claimed_sum_dev;
} cm31_logup_finalize_item;
int cm31_logup_finalize_small_batch(const cm31_logup_finalize_item* items, size_t n_items);
/* multiplicity histograms (P/src/preprocessed/range_check/range_check_macro.rs:72-84) */
int cm31_histogram(const uint32_t* values, size_t n, uint32_t* bins, uint32_t log_bins);
/* Pack::pack for ExecutionBundle + get_access_field (P/src/utils/execution_bundle.rs:31-75, P/src/utils/data_accesses.rs:10-28):
 * AoS bundles (12 words: pc, fp, clock, inst_prev_clock, inst[6], span start, span len) + the global access log (4 words per
 * access) -> the SoA input columns of a component's trace program, padded to 2^log_size rows with ExecutionBundle::default()
 * (P/src/adapter/memory.rs:112-124).  out_cols: 10 bundle columns, then 4 columns (address, prev_clock, prev_value, value) per
 * access slot; the _slots form writes only the first n_access_slots slots (the ones the component's program reads). */
int cm31_unpack_bundles(const uint32_t* bundles_dev, size_t n_real, uint32_t log_size, const uint32_t* accesses_dev,
                        size_t n_accesses, uint32_t* const* out_cols);
int cm31_unpack_bundles_slots(const uint32_t* bundles_dev, size_t n_real, uint32_t log_size, const uint32_t* accesses_dev,
                              size_t n_accesses, uint32_t* const* out_cols, uint32_t n_access_slots);
/* AoS rows of n_fields words (memory / merkle / clock_update / poseidon2 inputs: P/src/components/memory.rs:105-140,
 * merkle.rs:103-130, clock_update.rs:87-102, poseidon2.rs:180-200) -> n_fields columns of 2^log_size rows, zero padded */
int cm31_unpack_rows(const uint32_t* rows_dev, size_t n_real, uint32_t n_fields, uint32_t log_size, uint32_t* const* out_cols);
/* col[i] = i: the range-check preprocessed columns (P/src/preprocessed/range_check/mod.rs: RangeCheck::gen_column_simd) */
int cm31_iota(uint32_t* col, size_t n);
/* column k (0: op id, 1: a, 2: b, 3: result) of the stacked and/or/xor table (P/src/preprocessed/bitwise.rs:253-290), 2^18 rows */
int cm31_bitwise_table_col(int k, uint32_t* col);

/* ------------------------------------------------------------------ whole proofs
 * prove_cairo_m::<Blake2sMerkleChannel> (P/src/prover.rs:23-147) over the ops above.  The input
 * handle owns the reference's ProverInput (P/src/adapter/mod.rs:97-193): per-opcode ExecutionBundles,
 * the data-access log, boundary memory, clock-update rows, Merkle nodes.  A handle comes from cm31_input_create (a caller
 * that ran its own adapter) or from cm31_adapter_import (the runner's logs, adapted on the device). */
typedef struct cm31_prover_input cm31_prover_input;
/* Allocates the device buffers of this handle's prover input and RECORDS its host->device copies (cm31_bg_defer); returns at
 * once.  A proof that is running, or the next one, releases them at its STARK phase as a throttled copy kernel (DESIGN.md
 * section 5); the next cm31_prove_cairo_m[_async] on the handle consumes the staged input (oldest first) instead of copying
 * itself, releasing the copies by plain DMA if no proof did: the upload of segment i+1 overlaps the proof of segment i. */
int cm31_input_prefetch(cm31_prover_input* h);
int cm31_input_destroy(cm31_prover_input* h);
/* info[0] VM steps, [1] data accesses, [2] boundary-memory rows, [3] return value, [4] input bytes staged per proof */
int cm31_input_info(const cm31_prover_input* h, uint64_t info[5]);
/* stage the input in HBM once (later proofs on this handle skip the host->device copy) / drop that copy */
int cm31_input_upload(cm31_prover_input* h);
int cm31_input_release_device(cm31_prover_input* h);
/* proof bytes (ProofWriter layout, cairo/prover.hpp) into proof_out; timings_ms (optional, 5 doubles):
 * preprocessed, trace, interaction, stark, total.  Returns non-zero with
 * "ConstraintsNotSatisfied" (S/prover/src/core/prover/mod.rs:76-82) if the OODS check fails. */
int cm31_prove_cairo_m(const cm31_prover_input* h, uint32_t pow_bits, uint32_t n_queries, uint8_t* proof_out,
                       size_t proof_cap, size_t* proof_len, double* timings_ms);
/* Asynchronous form for a prover fed with a stream of segments: returns once the proof's kernels and the device->host copy of
 * its decommitment values are enqueued.  proof_out / proof_len are written later -- while the NEXT proof of this thread keeps
 * the GPU busy, or by cm31_prove_wait() -- and must stay valid until then; one proof may be pending at a time.  The bytes are
 * identical to cm31_prove_cairo_m's.  When the NEXT call returns 0 the previous proof is complete and its bytes are valid;
 * cm31_prove_wait() completes the pending proof and returns its status. */
int cm31_prove_cairo_m_async(const cm31_prover_input* h, uint32_t pow_bits, uint32_t n_queries, uint8_t* proof_out,
                             size_t proof_cap, size_t* proof_len, double* timings_ms);
int cm31_prove_wait(void);
/* The reference's proof wire format: serde JSON of Proof<Blake2sMerkleHasher> (P/src/lib.rs:61-73; CommitmentSchemeProof
 * S/prover/src/core/pcs/prover.rs:156-165, FriProof S/prover/src/core/fri.rs:675-699, MerkleDecommitment
 * S/prover/src/core/vcs/prover.rs:163-173), as `cairo-m-prover --output` writes it with sonic_rs (P/src/main.rs:88) and
 * `verify_cairo_m` reads it.  The blob of cm31_prove_cairo_m and the JSON carry the same proof: to_json / from_json are
 * inverse of each other (host only, no device work).  json_out may be NULL to query the length (without the final NUL). */
int cm31_proof_to_json(const uint8_t* proof, size_t proof_len, char* json_out, size_t cap, size_t* json_len);
int cm31_proof_from_json(const char* json, size_t json_len, uint8_t* proof_out, size_t cap, size_t* proof_len);
int cm31_prove_cairo_m_json(const cm31_prover_input* h, uint32_t pow_bits, uint32_t n_queries, char* json_out, size_t cap,
                            size_t* json_len);
/* The reference's ProverInput (P/src/adapter/mod.rs:40-95: Instructions {initial/final registers, states_by_opcodes,
 * data_accesses}, Memory {initial_memory, final_memory, clock_update_data}, MerkleTrees, public_address_ranges) as flat
 * u32 tables — what a Rust caller that ran its own adapter hands over.  Record layouts (words):
 *   bundle (12)       pc, fp, clock, inst_prev_clock, inst[6], access span start, access span len   (ExecutionBundle)
 *   data access (4)   address, prev_clock, prev_value, value                                       (DataAccess, memory.rs:57-68)
 *   memory row (8)    address, clock, value[4], multiplicity, root   (initial/final_memory entries, ascending address)
 *   clock update (6)  address, prev_clock, value[4]
 *   merkle node (9)   index, depth, left, right, parent, left/right/parent multiplicity, root   (initial tree, then final tree)
 * states_by_opcodes: group g holds the bundles of opcode opcode_ids[g], rows bundle_start[g] .. bundle_start[g+1] of `bundles`,
 * in execution order. */
typedef struct cm31_prover_input_desc {
    uint32_t initial_pc, initial_fp, final_pc, final_fp;
    uint32_t public_ranges[6]; /* program, input, output: [start, end) each */
    uint32_t initial_root, final_root;
    uint64_t n_steps;
    uint64_t n_opcodes;
    const uint32_t* opcode_ids;   /* n_opcodes */
    const uint64_t* bundle_start; /* n_opcodes + 1 */
    const uint32_t* bundles;
    const uint32_t* data_accesses;
    uint64_t n_data_accesses;
    const uint32_t* initial_memory;
    uint64_t n_initial_memory;
    const uint32_t* final_memory;
    uint64_t n_final_memory;
    const uint32_t* clock_updates;
    uint64_t n_clock_updates;
    const uint32_t* merkle_nodes;
    uint64_t n_merkle_nodes;
} cm31_prover_input_desc;
/* copies the tables into a new handle (prove with cm31_prove_cairo_m, stage with cm31_input_upload / cm31_input_prefetch) */
int cm31_input_create(const cm31_prover_input_desc* desc, cm31_prover_input** out);
/* the same view of an existing host-adapted handle (pointers stay valid until the handle is destroyed or tampered with) */
int cm31_input_describe(cm31_prover_input* h, cm31_prover_input_desc* out);

/* ------------------------------------------------------------------ adapter on the device
 * import_from_runner_output (P/src/adapter/mod.rs:233-…; import_internal :97-193; ExecutionBundleIterator and Memory::push,
 * P/src/adapter/memory.rs:264-403, 470-…) with the per-step work done in HBM: the runner's logs go in, a prover input
 * resident on the device comes out (prev clock / prev value of every access, clock-update rows, per-opcode bundles).
 *   trace          : IoTraceEntry {fp, pc} x n_trace   (P/src/adapter/io.rs:38-43; one per step + the final state)
 *   memory_trace   : IoMemoryEntry {address, value[4]} x n_mem, access order (io.rs:54-59)
 *   initial_memory : preloaded cells, QM31 x n_initial, address = index (Segment::initial_memory)
 *   public_ranges  : program [start, end), input [start, end), output [start, end) (PublicAddressRanges)
 * Errors (non-zero + cm31_last_error): "empty trace", "invalid opcode", "unexpected end of the memory trace",
 * "unexpected memory access" — VmImportError's cases (io.rs:12-36).  Stricter than the reference in one point: log entries
 * left over after the last step are an error instead of being ignored. */
int cm31_adapter_import(const uint32_t* trace, size_t n_trace, const uint32_t* memory_trace, size_t n_mem,
                        const uint32_t* initial_memory, size_t n_initial, const uint32_t public_ranges[6],
                        cm31_prover_input** out);
/* Pipelined form (continuation segments): cm31_adapter_prefetch allocates the device buffers and RECORDS the upload of a
 * segment's logs; a proof that is running (or runs next) releases it at its STARK phase as a throttled copy kernel,
 * cm31_adapter_import_prefetched at the latest (plain DMA).  The logs must stay valid (and should be page-locked) until
 * cm31_adapter_import_prefetched, which consumes the handle, returns; the proof of the previous segment runs while the next
 * logs arrive. */
typedef struct cm31_adapter_logs cm31_adapter_logs;
int cm31_adapter_prefetch(const uint32_t* trace, size_t n_trace, const uint32_t* memory_trace, size_t n_mem,
                          const uint32_t* initial_memory, size_t n_initial, const uint32_t public_ranges[6],
                          cm31_adapter_logs** out);
int cm31_adapter_import_prefetched(cm31_adapter_logs* logs, cm31_prover_input** out);
int cm31_adapter_logs_destroy(cm31_adapter_logs* logs); /* only for logs that were never imported */
/* The three phases cm31_adapter_import is made of, for a caller that owns the output buffers (the Rust shim allocates the
 * per-opcode row buffers itself once phase 1 has told it their sizes).  `plan` is an opaque device-side work area.
 *   stage_logs  : device buffers + upload of the logs; background != 0 issues the copies on the background copy stream
 *   scan(_staged): everything whose size is data dependent -> counts[0..63] steps per opcode, [64] data accesses,
 *                 [65] clock-update rows, [66] distinct cells touched (cm31_adapter_scan = stage_logs + scan_staged)
 *   emit        : opcode_rows_dev[op] = device buffer for the bundles of opcode op (12 words per step, NULL iff no steps),
 *                 accesses_dev 4 words per access, clock_update_dev 6 words per row, cells_host 10 words per distinct cell
 *                 {address, first value[4], last value[4], last clock} in ascending address order (HOST pointer)
 *   free        : releases the plan (after emit, or on any error path) */
int cm31_adapter_stage_logs(const uint32_t* trace_host, size_t n_trace, const uint32_t* mem_host, size_t n_mem,
                            const uint32_t* init_host, size_t n_init, int background, void** plan_out);
int cm31_adapter_scan_staged(void* plan, uint64_t counts_out[67]);
int cm31_adapter_scan(const uint32_t* trace_host, size_t n_trace, const uint32_t* mem_host, size_t n_mem,
                      const uint32_t* init_host, size_t n_init, void** plan_out, uint64_t counts_out[67]);
int cm31_adapter_emit(void* plan, uint32_t* const opcode_rows_dev[64], const uint64_t counts[67], uint32_t* accesses_dev,
                      uint32_t* clock_update_dev, uint32_t* cells_host);
int cm31_adapter_free(void* plan);
/* One table of a resident input read back to the host (parity tests): table 0 = data-access log (4 words per access),
 * 1..26 = opcode components in claim order (12 words per step), 100 = memory rows (8), 101 = merkle rows (9),
 * 102 = clock-update rows (6), 103 = poseidon2 states (16).  out may be NULL to query the size. */
int cm31_input_staged_words(const cm31_prover_input* h, uint32_t table, uint32_t* out, size_t cap_words, size_t* n_words_out);

/* The AIR shapes of the 34 components as captured from their `evaluate` bodies — what FrameworkComponent::new learns from its
 * InfoEvaluator pass (S/constraint_framework/src/component.rs:139-180, info.rs) — as JSON: {"relations": {name: size},
 * "components": [{name, opcodes, n_trace_columns, n_interaction_columns, n_preprocessed_columns, n_constraints,
 * n_cumsum_columns, n_lookups, lookups: {relation: [count, widest tuple]}}]} in claim order.  Host only (no device work).
 * The Rust shim asserts these against its own InfoEvaluator counts before it trusts a generated kernel; tests/test_air_shapes.py
 * checks them against the constants of P/src/components/**. buf may be NULL to query the length. */
int cm31_air_shapes(char* buf, size_t cap, size_t* len);


/* ------------------------------------------------------------------ test conveniences (cm31_test_*)
 * NOT part of the drop-in surface: built-in hand-assembled programs run on the host VM + host adapter of csrc/cairo/vm.hpp
 * (the serial steps BEFORE the hot path), a tamper hook, and the bring-up AIR.  Tests, bench.py and smoke() use them to have
 * inputs; a cairo-m-prover integration never calls them. */
/* program_id 0 = fibonacci_loop(n); 1 = array_sum(n): call/ret, frame pointer, double-deref, assert, le; 2 = u32_counter(n):
 * u32 limb ops; 3 = u32_mix(n): u32 mul/divrem/eq/lt + two-word *_fp_imm u32 instructions; 4 = sha256 (n chained compressions of the padded block of "abc"); 5 = all_opcodes (n iterations, every opcode family) */
int cm31_test_program_input_create(uint32_t program_id, uint32_t n, cm31_prover_input** out);
int cm31_test_fib_input_create(uint32_t n, cm31_prover_input** out);
/* corrupt the adapter output so a store_fp_fp constraint fails (kind 0: a written value, 1: an operand read) */
int cm31_test_input_tamper(cm31_prover_input* h, uint32_t kind);
/* The runner's output for a built-in program: the host VM only, no adapter. */
typedef struct cm31_test_vm_trace cm31_test_vm_trace;
int cm31_test_vm_trace_create(uint32_t program_id, uint32_t n, cm31_test_vm_trace** out);
/* info[0] trace entries (steps + 1), [1] memory-log entries, [2] preloaded cells, [3] return value */
int cm31_test_vm_trace_info(const cm31_test_vm_trace* h, uint64_t info[4]);
/* pointers into the handle, laid out as cm31_adapter_import takes them */
int cm31_test_vm_trace_data(const cm31_test_vm_trace* h, const uint32_t** trace, const uint32_t** memory_trace,
                            const uint32_t** initial_memory, uint32_t public_ranges[6]);
int cm31_test_vm_trace_destroy(cm31_test_vm_trace* h);
/* Continuation segments: the same run cut every `segment_steps` steps as RunnerOptions::max_steps cuts it
 * (R/crates/runner/src/vm/mod.rs:158-285); segment `index` as a runner output of its own, *n_segments = how many there are.
 * The reference chains consecutive segment proofs by their memory roots (P/tests/prover.rs:204-243). */
int cm31_test_vm_segment_create(uint32_t program_id, uint32_t n, uint64_t segment_steps, uint32_t index, uint32_t* n_segments,
                                cm31_test_vm_trace** out);
/* the host restatement of import_from_runner_output (csrc/cairo/vm.hpp) on a given runner output */
int cm31_test_vm_trace_to_input(const cm31_test_vm_trace* t, cm31_prover_input** out);
/* S/examples/src/wide_fibonacci/mod.rs:22-43 — the bring-up AIR (parity tests only) */
int cm31_test_prove_wide_fibonacci(uint32_t log_n_rows, uint32_t n_cols, uint32_t pow_bits, uint32_t n_queries,
                                   uint8_t* proof_out, size_t proof_cap, size_t* proof_len);

/* ------------------------------------------------------------------ per-kernel device timing
 * (the reference's tracing spans, S/prover/src/tracing/mod.rs:22-50).  When enabled every kernel
 * launch is bracketed by CUDA events on the launch stream; the report is a JSON array of
 * {"kernel", "ms", "launches", "alg_bytes"} (algorithmic bytes per SURVEY.md §8d). */
int cm31_profile_enable(int on);
int cm31_profile_reset(void);
int cm31_profile_launches(uint64_t* out); /* kernels launched since the last reset (always counted) */
int cm31_profile_report(char* buf, size_t cap, size_t* len);
int cm31_profile_trace(char* buf, size_t cap, size_t* len); /* CSV: kernel,start_ms,dur_ms per launch */

/* Integer issue-rate microbenchmark (the binding roof of this path; not in MEASURED_PEAKS.json):
 * tera lane-operations per second for [0] ALU-pipe ops (LOP3/SHF), [1] FMA-pipe ops (IMAD), [2] the 1:1 mix. */
int cm31_int_peak(double tera_lane_ops_per_s[3]);

#ifdef __cplusplus
}
#endif
#endif /* CM31_H */

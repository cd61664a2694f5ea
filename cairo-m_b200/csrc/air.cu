// AIR program kernel: constraint/quotient evaluation on the LDE, logup column generation and
// lookup-multiplicity histograms, all driven by the register bytecode of host/air_expr.hpp.
//
// Replaces the per-row loops of
//   external/stwo/crates/constraint_framework/src/component.rs:376-423 (SimdDomainEvaluator loop),
//   constraint_framework/src/logup.rs:123-320 (LogupTraceGenerator),
//   crates/prover/src/preprocessed/range_check/range_check_macro.rs:72-84 (multiplicity counts).
// One thread per row; every column access is a coalesced 4-byte-per-lane load; the program and
// its constants are warp-uniform reads.  The register file lives in per-thread local memory
// (hardware-interleaved, L1 resident), sized by the program's register allocation.
#include <cstdlib>
#include <unordered_map>

#include "air_gen.cuh"
#include "common.cuh"
#include "host/air_expr.hpp"

namespace cm31 {

__device__ __forceinline__ u32 dev_offset_row(u32 row, u32 trace_log, u32 eval_log, int off) {
    // core/utils.rs:74-90 offset_bit_reversed_circle_domain_index
    u32 idx = bit_reverse(row, eval_log);
    u32 half = 1u << (eval_log - 1);
    int step = off * (int)(1u << (eval_log - trace_log - 1));
    u32 mask = half - 1;
    if (idx < half) idx = (u32)((int)idx + step) & mask;
    else idx = ((u32)((int)(idx - half) - step) & mask) + half;
    return bit_reverse(idx, eval_log);
}

__device__ __forceinline__ QM31 rd4(const u32* r, u32 i) { return qm_make(r[i], r[i + 1], r[i + 2], r[i + 3]); }
__device__ __forceinline__ void wr4(u32* r, u32 i, QM31 v) {
    r[i] = v.a;
    r[i + 1] = v.b;
    r[i + 2] = v.c;
    r[i + 3] = v.d;
}

// one row of a bytecode program; returns the row's constraint accumulator
template <int NREGS>
__device__ __forceinline__ QM31 air_interpret_row(const u32* const* __restrict__ in_cols, u32* const* __restrict__ out_cols, u32 row,
                                                  u32 row_log, u32 trace_log, const uint64_t* __restrict__ code, u32 n_instr,
                                                  const u32* __restrict__ consts, u32 hist_bins, u32* err) {
    u32 regs[NREGS];
    QM31 acc = qm_zero();
    for (u32 pc = 0; pc < n_instr; pc++) {
        const uint64_t ins = __ldg(code + pc);
        const u32 op = (u32)(ins & 0xff);
        const u32 dst = (u32)((ins >> 8) & 0xffff);
        const u32 a = (u32)((ins >> 24) & 0xfffff);
        const u32 b = (u32)((ins >> 44) & 0xfffff);
        switch (op) {
            case OP_LOAD: {
                u32 r = row;
                if (b != 0) {
                    int off = (int)(b << 12) >> 12;  // sign-extend 20 bits
                    r = dev_offset_row(row, trace_log, row_log, off);
                }
                regs[dst] = __ldg(in_cols[a] + r);
                break;
            }
            case OP_CONSTF: regs[dst] = __ldg(consts + a); break;
            case OP_CONSTE:
                regs[dst] = __ldg(consts + a);
                regs[dst + 1] = __ldg(consts + a + 1);
                regs[dst + 2] = __ldg(consts + a + 2);
                regs[dst + 3] = __ldg(consts + a + 3);
                break;
            case OP_ADD: regs[dst] = m31_add(regs[a], regs[b]); break;
            case OP_SUB: regs[dst] = m31_sub(regs[a], regs[b]); break;
            case OP_MUL: regs[dst] = m31_mul(regs[a], regs[b]); break;
            case OP_NEG: regs[dst] = m31_neg(regs[a]); break;
            case OP_EADD: wr4(regs, dst, qm_add(rd4(regs, a), rd4(regs, b))); break;
            case OP_ESUB: wr4(regs, dst, qm_sub(rd4(regs, a), rd4(regs, b))); break;
            case OP_EMUL: wr4(regs, dst, qm_mul(rd4(regs, a), rd4(regs, b))); break;
            case OP_ENEG: wr4(regs, dst, qm_neg(rd4(regs, a))); break;
            case OP_EMULF: wr4(regs, dst, qm_mul_m31(rd4(regs, a), regs[b])); break;
            case OP_EADDF: wr4(regs, dst, qm_add_m31(rd4(regs, a), regs[b])); break;
            case OP_ESUBF: wr4(regs, dst, qm_sub_m31(rd4(regs, a), regs[b])); break;
            case OP_F2E: wr4(regs, dst, qm_from_m31(regs[a])); break;
            case OP_MOV: regs[dst] = regs[a]; break;
            case OP_EINV: wr4(regs, dst, qm_inv(rd4(regs, a))); break;
            case OP_CONSTRAINT_E:
                acc = qm_add(acc, qm_mul(qm_make(__ldg(consts + b), __ldg(consts + b + 1), __ldg(consts + b + 2), __ldg(consts + b + 3)),
                                         rd4(regs, a)));
                break;
            case OP_CONSTRAINT_F:
                acc = qm_add(acc, qm_mul_m31(qm_make(__ldg(consts + b), __ldg(consts + b + 1), __ldg(consts + b + 2), __ldg(consts + b + 3)),
                                             regs[a]));
                break;
            case OP_STORE_E:
                out_cols[b][row] = regs[a];
                out_cols[b + 1][row] = regs[a + 1];
                out_cols[b + 2][row] = regs[a + 2];
                out_cols[b + 3][row] = regs[a + 3];
                break;
            case OP_STORE_F: out_cols[b][row] = regs[a]; break;
            case OP_HIST: {
                // lookups of a fibonacci-like trace hit a handful of bins (clock deltas, small offsets):
                // aggregate equal values across the warp so each distinct value costs one atomic
                gen_hist(out_cols[b], regs[a], hist_bins, err);  // bounds-checked (air_gen.cuh)
                break;
            }
            case OP_INV: regs[dst] = m31_inv(regs[a]); break;
            case OP_SHR: regs[dst] = regs[a] >> b; break;
            case OP_AND: regs[dst] = regs[a] & b; break;
            case OP_ROWLT: regs[dst] = row < __ldg(consts + a) ? 1u : 0u; break;
            case OP_LE: regs[dst] = regs[a] <= regs[b] ? 1u : 0u; break;
            case OP_DIVC: regs[dst] = regs[a] / b; break;
            case OP_MODC: regs[dst] = regs[a] % b; break;
            case OP_U32DIVREM: regs[dst] = u32_divrem_part(qm_make(regs[a], regs[a + 1], regs[a + 2], regs[a + 3]), b); break;
            default: break;
        }
    }
    return acc;
}

template <int NREGS>
__global__ void __launch_bounds__(128) air_program_kernel(const u32* const* __restrict__ in_cols, u32* const* __restrict__ out_cols,
                                                          u32 row_log, u32 trace_log, const uint64_t* __restrict__ code, u32 n_instr,
                                                          const u32* __restrict__ consts, const u32* __restrict__ denom_inv,
                                                          u32* acc0, u32* acc1, u32* acc2, u32* acc3, u32 hist_bins, u32* err) {
    const u32 row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (1u << row_log)) return;
    const QM31 acc = air_interpret_row<NREGS>(in_cols, out_cols, row, row_log, trace_log, code, n_instr, consts, hist_bins, err);
    if (acc0 != nullptr) {
        // component.rs:413-421: col[row] += row_res * denom_inv[row >> trace_log]
        u32 di = __ldg(denom_inv + (row >> trace_log));
        QM31 v = qm_mul_m31(acc, di);
        acc0[row] = m31_add(acc0[row], v.a);
        acc1[row] = m31_add(acc1[row], v.b);
        acc2[row] = m31_add(acc2[row], v.c);
        acc3[row] = m31_add(acc3[row], v.d);
    }
}

// MANY small programs in one launch (blockIdx.y = program): the trace-fill, lookup and logup programs of the components that
// only prove their 16 padding rows -- 21 of the 26 opcode components of a fibonacci_loop proof -- are pure launch latency when
// issued one by one (~7 us of host time each, ~135 launches per proof).  The descriptors, pointer arrays, bytecode and constants
// of the whole batch travel in one table.  No constraint programs here: those accumulate into shared per-size columns.
struct AirBatchDesc {
    const u32* const* in;
    u32* const* out;
    const uint64_t* code;
    const u32* consts;
    u32 row_log, n_instr, hist_bins, pad;
};
template <int NREGS>
__global__ void __launch_bounds__(128) air_program_batch_kernel(const AirBatchDesc* __restrict__ descs, u32* err) {
    const AirBatchDesc d = descs[blockIdx.y];
    const u32 row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (1u << d.row_log)) return;
    air_interpret_row<NREGS>(d.in, d.out, row, d.row_log, d.row_log, d.code, d.n_instr, d.consts, d.hist_bins, err);
}

static int g_gen_sync = getenv("CM31_GEN_SYNC") ? atoi(getenv("CM31_GEN_SYNC")) : 1;  // CTA barriers in the generated bodies (air_gen.cuh GEN_SYNC)
static int g_air_mode = 0;  // 0 = AOT-specialised kernel when one exists, 1 = always the bytecode interpreter

// device word collecting AIR_ERR_* bits of every AIR program launched since the last cm31_air_error_check
static u32* air_err_flag() {
    static u32* flag[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!flag[dev & 63]) {
        if (cudaMalloc(&flag[dev & 63], 4) != cudaSuccess) return nullptr;
        cudaMemset(flag[dev & 63], 0, 4);
    }
    return flag[dev & 63];
}

// ALGORITHMIC M31 operations of one row of a bytecode program (S/constraint_framework/src/info.rs arithmetic counts, in M31
// units: QM31 mul = 16 mul + 15 add, QM31 x M31 = 4, QM31 add = 4, QM31 inverse ~ 60 + an M31 inverse of 37 mul)
static uint64_t program_m31_ops(const uint64_t* code, size_t n_instr, uint64_t hash) {
    static std::unordered_map<uint64_t, uint64_t> cache;  // programs are few and immutable: count once per program
    auto it = cache.find(hash);
    if (it != cache.end()) return it->second;
    uint64_t ops = 0;
    for (size_t i = 0; i < n_instr; i++) {
        switch ((u32)(code[i] & 0xff)) {
            case OP_ADD: case OP_SUB: case OP_MUL: case OP_NEG: ops += 1; break;
            case OP_EADD: case OP_ESUB: case OP_ENEG: ops += 4; break;
            case OP_EMUL: ops += 31; break;
            case OP_EMULF: ops += 4; break;
            case OP_EADDF: case OP_ESUBF: ops += 1; break;
            case OP_EINV: ops += 97; break;
            case OP_INV: ops += 37; break;
            case OP_CONSTRAINT_E: ops += 35; break;
            case OP_CONSTRAINT_F: ops += 8; break;
            default: break;
        }
    }
    cache[hash] = ops;
    return ops;
}

static int run_program(const uint32_t* const* in_cols, size_t n_in, uint32_t* const* out_cols, size_t n_out, u32 row_log,
                       u32 trace_log, const uint64_t* code, size_t n_instr, u32 n_regs, const u32* consts, size_t n_consts,
                       const u32* denom_inv_host, size_t n_denom, uint32_t* const* acc4, u32 hist_bins = 0) {
    u32* err = air_err_flag();
    CM_REQUIRE(err != nullptr, "air: cannot allocate the error word");
    CM_REQUIRE(row_log <= 30, "air: too many rows");
    CM_REQUIRE(n_regs <= 2048, "air: program needs more than 2048 registers");
    CM_REQUIRE(trace_log <= row_log, "air: trace domain larger than the evaluation domain");
    const uint64_t code_hash = air_code_hash(code, n_instr);
    const GenEntry* gen = g_air_mode == 0 ? air_gen_lookup(code_hash) : nullptr;
    if (gen) {
        DeviceTable ddenom_gen;
        if (acc4) {
            CM_REQUIRE(n_denom == ((size_t)1 << (row_log - trace_log)), "air: wrong number of denominator inverses");
            if (int e = ddenom_gen.upload(denom_inv_host, n_denom * 4)) return e;
        }
        GenLaunch gl{in_cols, n_in, out_cols, n_out, row_log, trace_log, consts, n_consts, (const u32*)ddenom_gen.d, acc4, g_gen_sync, hist_bins, err};
        size_t rows = (size_t)1 << row_log;
        ProfScope prof(acc4 ? "constraint_eval" : "air_program", acc4 ? 4ull * rows * n_in + 32ull * rows : 4ull * rows * (n_in + n_out));
        if (prof_enabled()) prof_ops(rows * (program_m31_ops(code, n_instr, code_hash) + (acc4 ? 8 : 0)));
        if (int e = gen->launch(gl)) return e;
        CM_LAUNCH_CHECK();
        return 0;
    }
    DeviceTable din, dout, dcode, dconsts, ddenom;
    if (int e = din.upload(in_cols, n_in * sizeof(void*))) return e;
    if (int e = dout.upload(out_cols, n_out * sizeof(void*))) return e;
    if (int e = dcode.upload(code, n_instr * 8)) return e;
    if (int e = dconsts.upload(consts, n_consts * 4)) return e;
    if (acc4) {
        CM_REQUIRE(n_denom == ((size_t)1 << (row_log - trace_log)), "air: wrong number of denominator inverses");
        if (int e = ddenom.upload(denom_inv_host, n_denom * 4)) return e;
    }
    size_t n = (size_t)1 << row_log;
    unsigned threads = 128, blocks = (unsigned)((n + threads - 1) / threads);
    u32 *a0 = acc4 ? acc4[0] : nullptr, *a1 = acc4 ? acc4[1] : nullptr, *a2 = acc4 ? acc4[2] : nullptr,
        *a3 = acc4 ? acc4[3] : nullptr;
    ProfScope prof(acc4 ? "constraint_eval" : "air_program", acc4 ? 4ull * n * n_in + 32ull * n : 4ull * n * (n_in + n_out));
    if (prof_enabled()) prof_ops(n * (program_m31_ops(code, n_instr, code_hash) + (acc4 ? 8 : 0)));
#define CM_AIR_LAUNCH(NR)                                                                                              \
    air_program_kernel<NR><<<blocks, threads, 0, stream()>>>((const u32* const*)din.d, (u32* const*)dout.d, row_log,   \
                                                             trace_log, (const uint64_t*)dcode.d, (u32)n_instr,        \
                                                             (const u32*)dconsts.d, (const u32*)ddenom.d, a0, a1, a2, a3, hist_bins, err)
    if (n_regs <= 64) CM_AIR_LAUNCH(64);
    else if (n_regs <= 128) CM_AIR_LAUNCH(128);
    else if (n_regs <= 256) CM_AIR_LAUNCH(256);
    else if (n_regs <= 512) CM_AIR_LAUNCH(512);
    else if (n_regs <= 1024) CM_AIR_LAUNCH(1024);
    else CM_AIR_LAUNCH(2048);
#undef CM_AIR_LAUNCH
    CM_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ logup finalize_last
// (logup.rs:211-251): claimed_sum = Σ rows of the last cumulative column; subtract
// claimed_sum/n from every row; inclusive prefix sum in COSET order of the bit-reversed
// circle-domain column (simd/prefix_sum.rs:19-112, index maps core/utils.rs:92-143).
__device__ __forceinline__ u32 coset_pos_to_row(u32 j, u32 L) {
    // coset index j -> circle-domain index -> bit-reversed storage row
    u32 cd = (j & 1) ? (u32)((((u64)2 << L) - j) >> 1) : (j >> 1);
    return bit_reverse(cd, L);
}

constexpr u32 SCAN_CHUNK = 1024;  // elements per CTA in the chunked scan
struct Col4 {
    u32* p[4];
};

__global__ void __launch_bounds__(256) sum4_kernel(Col4 c, size_t n, unsigned long long* sums) {
    __shared__ unsigned long long sh[4][256];
    unsigned long long s[4] = {0, 0, 0, 0};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 4; k++) s[k] += c.p[k][i];
    }
    for (int k = 0; k < 4; k++) sh[k][threadIdx.x] = s[k];
    __syncthreads();
    for (u32 st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st)
            for (int k = 0; k < 4; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int k = 0; k < 4; k++) atomicAdd(&sums[k], sh[k][0] % P);
}
// claimed_sum = sums mod P (-> claimed_out, 4 words); shift = claimed_sum / n (-> shift_out, 4 words)
__global__ void sums_to_shift_kernel(const unsigned long long* sums, u32 n_inv, u32* claimed_out, u32* shift_out) {
    u32 k = threadIdx.x;
    if (k >= 4) return;
    u32 c = m31_reduce64(sums[k]);
    claimed_out[k] = c;
    shift_out[k] = m31_mul(c, n_inv);
}

// The three scan passes handle the 4 coordinate columns in one launch (blockIdx.y = coordinate).
// pass 1: per-chunk totals of (value - shift) in coset order
__global__ void __launch_bounds__(256) scan_chunk_sums_kernel(Col4 c, u32 L, const u32* __restrict__ shift4, u32* chunk_sums, u32 n_chunks) {
    __shared__ u32 sh[256];
    const u32* col = c.p[blockIdx.y];
    const u32 shift = shift4[blockIdx.y];
    u32 chunk = blockIdx.x;
    size_t n = (size_t)1 << L;
    u32 acc = 0;
    for (u32 k = threadIdx.x; k < SCAN_CHUNK; k += 256) {
        size_t j = (size_t)chunk * SCAN_CHUNK + k;
        if (j < n) acc = m31_add(acc, m31_sub(col[coset_pos_to_row((u32)j, L)], shift));
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (u32 st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) sh[threadIdx.x] = m31_add(sh[threadIdx.x], sh[threadIdx.x + st]);
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_sums[(size_t)blockIdx.y * n_chunks + chunk] = sh[0];
}
// pass 2: exclusive scan of the chunk / segment totals (one CTA per coordinate, tiles of 1024; warp-shuffle scans)
__device__ __forceinline__ u32 warp_inclusive_m31(u32 x, u32 lane) {
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x = m31_add(x, t);
    }
    return x;
}
__global__ void __launch_bounds__(1024) scan_chunk_offsets_kernel(u32* chunk_sums_all, u32 n_chunks) {
    __shared__ u32 wsum[32];
    __shared__ u32 tile_total;
    u32* chunk_sums = chunk_sums_all + (size_t)blockIdx.x * n_chunks;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 carry = 0;  // identical in every thread
    for (u32 base = 0; base < n_chunks; base += 1024) {
        u32 i = base + threadIdx.x;
        u32 v = i < n_chunks ? chunk_sums[i] : 0;
        u32 x = warp_inclusive_m31(v, lane);
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            u32 w = wsum[lane];
            u32 y = warp_inclusive_m31(w, lane);
            wsum[lane] = m31_sub(y, w);  // exclusive over warps
            if (lane == 31) tile_total = y;
        }
        __syncthreads();
        u32 incl = m31_add(x, wsum[warp]);
        if (i < n_chunks) chunk_sums[i] = m31_add(carry, m31_sub(incl, v));  // exclusive
        carry = m31_add(carry, tile_total);
        __syncthreads();
    }
}

// Segmented form of passes 1 and 3 for large columns.  The coset order is cut into S = 2^s contiguous segments of
// 2^(SEG_Q+1) positions.  With L = 1 + s + SEG_Q, position j = 2y (+1), y = sigma * 2^SEG_Q + i:
//   even j: cd = y        -> row = bitrev_q(i) << (s+1) | bitrev_s(sigma) << 1
//   odd  j: cd = n-1-y    -> row = n - 1 - (row of the even partner)          (every bit complemented)
// so thread tau = bitrev_s(sigma) walks rows (h << (s+1)) | 2 tau and their mirror images: at every step the threads
// of a warp touch CONSECUTIVE even (resp. odd) words — whole sectors, shared with the warp of the mirrored segment —
// instead of one word per sector (pass 1/3 of the chunked form: rows bitrev(cd) of consecutive cd, 8x sector
// amplification, 210-270 GB/s).  Each thread scans its 2^(SEG_Q+1) values sequentially; loads are issued 8 steps ahead.
constexpr u32 SEG_Q = 6;
template <bool APPLY>
__global__ void __launch_bounds__(256) seg_scan_kernel(Col4 c, u32 L, const u32* __restrict__ shift4, u32* seg_all, u32 s) {
    u32* col = c.p[blockIdx.y];
    const u32 shift = shift4[blockIdx.y];
    const u32 S = 1u << s;
    u32 tau = blockIdx.x * blockDim.x + threadIdx.x;
    if (tau >= S) return;
    u32* slot = seg_all + (size_t)blockIdx.y * S + (__brev(tau) >> (32 - s));  // totals / offsets are kept in sigma order
    const u32 n1 = (u32)(((u64)1 << L) - 1);
    u32 acc = APPLY ? *slot : 0u;
    for (u32 i0 = 0; i0 < (1u << SEG_Q); i0 += 8) {
        u32 re[8], xe[8], xo[8];
#pragma unroll
        for (u32 k = 0; k < 8; k++) {
            u32 h = __brev(i0 + k) >> (32 - SEG_Q);
            re[k] = (h << (s + 1)) | (tau << 1);
            xe[k] = col[re[k]];
            xo[k] = col[n1 - re[k]];
        }
#pragma unroll
        for (u32 k = 0; k < 8; k++) {
            acc = m31_add(acc, m31_sub(xe[k], shift));
            if (APPLY) col[re[k]] = acc;
            acc = m31_add(acc, m31_sub(xo[k], shift));
            if (APPLY) col[n1 - re[k]] = acc;
        }
    }
    if (!APPLY) *slot = acc;
}

// pass 3: in-chunk inclusive scan + chunk offset, written back in place
__global__ void __launch_bounds__(256) scan_apply_kernel(Col4 c, u32 L, const u32* __restrict__ shift4, const u32* chunk_offsets_all, u32 n_chunks) {
    __shared__ u32 sh[256];
    u32* col = c.p[blockIdx.y];
    const u32 shift = shift4[blockIdx.y];
    const u32* chunk_offsets = chunk_offsets_all + (size_t)blockIdx.y * n_chunks;
    u32 chunk = blockIdx.x;
    size_t n = (size_t)1 << L;
    const u32 per = SCAN_CHUNK / 256;  // 4 consecutive coset positions per thread
    u32 v[per];
    u32 rows[per];
    u32 acc = 0;
    for (u32 k = 0; k < per; k++) {
        size_t j = (size_t)chunk * SCAN_CHUNK + threadIdx.x * per + k;
        rows[k] = j < n ? coset_pos_to_row((u32)j, L) : 0xffffffffu;
        u32 x = j < n ? m31_sub(col[rows[k]], shift) : 0;
        acc = m31_add(acc, x);
        v[k] = acc;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (u32 st = 1; st < 256; st <<= 1) {
        u32 t = threadIdx.x >= st ? sh[threadIdx.x - st] : 0;
        __syncthreads();
        sh[threadIdx.x] = m31_add(sh[threadIdx.x], t);
        __syncthreads();
    }
    u32 excl = m31_sub(sh[threadIdx.x], acc);
    u32 base = m31_add(chunk_offsets[chunk], excl);
    for (u32 k = 0; k < per; k++)
        if (rows[k] != 0xffffffffu) col[rows[k]] = m31_add(base, v[k]);
}

// Small components (n <= 2^11 rows; most of cairo-m's 34 components for a given program): the
// whole finalize_last — sums, claimed sum, shift, coset-order scan of the 4 coordinates — in ONE
// single-CTA launch.  64 threads per coordinate, each scanning a contiguous run.
constexpr u32 FINALIZE_SMALL_LOG = 11;
__device__ __forceinline__ void logup_finalize_small_body(const Col4& c, u32 L, u32 n_inv, u32* claimed_out) {
    __shared__ u32 vals[4][1u << FINALIZE_SMALL_LOG];
    __shared__ u32 tot[4][64];
    __shared__ u32 shift[4];
    const u32 n = 1u << L;
    const u32 k = threadIdx.x >> 6, t = threadIdx.x & 63;
    u32* col = c.p[k];
    const u32 run = n >= 64 ? n / 64 : 1;  // elements per thread (threads >= n idle when n < 64)
    const bool active = t * run < n;
    u32 acc = 0;
    if (active)
        for (u32 i = 0; i < run; i++) {
            u32 j = t * run + i;
            u32 x = col[coset_pos_to_row(j, L)];
            vals[k][j] = x;
            acc = m31_add(acc, x);
        }
    tot[k][t] = acc;
    __syncthreads();
    if (t == 0) {
        u32 s = 0;
        for (u32 i = 0; i < 64; i++) s = m31_add(s, tot[k][i]);
        claimed_out[k] = s;
        shift[k] = m31_mul(s, n_inv);
    }
    __syncthreads();
    if (!active) return;
    // prefix of the shifted values: offset of this run = sum of previous runs - shift * (elements before)
    u32 before = 0;
    for (u32 i = 0; i < t; i++) before = m31_add(before, tot[k][i]);
    const u32 sh = shift[k];
    u32 running = m31_sub(before, m31_mul(sh, (t * run) % P));
    for (u32 i = 0; i < run; i++) {
        u32 j = t * run + i;
        running = m31_add(running, m31_sub(vals[k][j], sh));
        col[coset_pos_to_row(j, L)] = running;
    }
}
__global__ void __launch_bounds__(256) logup_finalize_small_kernel(Col4 c, u32 L, u32 n_inv, u32* claimed_out) {
    logup_finalize_small_body(c, L, n_inv, claimed_out);
}
struct LogupSmallItem {
    Col4 c;
    u32 L, n_inv;
    u32* claimed;
};
__global__ void __launch_bounds__(256) logup_finalize_small_batch_kernel(const LogupSmallItem* __restrict__ items) {
    const LogupSmallItem it = items[blockIdx.x];
    logup_finalize_small_body(it.c, it.L, it.n_inv, it.claimed);
}

__global__ void histogram_kernel(const u32* values, size_t n, u32* bins, u32 n_bins) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 v = values[i];
    if (v < n_bins) atomicAdd(bins + v, 1u);
}

}  // namespace cm31

using namespace cm31;

extern "C" {

int cm31_set_air_mode(int mode) {
    CM_REQUIRE(mode == 0 || mode == 1, "set_air_mode: 0 (specialised kernels) or 1 (interpreter)");
    g_air_mode = mode;
    return 0;
}

int cm31_constraint_eval(const uint32_t* const* cols, size_t n_cols, uint32_t trace_log_size, uint32_t eval_log_size,
                         const uint64_t* code, size_t n_instr, uint32_t n_regs, const uint32_t* consts, size_t n_consts,
                         const uint32_t* denom_inv_host, uint32_t* const acc4[4]) {
    CM_REQUIRE(acc4 != nullptr && denom_inv_host != nullptr, "constraint_eval: null accumulator");
    return run_program(cols, n_cols, nullptr, 0, eval_log_size, trace_log_size, code, n_instr, n_regs, consts, n_consts,
                       denom_inv_host, (size_t)1 << (eval_log_size - trace_log_size), acc4);
}

int cm31_air_program(const uint32_t* const* in_cols, size_t n_in, uint32_t* const* out_cols, size_t n_out,
                     uint32_t log_size, const uint64_t* code, size_t n_instr, uint32_t n_regs, const uint32_t* consts,
                     size_t n_consts) {
    for (size_t i = 0; i < n_instr; i++)
        CM_REQUIRE((code[i] & 0xff) != OP_HIST, "air_program: histogram programs go through cm31_air_lookups (the bin count bounds OP_HIST)");
    return run_program(in_cols, n_in, out_cols, n_out, log_size, log_size, code, n_instr, n_regs, consts, n_consts,
                       nullptr, 0, nullptr);
}

int cm31_air_lookups(const uint32_t* const* in_cols, size_t n_in, uint32_t* bins, uint32_t log_bins, uint32_t log_size,
                     const uint64_t* code, size_t n_instr, uint32_t n_regs, const uint32_t* consts, size_t n_consts) {
    CM_REQUIRE(bins != nullptr && log_bins <= 30, "air_lookups: bad bin column");
    for (size_t i = 0; i < n_instr; i++) {
        const u32 op = (u32)(code[i] & 0xff), b = (u32)((code[i] >> 44) & 0xfffff);
        CM_REQUIRE(op != OP_STORE_E && op != OP_STORE_F, "air_lookups: a lookup program only counts");
        CM_REQUIRE(op != OP_HIST || b == 0, "air_lookups: one bin column per call");
    }
    uint32_t* out[1] = {bins};
    return run_program(in_cols, n_in, out, 1, log_size, log_size, code, n_instr, n_regs, consts, n_consts, nullptr, 0, nullptr,
                       1u << log_bins);
}

// see air_program_batch_kernel.  Every item is a cm31_air_program call (hist_bins == 0) or a cm31_air_lookups call
// (hist_bins = bin count of out_cols[0]); items must be independent of each other (they run concurrently).
int cm31_air_program_batch(const cm31_air_batch_item* items, size_t n_items) {
    if (n_items == 0) return 0;
    u32* err = air_err_flag();
    CM_REQUIRE(err != nullptr, "air: cannot allocate the error word");
    CM_REQUIRE(items != nullptr && n_items <= 65535, "air_program_batch: bad batch");
    // blob = [descriptors][pointer arrays][bytecode][constants], 8-byte aligned sections
    size_t n_ptrs = 0, n_code = 0, n_consts = 0;
    u32 max_regs = 0, max_log = 0;
    for (size_t i = 0; i < n_items; i++) {
        const cm31_air_batch_item& it = items[i];
        CM_REQUIRE(it.log_size <= 12, "air_program_batch: small programs only");
        CM_REQUIRE((it.n_in == 0 || it.in_cols != nullptr) && (it.n_out == 0 || it.out_cols != nullptr) && it.code != nullptr &&
                       (it.n_consts == 0 || it.consts != nullptr),
                   "air_program_batch: null table in an item");
        CM_REQUIRE(it.hist_bins == 0 || it.n_out >= 1, "air_program_batch: a lookup item needs its bin column");
        CM_REQUIRE(it.n_regs <= 512, "air_program_batch: programs of at most 512 registers (larger register files live in local memory: see CudaBackend::BATCH_MAX_REGS)");
        for (size_t k = 0; k < it.n_instr; k++) {
            const u32 op = (u32)(it.code[k] & 0xff);
            CM_REQUIRE(op != OP_HIST || it.hist_bins != 0, "air_program_batch: a histogram program needs its bin count");
            CM_REQUIRE(it.hist_bins == 0 || (op != OP_STORE_E && op != OP_STORE_F), "air_program_batch: a lookup program only counts");
        }
        n_ptrs += it.n_in + it.n_out;
        n_code += it.n_instr;
        n_consts += it.n_consts + (it.n_consts & 1);
        max_regs = std::max(max_regs, it.n_regs);
        max_log = std::max(max_log, it.log_size);
    }
    const size_t off_ptrs = n_items * sizeof(AirBatchDesc), off_code = off_ptrs + n_ptrs * 8, off_consts = off_code + n_code * 8;
    const size_t bytes = off_consts + n_consts * 4;
    static thread_local std::vector<uint8_t> blob;
    blob.resize(bytes);
    // the device address of the table is known only after the upload slot is taken: stage with offsets, patch after
    DeviceTable table;
    if (int e = table.reserve(bytes)) return e;  // fixes the device address; the bytes follow (fill)
    uint8_t* dbase = (uint8_t*)table.d;
    AirBatchDesc* descs = (AirBatchDesc*)blob.data();
    size_t at_ptr = off_ptrs, at_code = off_code, at_consts = off_consts;
    uint64_t ops = 0;
    for (size_t i = 0; i < n_items; i++) {
        const cm31_air_batch_item& it = items[i];
        descs[i].in = (const u32* const*)(dbase + at_ptr);
        memcpy(blob.data() + at_ptr, it.in_cols, it.n_in * 8);
        at_ptr += it.n_in * 8;
        descs[i].out = (u32* const*)(dbase + at_ptr);
        memcpy(blob.data() + at_ptr, it.out_cols, it.n_out * 8);
        at_ptr += it.n_out * 8;
        descs[i].code = (const uint64_t*)(dbase + at_code);
        memcpy(blob.data() + at_code, it.code, it.n_instr * 8);
        at_code += it.n_instr * 8;
        descs[i].consts = (const u32*)(dbase + at_consts);
        if (it.n_consts) memcpy(blob.data() + at_consts, it.consts, it.n_consts * 4);
        at_consts += (it.n_consts + (it.n_consts & 1)) * 4;
        descs[i].row_log = it.log_size;
        descs[i].n_instr = (u32)it.n_instr;
        descs[i].hist_bins = it.hist_bins;
        descs[i].pad = 0;
        ops += ((uint64_t)1 << it.log_size) * (it.n_in + it.n_out) * 4;
    }
    if (int e = table.fill(blob.data(), bytes)) return e;
    ProfScope prof("air_program_batch", ops);
    dim3 grid((unsigned)((((size_t)1 << max_log) + 127) / 128), (unsigned)n_items);
#define CM_AIR_BATCH(NR) air_program_batch_kernel<NR><<<grid, 128, 0, stream()>>>((const AirBatchDesc*)table.d, err)
    if (max_regs <= 64) CM_AIR_BATCH(64);
    else if (max_regs <= 128) CM_AIR_BATCH(128);
    else if (max_regs <= 256) CM_AIR_BATCH(256);
    else CM_AIR_BATCH(512);
#undef CM_AIR_BATCH
    CM_LAUNCH_CHECK();
    return 0;
}

int cm31_air_error_check(void) {
    u32* err = air_err_flag();
    CM_REQUIRE(err != nullptr, "air: cannot allocate the error word");
    u32 bits = 0;
    CM_CUDA(cudaMemcpyAsync(&bits, err, 4, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    if (bits) {
        CM_CUDA(cudaMemsetAsync(err, 0, 4, stream()));
        CM_REQUIRE(!(bits & AIR_ERR_LOOKUP_OUT_OF_TABLE), "lookup outside its table");
        CM_REQUIRE(false, "air: device error");
    }
    return 0;
}

// claimed_sum_dev: 4 device words receiving the claimed sum; nothing is synchronised.
int cm31_logup_finalize_last_async(uint32_t* const last4[4], uint32_t log_size, uint32_t* claimed_sum_dev) {
    CM_REQUIRE(log_size >= 1 && log_size <= 30, "logup_finalize_last: bad log_size");
    CM_REQUIRE(claimed_sum_dev != nullptr, "logup_finalize_last: null output");
    size_t n = (size_t)1 << log_size;
    Col4 c;
    for (int k = 0; k < 4; k++) c.p[k] = last4[k];
    const u32 n_inv = m31_inv((u32)(n % P));
    if (log_size <= FINALIZE_SMALL_LOG) {
        ProfScope prof("logup_finalize_small", 32ull * n);
        logup_finalize_small_kernel<<<1, 256, 0, stream()>>>(c, log_size, n_inv, claimed_sum_dev);
        CM_LAUNCH_CHECK();
        return 0;
    }
    unsigned long long* dsums = nullptr;
    u32* dshift = nullptr;
    CM_CUDA(cudaMallocAsync(&dsums, 32, stream()));
    CM_CUDA(cudaMallocAsync(&dshift, 16, stream()));
    CM_CUDA(cudaMemsetAsync(dsums, 0, 32, stream()));
    unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 1024);
    {
        ProfScope prof("logup_sum4", 16ull * n, 2);
        sum4_kernel<<<blocks, 256, 0, stream()>>>(c, n, dsums);
        sums_to_shift_kernel<<<1, 32, 0, stream()>>>(dsums, n_inv, claimed_sum_dev, dshift);
    }
    CM_LAUNCH_CHECK();
    static const bool chunked = getenv("CM31_CHUNKED_SCAN") != nullptr;  // the first (gathering) form, kept for A/B runs
    const bool segmented = !chunked && log_size >= SEG_Q + 6;
    u32 n_chunks = segmented ? 1u << (log_size - 1 - SEG_Q) : (u32)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
    u32* dchunks = nullptr;
    CM_CUDA(cudaMallocAsync(&dchunks, (size_t)n_chunks * 16, stream()));
    {
        ProfScope prof("logup_prefix_sum", 32ull * n, 3);
        if (segmented) {
            const u32 s = log_size - 1 - SEG_Q;
            dim3 grid((n_chunks + 255) / 256, 4);
            seg_scan_kernel<false><<<grid, 256, 0, stream()>>>(c, log_size, dshift, dchunks, s);
            scan_chunk_offsets_kernel<<<4, 1024, 0, stream()>>>(dchunks, n_chunks);
            seg_scan_kernel<true><<<grid, 256, 0, stream()>>>(c, log_size, dshift, dchunks, s);
        } else {
            scan_chunk_sums_kernel<<<dim3(n_chunks, 4), 256, 0, stream()>>>(c, log_size, dshift, dchunks, n_chunks);
            scan_chunk_offsets_kernel<<<4, 1024, 0, stream()>>>(dchunks, n_chunks);
            scan_apply_kernel<<<dim3(n_chunks, 4), 256, 0, stream()>>>(c, log_size, dshift, dchunks, n_chunks);
        }
    }
    CM_LAUNCH_CHECK();
    CM_CUDA(cudaFreeAsync(dchunks, stream()));
    CM_CUDA(cudaFreeAsync(dshift, stream()));
    CM_CUDA(cudaFreeAsync(dsums, stream()));
    return 0;
}

// the single-CTA finalize_last of MANY small components in one launch (blockIdx.x = component)
int cm31_logup_finalize_small_batch(const cm31_logup_finalize_item* items, size_t n_items) {
    if (n_items == 0) return 0;
    CM_REQUIRE(items != nullptr, "logup_finalize_small_batch: null items");
    static thread_local std::vector<LogupSmallItem> host;
    host.resize(n_items);
    uint64_t bytes = 0;
    for (size_t i = 0; i < n_items; i++) {
        CM_REQUIRE(items[i].log_size >= 1 && items[i].log_size <= FINALIZE_SMALL_LOG && items[i].claimed_sum_dev != nullptr,
                   "logup_finalize_small_batch: bad item");
        for (int k = 0; k < 4; k++) host[i].c.p[k] = items[i].last4[k];
        host[i].L = items[i].log_size;
        host[i].n_inv = m31_inv((u32)(((size_t)1 << items[i].log_size) % P));
        host[i].claimed = items[i].claimed_sum_dev;
        bytes += 32ull << items[i].log_size;
    }
    DeviceTable table;
    if (int e = table.upload(host.data(), n_items * sizeof(LogupSmallItem))) return e;
    ProfScope prof("logup_finalize_small", bytes);
    logup_finalize_small_batch_kernel<<<(unsigned)n_items, 256, 0, stream()>>>((const LogupSmallItem*)table.d);
    CM_LAUNCH_CHECK();
    return 0;
}

int cm31_logup_finalize_last(uint32_t* const last4[4], uint32_t log_size, uint32_t claimed_sum_out[4]) {
    u32* dsum = nullptr;
    CM_CUDA(cudaMallocAsync(&dsum, 16, stream()));
    if (int e = cm31_logup_finalize_last_async(last4, log_size, dsum)) return e;
    CM_CUDA(cudaMemcpyAsync(claimed_sum_out, dsum, 16, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    CM_CUDA(cudaFreeAsync(dsum, stream()));
    return 0;
}

int cm31_histogram(const uint32_t* values, size_t n, uint32_t* bins, uint32_t log_bins) {
    if (n == 0) return 0;
    ProfScope prof("histogram", 4ull * n);
    histogram_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>(values, n, bins, 1u << log_bins);
    CM_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

// Runtime side of the AOT-specialised AIR kernels (csrc/generated/*.cu, written by
// tools/gen_air_kernels.cpp from the same captured AIR graphs that produce the bytecode).
//
// A generated kernel is the straight-line CUDA form of ONE bytecode program: one thread per row,
// every SSA value in a register (nvcc allocates and schedules), column pointers and the constant /
// parameter table passed BY VALUE as a __grid_constant__ kernel argument so they are constant-bank
// operands.  It is selected at run time by the hash of the program's instruction words
// (air_code_hash); programs without a generated kernel run on the bytecode interpreter (air.cu).
#pragma once
#include "common.cuh"

namespace cm31 {

struct GenLaunch {
    const uint32_t* const* in_cols;  // host array of device pointers
    size_t n_in;
    uint32_t* const* out_cols;
    size_t n_out;
    u32 row_log, trace_log;
    const u32* consts;  // host
    size_t n_consts;
    const u32* denom_inv_dev;  // device, or null
    uint32_t* const* acc4;     // host array of 4 device pointers, or null
    int sync;                  // CTA-wide barriers inside the generated bodies (instruction-fetch sharing), see GEN_SYNC
    u32 hist_bins;             // size of the bin column OP_HIST counts into (0 for programs without histograms)
    u32* err_flag;             // device word: AIR_ERR_* bits (cm31_air_error_check)
};
enum : u32 { AIR_ERR_LOOKUP_OUT_OF_TABLE = 1 };

struct GenEntry {
    uint64_t hash;
    const char* name;
    int (*launch)(const GenLaunch&);
};

// generated/air_registry.cu
const GenEntry* air_gen_lookup(uint64_t hash);

__device__ __forceinline__ u32 gen_offset_row(u32 row, u32 trace_log, u32 eval_log, int off) {
    // core/utils.rs:74-90 offset_bit_reversed_circle_domain_index
    u32 idx = bit_reverse(row, eval_log);
    u32 half = 1u << (eval_log - 1);
    int step = off * (int)(1u << (eval_log - trace_log - 1));
    u32 mask = half - 1;
    if (idx < half) idx = (u32)((int)idx + step) & mask;
    else idx = ((u32)((int)(idx - half) - step) & mask) + half;
    return bit_reverse(idx, eval_log);
}

__device__ __forceinline__ void gen_hist(u32* bins, u32 v, u32 n_bins, u32* err) {
    // A looked-up value comes from witness data (clock deltas, u32 limbs, bitwise tuples): outside the table it must not
    // index the bin column -- the reference panics on the slice index (range_check_macro.rs:72-84), here the proof fails
    // with "lookup outside its table" (cm31_air_error_check).
    if (v >= n_bins) {
        atomicOr(err, AIR_ERR_LOOKUP_OUT_OF_TABLE);
        return;
    }
    // equal values across the warp cost one atomic (lookups of a loop-shaped trace hit few bins)
    const unsigned peers = __match_any_sync(__activemask(), v);
    if ((threadIdx.x & 31u) == (u32)(__ffs(peers) - 1)) atomicAdd(bins + v, (u32)__popc(peers));
}

// ---- lazily reduced arithmetic of the generated kernels.  Canonical inputs (< P < 2^31) give
// products < 2^62, so four of them fit a u64 accumulator (4*P*(P-1) < 2^64); dot_fold brings an
// accumulator back below 2^34 so the next four fit again.  P - f (in [1, P]) stands for -f.
__device__ __forceinline__ u64 g_fold64(u64 x) { return (x & P) + (x >> 31); }
__device__ __forceinline__ void dot_term(u64& d0, u64& d1, u64& d2, u64& d3, QM31 e, u32 f) {
    d0 += (u64)e.a * f;
    d1 += (u64)e.b * f;
    d2 += (u64)e.c * f;
    d3 += (u64)e.d * f;
}
__device__ __forceinline__ void dot_fold(u64& d0, u64& d1, u64& d2, u64& d3) {
    d0 = g_fold64(d0);
    d1 = g_fold64(d1);
    d2 = g_fold64(d2);
    d3 = g_fold64(d3);
}
__device__ __forceinline__ QM31 dot_done(u64 d0, u64 d1, u64 d2, u64 d3) {
    return qm_make(m31_reduce64(d0), m31_reduce64(d1), m31_reduce64(d2), m31_reduce64(d3));
}
// QM31 product as six u64 dot products (16 multiply-adds, 6 reductions) instead of four CM31
// products with a reduction per partial product.  (A + Bu)(C + Du) = (AC + (2+i)BD) + (AD + BC)u.
__device__ __forceinline__ QM31 g_qm_mul(QM31 x, QM31 y) {
    const u32 nyb = P - y.b, nyd = P - y.d;
    const u32 p = m31_reduce64((u64)x.c * y.c + (u64)x.d * nyd);  // Re(BD)
    const u32 q = m31_reduce64((u64)x.c * y.d + (u64)x.d * y.c);  // Im(BD)
    const u64 lre = (u64)x.a * y.a + (u64)x.b * nyb + 2ull * p + (P - q);
    const u64 lim = (u64)x.a * y.b + (u64)x.b * y.a + p + 2ull * q;
    const u64 hre = (u64)x.a * y.c + (u64)x.b * nyd + (u64)x.c * y.a + (u64)x.d * nyb;
    const u64 him = (u64)x.a * y.d + (u64)x.b * y.c + (u64)x.c * y.b + (u64)x.d * y.a;
    return qm_make(m31_reduce64(lre), m31_reduce64(lim), m31_reduce64(hre), m31_reduce64(him));
}

}  // namespace cm31

// ---- vocabulary of the generated kernel bodies (a = the kernel's argument struct)
using cm31::dot_term;
using cm31::dot_fold;
using cm31::dot_done;
using cm31::g_qm_mul;
#define ldcol(i) __ldg(a.in[i] + row)
#define ldcol_off(i, off) __ldg(a.in[i] + cm31::gen_offset_row(row, a.trace_log, a.row_log, (off)))
#define cw(s) (a.c[s])
#define cq(s) cm31::qm_make(a.c[s], a.c[(s) + 1], a.c[(s) + 2], a.c[(s) + 3])
// `live`: the thread's row exists (every thread of a CTA runs the whole body so the CTA-wide barriers of GEN_SYNC are legal;
// rows past the end compute on row 0 and store nothing)
#define st1(slot, f)                      \
    do {                                  \
        if (live) a.out[slot][row] = (f); \
    } while (0)
#define st4(slot, e)                        \
    do {                                    \
        if (live) {                         \
            a.out[slot][row] = (e).a;       \
            a.out[(slot) + 1][row] = (e).b; \
            a.out[(slot) + 2][row] = (e).c; \
            a.out[(slot) + 3][row] = (e).d; \
        }                                   \
    } while (0)
#define hist(slot, f)                                                     \
    do {                                                                  \
        if (live) cm31::gen_hist(a.out[slot], (f), a.hist_bins, a.err);   \
    } while (0)
// Instruction fetch is what these 2-7k-instruction straight-line bodies are short of (ncu r01b: `no_inst` 29-32 % of the
// stall samples): every warp streams the whole body from L2 unless another warp of its SM fetched the same lines a moment
// ago.  A CTA-wide barrier every few dozen statements keeps the warps of a CTA within the 32 KB L1.5 instruction cache of
// each other, so a line is fetched once per CTA instead of once per warp.  a.sync is uniform (kernel argument).
#define GEN_SYNC()                    \
    do {                              \
        if (a.sync) __syncthreads();  \
    } while (0)
// lazily reduced accumulation of  coeff (QM31) x f  over the base-field constraints of a component
#define CACC(e, f)                                  \
    do {                                            \
        dot_term(ca0, ca1, ca2, ca3, (e), (f));     \
        if ((++ca_n & 3) == 0) dot_fold(ca0, ca1, ca2, ca3); \
    } while (0)

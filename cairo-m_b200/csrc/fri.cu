// FriOps, AccumulationOps and QuotientOps on sm_100a.
//
// Replaces external/stwo/crates/prover/src/core/backend/simd/fri.rs:24-165,
// simd/accumulation.rs:11-39, simd/quotients.rs:33-265; defined by core/fri.rs:1132-1189,
// cpu/fri.rs:29-85, cpu/accumulation.rs:8-25, cpu/quotients.rs:18-146.
// All kernels are one-thread-per-output-row streaming kernels over SoA QM31 columns.
#include "circle.hpp"
#include "common.cuh"

namespace cm31 {

struct Ptr4 {
    u32* p[4];
};
struct CPtr4 {
    const u32* p[4];
};
__device__ __forceinline__ QM31 ld4(const CPtr4& c, size_t i) {
    return qm_make(__ldg(c.p[0] + i), __ldg(c.p[1] + i), __ldg(c.p[2] + i), __ldg(c.p[3] + i));
}
__device__ __forceinline__ QM31 ld4(const Ptr4& c, size_t i) { return qm_make(c.p[0][i], c.p[1][i], c.p[2][i], c.p[3][i]); }
__device__ __forceinline__ void st4(const Ptr4& c, size_t i, QM31 v) {
    c.p[0][i] = v.a;
    c.p[1][i] = v.b;
    c.p[2][i] = v.c;
    c.p[3][i] = v.d;
}

// fold_line (fri.rs:1132-1157): pair (2i, 2i+1), twiddle = 1/x of domain.at(bit_reverse(2i)).
// itw_layer = inverse twiddles of the tree level whose coset is the line domain's coset.
__global__ void fold_line_kernel(CPtr4 src, Ptr4 dst, size_t n_out, QM31 alpha, const u32* __restrict__ itw_layer) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    QM31 f0 = ld4(src, 2 * i), f1 = ld4(src, 2 * i + 1);
    u32 it = __ldg(itw_layer + i);
    QM31 s = qm_add(f0, f1);
    QM31 d = qm_mul_m31(qm_sub(f0, f1), it);
    st4(dst, i, qm_add(s, qm_mul(alpha, d)));
}

// fold_circle_into_line (fri.rs:1159-1189): twiddle = 1/y of domain.at(bit_reverse(2i)), derived
// from the first line layer pairs [x, y] -> [y, -y, -x, x]  (cpu/circle.rs:209-229).
__global__ void fold_circle_kernel(Ptr4 dst, CPtr4 src, size_t n_out, QM31 alpha, QM31 alpha_sq,
                                   const u32* __restrict__ itw_line0) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    QM31 f0 = ld4(src, 2 * i), f1 = ld4(src, 2 * i + 1);
    u32 x = __ldg(itw_line0 + 2 * (i >> 2)), y = __ldg(itw_line0 + 2 * (i >> 2) + 1);
    u32 sel = (u32)i & 3;
    u32 v = sel < 2 ? y : x;
    u32 it = (sel == 1 || sel == 2) ? m31_neg(v) : v;
    QM31 s = qm_add(f0, f1);
    QM31 d = qm_mul_m31(qm_sub(f0, f1), it);
    QM31 fp = qm_add(qm_mul(alpha, d), s);
    st4(dst, i, qm_add(qm_mul(ld4(dst, i), alpha_sq), fp));
}

// circle domains of log size <= 2 have no [x,y] twiddle pair in the tree: explicit 1/y values.
__global__ void fold_circle_small_kernel(Ptr4 dst, CPtr4 src, u32 n_out, QM31 alpha, QM31 alpha_sq, u32 iy0, u32 iy1) {
    u32 i = threadIdx.x;
    if (i >= n_out) return;
    QM31 f0 = ld4(src, 2 * i), f1 = ld4(src, 2 * i + 1);
    u32 it = i == 0 ? iy0 : iy1;
    QM31 s = qm_add(f0, f1);
    QM31 d = qm_mul_m31(qm_sub(f0, f1), it);
    QM31 fp = qm_add(qm_mul(alpha, d), s);
    st4(dst, i, qm_add(qm_mul(ld4(dst, i), alpha_sq), fp));
}

__global__ void accumulate_kernel(Ptr4 dst, CPtr4 src, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 4; k++) dst.p[k][i] = m31_add(dst.p[k][i], __ldg(src.p[k] + i));
}

// decompose (cpu/fri.rs:29-85): lambda = (sum first half - sum second half) / n; g = f -/+ lambda.
__global__ void decompose_sum_kernel(CPtr4 src, size_t n, unsigned long long* sums /* 8: lo[4], hi[4] */) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool hi = i >= n / 2;
#pragma unroll
    for (int k = 0; k < 4; k++) atomicAdd(&sums[(hi ? 4 : 0) + k], (unsigned long long)__ldg(src.p[k] + i));
}
__global__ void decompose_apply_kernel(CPtr4 src, Ptr4 dst, size_t n, QM31 lambda) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    QM31 v = ld4(src, i);
    st4(dst, i, i < n / 2 ? qm_sub(v, lambda) : qm_add(v, lambda));
}

// ------------------------------------------------------------------ DEEP quotients
struct QuotBatch {
    u32 prx[2], pry[2], pix[2], piy[2];  // point = (prx + u*pix, pry + u*piy), each CM31
    u32 batch_coeff[4];                  // alpha^(#columns in batch)
    u32 sum_a[4], sum_b[4];              // sum_i alpha^i a_i, sum_i alpha^i b_i
    u32 start, end;                      // range into the (col_idx, c) tables
};

__device__ __forceinline__ CirclePointM31 domain_point_bitrev(u32 row, u32 log_size, u32 half_initial, u32 half_step,
                                                               const CirclePointM31* __restrict__ gen_pow) {
    // CircleDomain::at(bit_reverse(row))  (poly/circle/domain.rs:57-63)
    u32 d = bit_reverse(row, log_size);
    u32 half = 1u << (log_size - 1);
    bool conj = d >= half;
    u32 dd = conj ? d - half : d;
    u32 idx = (u32)((half_initial + (u64)half_step * dd) & 0x7fffffffu);
    if (conj) idx = (u32)(((1ull << 31) - idx) & 0x7fffffffu);
    CirclePointM31 res = {1, 0};
#pragma unroll 1
    for (u32 bit = 0; bit < 31; bit++) {
        if (idx & (1u << bit)) res = cp_add(res, gen_pow[bit]);
    }
    return res;
}

// accumulate_row_quotients (cpu/quotients.rs:45-78) with the per-batch sums of the line
// coefficients a_i, b_i hoisted out of the column loop:
//   num = sum_i c_i * f_i(row)  -  (A * y + B),   row = row*alpha^n_b + num / den.
__global__ void __launch_bounds__(256) quotients_kernel(u32 log_size, const u32* const* __restrict__ cols,
                                                        const QuotBatch* __restrict__ batches, u32 n_batches,
                                                        const u32* __restrict__ col_idx, const u32* __restrict__ coef_c,
                                                        u32 half_initial, u32 half_step,
                                                        const CirclePointM31* __restrict__ gen_pow, Ptr4 out) {
    size_t row = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (row >= ((size_t)1 << log_size)) return;
    CirclePointM31 p = domain_point_bitrev((u32)row, log_size, half_initial, half_step, gen_pow);
    QM31 acc = qm_zero();
    for (u32 b = 0; b < n_batches; b++) {
        const QuotBatch qb = batches[b];
        QM31 num = qm_zero();
        for (u32 k = qb.start; k < qb.end; k++) {
            u32 f = __ldg(cols[__ldg(col_idx + k)] + row);
            const u32* c = coef_c + (size_t)k * 4;
            num = qm_add(num, qm_mul_m31(qm_make(__ldg(c), __ldg(c + 1), __ldg(c + 2), __ldg(c + 3)), f));
        }
        QM31 A = qm_make(qb.sum_a[0], qb.sum_a[1], qb.sum_a[2], qb.sum_a[3]);
        QM31 B = qm_make(qb.sum_b[0], qb.sum_b[1], qb.sum_b[2], qb.sum_b[3]);
        num = qm_sub(num, qm_add(qm_mul_m31(A, p.y), B));
        // den = (prx - x) * piy - (pry - y) * pix      (cpu/quotients.rs:127-139)
        CM31 prx = cm_make(qb.prx[0], qb.prx[1]), pry = cm_make(qb.pry[0], qb.pry[1]);
        CM31 pix = cm_make(qb.pix[0], qb.pix[1]), piy = cm_make(qb.piy[0], qb.piy[1]);
        CM31 den = cm_sub(cm_mul(cm_make(m31_sub(prx.a, p.x), prx.b), piy),
                          cm_mul(cm_make(m31_sub(pry.a, p.y), pry.b), pix));
        CM31 di = cm_inv(den);
        QM31 bc = qm_make(qb.batch_coeff[0], qb.batch_coeff[1], qb.batch_coeff[2], qb.batch_coeff[3]);
        acc = qm_add(qm_mul(acc, bc), qm_mul_cm31(num, di));
    }
    st4(out, row, acc);
}

// ---- fast path (log_size >= 9): one CTA = 256 consecutive rows.
// * Domain point: row = blk*256 + t  =>  bit_reverse(row) = bit_reverse(t,8) << (L-8) | bit_reverse(blk),
//   so point(row) = Q_blk + R[t >> 1], conjugated when t is odd: one 31-step double-and-add per CTA
//   (thread 0) and one circle-group addition per row instead of a 31-step loop per row.
// * Numerator: sum_i c_i * f_i(row) accumulated as four u64 dot products (IMAD.WIDE with carry-in),
//   folded to 34 bits every 4 columns and reduced mod P once per batch: 4 multiply-adds per column
//   instead of 4 reduced multiplications + 4 reduced additions.
struct QuotEntry {  // one (batch, column) term, 32 bytes = two 128-bit uniform loads
    const u32* col;
    u32 pad[2];
    u32 c[4];
};
struct QuotPointTable {
    CirclePointM31 r[128];
};
__device__ __forceinline__ u64 fold64(u64 x) { return (x & P) + (x >> 31); }

// one thread per CTA of the main kernel: Q_blk = point(half_initial + half_step * bit_reverse(blk))
__global__ void quotients_block_points_kernel(u32 log_size, u32 half_initial, u32 half_step, const CirclePointM31* __restrict__ gen_pow,
                                              CirclePointM31* __restrict__ q_out) {
    u32 blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= (1u << (log_size - 8))) return;
    u32 bk = bit_reverse(blk, log_size - 8);
    u32 idx = (u32)((half_initial + (u64)half_step * bk) & 0x7fffffffu);
    CirclePointM31 res = {1, 0};
#pragma unroll 1
    for (u32 bit = 0; bit < 31; bit++)
        if (idx & (1u << bit)) res = cp_add(res, gen_pow[bit]);
    q_out[blk] = res;
}

constexpr u32 QUOT_TILE = 256;  // (batch, column) terms staged in shared memory at a time

__global__ void __launch_bounds__(256) quotients_fast_kernel(u32 log_size, const QuotEntry* __restrict__ entries,
                                                             const QuotBatch* __restrict__ batches, u32 n_batches,
                                                             const CirclePointM31* __restrict__ q_blk,
                                                             const __grid_constant__ QuotPointTable table, Ptr4 out, u32 blk0) {
    __shared__ uint4 sh_entries[QUOT_TILE * 2];
    const u32 t = threadIdx.x;
    const u32 blk = blockIdx.x + blk0;  // blk0: first 256-row block of this launch (a rank's row range when a proof is sharded)
    const size_t row = blk * (size_t)256 + t;
    CirclePointM31 p = cp_add(q_blk[blk], table.r[t >> 1]);
    if (t & 1) p.y = m31_neg(p.y);
    QM31 acc = qm_zero();
    for (u32 b = 0; b < n_batches; b++) {
        const QuotBatch qb = batches[b];
        u64 n0 = 0, n1 = 0, n2 = 0, n3 = 0;
        for (u32 k0 = qb.start; k0 < qb.end; k0 += QUOT_TILE) {
            const u32 cnt = min(QUOT_TILE, qb.end - k0);
            __syncthreads();  // previous tile fully consumed
            for (u32 i = t; i < cnt * 2; i += 256) sh_entries[i] = __ldg(reinterpret_cast<const uint4*>(entries + k0) + i);
            __syncthreads();
            u32 k = 0;
            for (; k + 8 <= cnt; k += 8) {  // 8 independent column reads in flight (ncu r02: long_scoreboard 31 % with 4)
                u32 f[8];
#pragma unroll
                for (u32 j = 0; j < 8; j++) {
                    const uint4 e0 = sh_entries[2 * (k + j)];
                    f[j] = __ldg(reinterpret_cast<const u32*>(((u64)e0.y << 32) | e0.x) + row);
                }
#pragma unroll
                for (u32 h = 0; h < 2; h++) {  // 4 products of < 2^62 fit a u64 on top of a folded carry
#pragma unroll
                    for (u32 j = 4 * h; j < 4 * h + 4; j++) {
                        const uint4 c = sh_entries[2 * (k + j) + 1];
                        n0 += (u64)f[j] * c.x;
                        n1 += (u64)f[j] * c.y;
                        n2 += (u64)f[j] * c.z;
                        n3 += (u64)f[j] * c.w;
                    }
                    n0 = fold64(n0);
                    n1 = fold64(n1);
                    n2 = fold64(n2);
                    n3 = fold64(n3);
                }
            }
            for (; k + 4 <= cnt; k += 4) {
                u32 f[4];
#pragma unroll
                for (u32 j = 0; j < 4; j++) {  // the 4 column reads are independent: all in flight together
                    const uint4 e0 = sh_entries[2 * (k + j)];
                    f[j] = __ldg(reinterpret_cast<const u32*>(((u64)e0.y << 32) | e0.x) + row);
                }
#pragma unroll
                for (u32 j = 0; j < 4; j++) {
                    const uint4 c = sh_entries[2 * (k + j) + 1];
                    n0 += (u64)f[j] * c.x;
                    n1 += (u64)f[j] * c.y;
                    n2 += (u64)f[j] * c.z;
                    n3 += (u64)f[j] * c.w;
                }
                n0 = fold64(n0);
                n1 = fold64(n1);
                n2 = fold64(n2);
                n3 = fold64(n3);
            }
            for (; k < cnt; k++) {  // <= 3 leftover terms: below 2^64 on top of a folded carry
                const uint4 e0 = sh_entries[2 * k];
                const uint4 c = sh_entries[2 * k + 1];
                const u64 f = __ldg(reinterpret_cast<const u32*>(((u64)e0.y << 32) | e0.x) + row);
                n0 += f * c.x;
                n1 += f * c.y;
                n2 += f * c.z;
                n3 += f * c.w;
            }
            n0 = fold64(n0);
            n1 = fold64(n1);
            n2 = fold64(n2);
            n3 = fold64(n3);
        }
        QM31 num = qm_make(m31_reduce64(n0), m31_reduce64(n1), m31_reduce64(n2), m31_reduce64(n3));
        QM31 A = qm_make(qb.sum_a[0], qb.sum_a[1], qb.sum_a[2], qb.sum_a[3]);
        QM31 B = qm_make(qb.sum_b[0], qb.sum_b[1], qb.sum_b[2], qb.sum_b[3]);
        num = qm_sub(num, qm_add(qm_mul_m31(A, p.y), B));
        CM31 prx = cm_make(qb.prx[0], qb.prx[1]), pry = cm_make(qb.pry[0], qb.pry[1]);
        CM31 pix = cm_make(qb.pix[0], qb.pix[1]), piy = cm_make(qb.piy[0], qb.piy[1]);
        CM31 den = cm_sub(cm_mul(cm_make(m31_sub(prx.a, p.x), prx.b), piy),
                          cm_mul(cm_make(m31_sub(pry.a, p.y), pry.b), pix));
        CM31 di = cm_inv(den);
        QM31 bc = qm_make(qb.batch_coeff[0], qb.batch_coeff[1], qb.batch_coeff[2], qb.batch_coeff[3]);
        acc = qm_add(qm_mul(acc, bc), qm_mul_cm31(num, di));
    }
    st4(out, row, acc);
}

static const u32* tree_level_for_line_coset(const cm31_twiddles* tw, bool inverse, u32 coset_log_size) {
    // level whose coset is half_odds(coset_log_size): offset 2^(M-1) - 2^coset_log_size
    const u32* base = inverse ? tw->itw : tw->tw;
    return base + (((size_t)1 << (tw->log_size - 1)) - ((size_t)1 << coset_log_size));
}

}  // namespace cm31

using namespace cm31;

static inline QM31 qm_from_arr(const uint32_t v[4]) { return qm_make(v[0], v[1], v[2], v[3]); }

extern "C" {

int cm31_fold_line(const uint32_t* const src4[4], uint32_t log_size, const uint32_t alpha[4], const cm31_twiddles* tw,
                   uint32_t* const dst4[4]) {
    CM_REQUIRE(log_size >= 1, "fold_line: evaluation too small");
    CM_REQUIRE(tw && log_size + 1 <= tw->log_size, "fold_line: twiddle tree too small");
    CPtr4 s;
    Ptr4 d;
    for (int k = 0; k < 4; k++) {
        s.p[k] = src4[k];
        d.p[k] = dst4[k];
    }
    size_t n_out = (size_t)1 << (log_size - 1);
    const u32* itw = tree_level_for_line_coset(tw, true, log_size);
    ProfScope prof("fold_line", 24ull * (n_out * 2));
    prof_ops(n_out * (4 + 4 + 4 + 31 + 4));  // per output: ibutterfly on QM31 (add, sub, x M31) + alpha * f1 (QM31 mul) + add
    fold_line_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, stream()>>>(s, d, n_out, qm_from_arr(alpha), itw);
    CM_LAUNCH_CHECK();
    return 0;
}

int cm31_fold_circle_into_line(uint32_t* const dst4[4], const uint32_t* const src4[4], uint32_t log_size,
                               const uint32_t alpha[4], const cm31_twiddles* tw) {
    CM_REQUIRE(log_size >= 1, "fold_circle_into_line: evaluation too small");
    CPtr4 s;
    Ptr4 d;
    for (int k = 0; k < 4; k++) {
        s.p[k] = src4[k];
        d.p[k] = dst4[k];
    }
    size_t n_out = (size_t)1 << (log_size - 1);
    QM31 a = qm_from_arr(alpha);
    QM31 a2 = qm_sqr(a);
    ProfScope prof("fold_circle_into_line", 32ull * (n_out * 2));
    prof_ops(n_out * (4 + 4 + 4 + 31 + 4 + 31 + 4));  // as fold_line + dst * alpha^2 + f'
    if (log_size <= 2) {
        CircleDomain dom = CanonicCoset(log_size).circle_domain();
        u32 iy0 = m31_inv(dom.at(bit_reverse(0, log_size)).y);
        u32 iy1 = log_size == 2 ? m31_inv(dom.at(bit_reverse(2, log_size)).y) : 0;
        fold_circle_small_kernel<<<1, 32, 0, stream()>>>(d, s, (u32)n_out, a, a2, iy0, iy1);
    } else {
        CM_REQUIRE(tw && log_size <= tw->log_size, "fold_circle_into_line: twiddle tree too small");
        const u32* itw0 = tree_level_for_line_coset(tw, true, log_size - 1);
        fold_circle_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, stream()>>>(d, s, n_out, a, a2, itw0);
    }
    CM_LAUNCH_CHECK();
    return 0;
}

int cm31_accumulate(uint32_t* const dst4[4], const uint32_t* const src4[4], size_t n) {
    CPtr4 s;
    Ptr4 d;
    for (int k = 0; k < 4; k++) {
        s.p[k] = src4[k];
        d.p[k] = dst4[k];
    }
    if (n == 0) return 0;
    ProfScope prof("accumulate", 48ull * n);
    prof_ops(4ull * n);
    accumulate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>(d, s, n);
    CM_LAUNCH_CHECK();
    return 0;
}

int cm31_secure_powers(const uint32_t felt[4], size_t n_powers, uint32_t* out_host) {
    // sequential by definition (cpu/accumulation.rs:17-25); n is the constraint count (hundreds)
    QM31 f = qm_from_arr(felt), acc = qm_one();
    for (size_t i = 0; i < n_powers; i++) {
        out_host[4 * i] = acc.a;
        out_host[4 * i + 1] = acc.b;
        out_host[4 * i + 2] = acc.c;
        out_host[4 * i + 3] = acc.d;
        acc = qm_mul(acc, f);
    }
    return 0;
}

int cm31_decompose(const uint32_t* const src4[4], uint32_t log_size, uint32_t* const dst4[4], uint32_t lambda_out[4]) {
    CM_REQUIRE(log_size >= 1, "decompose: evaluation too small");
    CPtr4 s;
    Ptr4 d;
    for (int k = 0; k < 4; k++) {
        s.p[k] = src4[k];
        d.p[k] = dst4[k];
    }
    size_t n = (size_t)1 << log_size;
    unsigned long long* dsums = nullptr;
    CM_CUDA(cudaMallocAsync(&dsums, 64, stream()));
    CM_CUDA(cudaMemsetAsync(dsums, 0, 64, stream()));
    {
        ProfScope prof("decompose_sum", 16ull * n);
        decompose_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>(s, n, dsums);
    }
    CM_LAUNCH_CHECK();
    unsigned long long h[8];
    CM_CUDA(cudaMemcpyAsync(h, dsums, 64, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    CM_CUDA(cudaFreeAsync(dsums, stream()));
    u32 lo[4], hi[4];
    for (int k = 0; k < 4; k++) {
        lo[k] = m31_reduce64(h[k]);
        hi[k] = m31_reduce64(h[4 + k]);
    }
    u32 n_inv = m31_inv((u32)(n % P));
    QM31 lam = qm_mul_m31(qm_sub(qm_make(lo[0], lo[1], lo[2], lo[3]), qm_make(hi[0], hi[1], hi[2], hi[3])), n_inv);
    lambda_out[0] = lam.a;
    lambda_out[1] = lam.b;
    lambda_out[2] = lam.c;
    lambda_out[3] = lam.d;
    ProfScope prof("decompose_apply", 32ull * n);
    decompose_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>(s, d, n, lam);
    CM_LAUNCH_CHECK();
    return 0;
}

int cm31_accumulate_quotients(uint32_t log_size, const uint32_t* const* cols, size_t n_cols,
                              const uint32_t random_coeff[4], size_t n_batches, const uint32_t* batch_points_host,
                              const uint32_t* batch_start_host, const uint32_t* col_idx_host,
                              const uint32_t* values_host, uint32_t* const out4[4]) {
    return cm31_accumulate_quotients_range(log_size, cols, n_cols, random_coeff, n_batches, batch_points_host, batch_start_host, col_idx_host,
                                           values_host, out4, 0, (size_t)1 << log_size);
}

int cm31_accumulate_quotients_range(uint32_t log_size, const uint32_t* const* cols, size_t n_cols,
                                    const uint32_t random_coeff[4], size_t n_batches, const uint32_t* batch_points_host,
                                    const uint32_t* batch_start_host, const uint32_t* col_idx_host,
                                    const uint32_t* values_host, uint32_t* const out4[4], size_t first_row, size_t n_rows) {
    return cm31_accumulate_quotients_partial(log_size, cols, n_cols, random_coeff, n_batches, batch_points_host, batch_start_host, col_idx_host,
                                             values_host, out4, first_row, n_rows, nullptr);
}

// entry_active (one byte per (batch, column) entry, or NULL = all): the quotient is LINEAR in the per-column terms
// (c * f(row) - (a * y + b)) / den, so the entries can be split over the ranks of a sharded proof -- every rank accumulates
// the terms of the columns it owns over all rows, reading only local memory, and the partial quotients are summed mod P
// (cm31_shard_reduce_m31).  Inactive entries still take their power of the random coefficient.
int cm31_accumulate_quotients_partial(uint32_t log_size, const uint32_t* const* cols, size_t n_cols,
                                      const uint32_t random_coeff[4], size_t n_batches, const uint32_t* batch_points_host,
                                      const uint32_t* batch_start_host, const uint32_t* col_idx_host,
                                      const uint32_t* values_host, uint32_t* const out4[4], size_t first_row, size_t n_rows,
                                      const uint8_t* entry_active) {
    CM_REQUIRE(log_size >= 1 && log_size <= 30, "accumulate_quotients: bad log_size");
    const bool whole = first_row == 0 && n_rows == ((size_t)1 << log_size);
    CM_REQUIRE(whole || (log_size >= 9 && first_row % 256 == 0 && n_rows % 256 == 0 && first_row + n_rows <= ((size_t)1 << log_size)),
               "accumulate_quotients: a row range must be made of whole 256-row blocks of a domain of at least 2^9 rows");
    QM31 alpha = qm_from_arr(random_coeff);
    std::vector<QuotBatch> qb(n_batches);
    std::vector<u32> coef_c, act_idx;  // the ACTIVE entries, compacted: coefficient (4 words) and column index per entry
    coef_c.reserve((size_t)batch_start_host[n_batches] * 4 + 4);
    for (size_t b = 0; b < n_batches; b++) {
        const u32* pt = batch_points_host + b * 8;
        QM31 px = qm_make(pt[0], pt[1], pt[2], pt[3]), py = qm_make(pt[4], pt[5], pt[6], pt[7]);
        CM_REQUIRE(!(py.c == 0 && py.d == 0), "accumulate_quotients: sample point on the conjugate line");
        QuotBatch& q = qb[b];
        q.prx[0] = px.a; q.prx[1] = px.b; q.pix[0] = px.c; q.pix[1] = px.d;
        q.pry[0] = py.a; q.pry[1] = py.b; q.piy[0] = py.c; q.piy[1] = py.d;
        q.start = (u32)act_idx.size();
        // column_line_coeffs (cpu/quotients.rs:84-108) + complex_conjugate_line_coeffs (constraints.rs:98-113)
        QM31 al = qm_one(), sa = qm_zero(), sb = qm_zero();
        QM31 c = qm_sub(qm_conj(py), py);
        for (u32 k = batch_start_host[b]; k < batch_start_host[b + 1]; k++) {
            CM_REQUIRE(col_idx_host[k] < n_cols, "accumulate_quotients: column index out of range");
            al = qm_mul(al, alpha);  // every entry takes its power, active or not
            if (entry_active && !entry_active[k]) continue;
            QM31 v = qm_from_arr(values_host + (size_t)k * 4);
            QM31 a = qm_sub(qm_conj(v), v);
            QM31 bq = qm_sub(qm_mul(v, c), qm_mul(a, py));
            sa = qm_add(sa, qm_mul(al, a));
            sb = qm_add(sb, qm_mul(al, bq));
            QM31 ac = qm_mul(al, c);
            for (u32 w : {ac.a, ac.b, ac.c, ac.d}) coef_c.push_back(w);
            act_idx.push_back(col_idx_host[k]);
        }
        q.end = (u32)act_idx.size();
        QM31 bc = qm_pow(alpha, batch_start_host[b + 1] - batch_start_host[b]);
        q.batch_coeff[0] = bc.a; q.batch_coeff[1] = bc.b; q.batch_coeff[2] = bc.c; q.batch_coeff[3] = bc.d;
        q.sum_a[0] = sa.a; q.sum_a[1] = sa.b; q.sum_a[2] = sa.c; q.sum_a[3] = sa.d;
        q.sum_b[0] = sb.a; q.sum_b[1] = sb.b; q.sum_b[2] = sb.c; q.sum_b[3] = sb.d;
    }
    const size_t n_entries = act_idx.size();
    const u32* col_idx_act = act_idx.data();
    for (int k = 0; k < 4; k++) coef_c.push_back(0);
    std::vector<CirclePointM31> gen_pow(31);
    CirclePointM31 g = {M31_CIRCLE_GEN_X, M31_CIRCLE_GEN_Y};
    for (int i = 0; i < 31; i++) {
        gen_pow[i] = g;
        g = cp_double(g);
    }
    DeviceTable dcols, dqb, didx, dc, dgen;
    if (int e = dcols.upload(cols, n_cols * sizeof(void*))) return e;
    if (int e = dqb.upload(qb.data(), qb.size() * sizeof(QuotBatch))) return e;
    if (int e = didx.upload(col_idx_act, n_entries * 4)) return e;
    if (int e = dc.upload(coef_c.data(), coef_c.size() * 4)) return e;
    if (int e = dgen.upload(gen_pow.data(), gen_pow.size() * sizeof(CirclePointM31))) return e;
    Ptr4 o;
    for (int k = 0; k < 4; k++) o.p[k] = out4[k];
    Coset half = CanonicCoset(log_size).half_coset();
    size_t n = (size_t)1 << log_size;
    if (log_size >= 9) {
        std::vector<QuotEntry> entries(n_entries + 1);
        for (size_t k = 0; k < n_entries; k++) {
            entries[k].col = cols[col_idx_act[k]];
            entries[k].pad[0] = entries[k].pad[1] = 0;
            for (int j = 0; j < 4; j++) entries[k].c[j] = coef_c[k * 4 + j];
        }
        QuotPointTable table;
        for (u32 j = 0; j < 128; j++) {
            u64 mult = (u64)bit_reverse(j, 7) << (log_size - 8);
            table.r[j] = cp_from_index((u32)(((u64)half.step_size * mult) & 0x7fffffffu));
        }
        DeviceTable dent;
        if (int e = dent.upload(entries.data(), entries.size() * sizeof(QuotEntry))) return e;
        CirclePointM31* dq = nullptr;
        CM_CUDA(cudaMallocAsync(&dq, (n / 256) * sizeof(CirclePointM31), stream()));
        ProfScope prof("accumulate_quotients", (4ull * n_cols + 16ull) * n_rows, 2);
        // per row: every column term c * f(row) is QM31 x M31 + QM31 add (8); per batch: line value (8), acc * alpha^k (31),
        // CM31 denominator + inverse share (~12), numerator x 1/den (QM31 x CM31 = 8 mul + 4 add)
        prof_ops(n_rows * (8ull * entries.size() + (8 + 31 + 12 + 12) * (uint64_t)n_batches));
        quotients_block_points_kernel<<<(unsigned)((n / 256 + 127) / 128), 128, 0, stream()>>>(log_size, half.initial_index, half.step_size,
                                                                                             (const CirclePointM31*)dgen.d, dq);
        quotients_fast_kernel<<<(unsigned)(n_rows / 256), 256, 0, stream()>>>(log_size, (const QuotEntry*)dent.d, (const QuotBatch*)dqb.d,
                                                                            (u32)n_batches, dq, table, o, (u32)(first_row / 256));
        CM_LAUNCH_CHECK();
        CM_CUDA(cudaFreeAsync(dq, stream()));
        return 0;
    }
    ProfScope prof("accumulate_quotients", (4ull * n_cols + 16ull) * n);
    quotients_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>(
        log_size, (const u32* const*)dcols.d, (const QuotBatch*)dqb.d, (u32)n_batches, (const u32*)didx.d,
        (const u32*)dc.d, half.initial_index, half.step_size, (const CirclePointM31*)dgen.d, o);
    CM_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

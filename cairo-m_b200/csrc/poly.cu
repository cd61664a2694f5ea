// PolyOps on sm_100a: twiddle tree, batched circle (I)FFT, LDE, bit reversal, eval_at_point.
//
// Replaces the SimdBackend kernels
//   external/stwo/crates/prover/src/core/backend/simd/fft/ifft.rs:33-57, rfft.rs:35-60,
//   simd/circle.rs:132-297,303-341, simd/bit_reverse.rs:48-147
// and is defined by (bit-exact against) the CpuBackend
//   external/stwo/crates/prover/src/core/backend/cpu/circle.rs:18-229, poly/utils.rs:44-99.
//
// Layout: one column = 2^L packed u32 M31 words, contiguous (column-major trace).  The FFT is
// radix-2 over layers 0 (circle layer) .. L-1; a "pass" loads a tile into shared memory,
// performs up to 12 layers (radix-8 register rounds) and stores it back, so a 2^22 column is
// transformed with two HBM round trips.  Twiddles use the reference CPU tree layout.
#include "circle.hpp"
#include "common.cuh"

namespace cm31 {

// --------------------------------------------------------------------------- twiddles
// tree[level j][i] = x( C_j.at(bit_reverse(i, k-1-j)) ),  C_j = root.repeated_double(j),
// root = half_odds(k), level j at offset 2^k - 2^(k-j)   (cpu/circle.rs:171-188)
__device__ __forceinline__ CirclePointM31 dev_point_from_index(u32 index) {
    CirclePointM31 res = {1, 0};
    CirclePointM31 cur = {M31_CIRCLE_GEN_X, M31_CIRCLE_GEN_Y};
    index &= 0x7fffffffu;
    while (index) {
        if (index & 1) res = cp_add(res, cur);
        cur = cp_double(cur);
        index >>= 1;
    }
    return res;
}

// Point of index initial_j + step_j * nat as I_j + sum over set bits b of nat of T[j + b]
// (T[b] = point(step * 2^b), I_j = point(initial * 2^j)): <= k-1 group additions instead of a
// 31-step double-and-add; the 1/x tree uses one Montgomery batch inversion per 4 entries.
struct TwiddleTables {
    CirclePointM31 t[32];  // T[b]
    CirclePointM31 i[32];  // I[j]
};
__device__ __forceinline__ u32 twiddle_x(const TwiddleTables& tab, u32 k, u32 t, u32 total) {
    u32 rem = total - t;                  // in (1, 2^k]
    u32 j = k - (32 - __clz(rem - 1));    // level: 2^(k-j-1) < rem <= 2^(k-j)
    u32 level_off = total - (1u << (k - j));
    u32 nat = bit_reverse(t - level_off, k - 1 - j);
    CirclePointM31 p = tab.i[j];
    for (u32 b = 0; nat != 0; b++, nat >>= 1)
        if (nat & 1) p = cp_add(p, tab.t[j + b]);
    return p.x;
}
__global__ void __launch_bounds__(256) twiddle_kernel(u32* tw, u32* itw, u32 k, const __grid_constant__ TwiddleTables tab) {
    const u32 total = 1u << k;
    const u32 t0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (t0 >= total) return;
    u32 x[4];
    const u32 rem = total - t0;
    const u32 j = k - (32 - __clz(rem - 1));
    const u32 bits = k - 1 - j;  // level j holds 2^bits entries
    if (bits >= 2) {
        // The thread's 4 entries sit in one level and their indices differ only in the low 2 bits, i.e. (bit-reversed) in the
        // TOP 2 bits of `nat`: the group additions for the shared low bits are done once, then x(P + C_q) for
        // C_q in {0, T[top], T[top-1], T[top] + T[top-1]} (2 products each) — ~2.4x fewer products than 4 independent sums.
        const u32 level_off = total - (1u << (k - j));
        u32 nat = bit_reverse(t0 - level_off, bits);  // low 2 bits of the index are 0 -> top 2 bits of nat are 0
        CirclePointM31 p = tab.i[j];
        for (u32 b = 0; nat != 0; b++, nat >>= 1)
            if (nat & 1) p = cp_add(p, tab.t[j + b]);
        const CirclePointM31 c1 = tab.t[j + bits - 1], c2 = tab.t[j + bits - 2], c3 = cp_add(c1, c2);
        x[0] = p.x;
        x[1] = m31_sub(m31_mul(p.x, c1.x), m31_mul(p.y, c1.y));  // index +1 -> top bit of nat
        x[2] = m31_sub(m31_mul(p.x, c2.x), m31_mul(p.y, c2.y));  // index +2 -> second bit from the top
        x[3] = m31_sub(m31_mul(p.x, c3.x), m31_mul(p.y, c3.y));
    } else {
#pragma unroll
        for (u32 q = 0; q < 4; q++) x[q] = (t0 + q == total - 1) ? 1u : twiddle_x(tab, k, t0 + q, total);  // pad element (cpu/circle.rs:185-187)
    }
    // batch inverse of 4 (fields/mod.rs:69-99)
    u32 p01 = m31_mul(x[0], x[1]), p012 = m31_mul(p01, x[2]), p0123 = m31_mul(p012, x[3]);
    u32 inv = m31_inv(p0123);
    u32 i3 = m31_mul(inv, p012);
    inv = m31_mul(inv, x[3]);
    u32 i2 = m31_mul(inv, p01);
    inv = m31_mul(inv, x[2]);
    u32 i1 = m31_mul(inv, x[0]);
    u32 i0 = m31_mul(inv, x[1]);
    *reinterpret_cast<uint4*>(tw + t0) = make_uint4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<uint4*>(itw + t0) = make_uint4(i0, i1, i2, i3);
}

// --------------------------------------------------------------------------- FFT passes
__device__ __forceinline__ u32 pad_idx(u32 s) { return s + ((s >> 6) << 3); }

// Twiddle for FFT layer `layer` (0 = circle layer) and butterfly group h, domain log size L,
// tree log size M (tree holds 2^(M-1) words).
__device__ __forceinline__ u32 load_twiddle(const u32* __restrict__ tree, u32 M, u32 L, u32 layer, u32 h) {
    u32 end = 1u << (M - 1);
    if (layer == 0) {
        const u32* line0 = tree + (end - (1u << (L - 1)));
        u32 x = __ldg(line0 + 2 * (h >> 2));
        u32 y = __ldg(line0 + 2 * (h >> 2) + 1);
        u32 sel = h & 3;  // [y, -y, -x, x]   (cpu/circle.rs:209-229)
        u32 v = (sel < 2) ? y : x;
        return (sel == 1 || sel == 2) ? m31_neg(v) : v;
    }
    return __ldg(tree + (end - (1u << (L - layer))) + h);
}

template <bool INV>
__device__ __forceinline__ void bfly(u32& v0, u32& v1, u32 t) {
    if (INV) {  // ibutterfly (core/fft.rs:14-21)
        u32 tmp = v0;
        v0 = m31_add(tmp, v1);
        v1 = m31_mul(m31_sub(tmp, v1), t);
    } else {  // butterfly (core/fft.rs:5-12)
        u32 tmp = m31_mul(v1, t);
        v1 = m31_sub(v0, tmp);
        v0 = m31_add(v0, tmp);
    }
}

// One pass over layers [lo, lo+nl).  Tile = 2^nl strided rows x 2^b contiguous words.
// grid.x = tile * n_cols + column.
template <bool INV>
__global__ void __launch_bounds__(1024) fft_pass_kernel(const u32* const* __restrict__ src_cols,
                                                        u32* const* __restrict__ dst_cols, u32 L, u32 log_in,
                                                        u32 lo, u32 nl, u32 b, const u32* __restrict__ tree,
                                                        u32 M, u32 scale, u32 n_cols) {
    extern __shared__ u32 smem[];
    const u32 tile_log = nl + b;
    const u32 tile = 1u << tile_log;
    // 1-D grid, column fastest: CTAs that run together work on the same tile of different
    // columns and share its twiddles through L2.
    const u32 col = blockIdx.x % n_cols;
    const u32 tile_id = blockIdx.x / n_cols;
    const u32* __restrict__ src = src_cols[col];
    u32* __restrict__ dst = dst_cols[col];
    const u32 lo_hi = tile_id & ((1u << (lo - b)) - 1);
    const u32 hi = tile_id >> (lo - b);
    const u32 n_in = 1u << log_in;
    const u32 bmask = (1u << b) - 1;
    const size_t gbase = ((size_t)hi << (lo + nl)) | ((size_t)lo_hi << b);

    // ---- load
    if (b == 0 && log_in >= 2) {
        // contiguous tile: 128-bit loads
        const uint4* s4 = reinterpret_cast<const uint4*>(src + gbase);
        for (u32 s = threadIdx.x * 4; s < tile; s += blockDim.x * 4) {
            uint4 v;
            if (gbase + s < n_in) v = __ldg(s4 + (s >> 2));
            else v = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(&smem[pad_idx(s)]) = v;
        }
    } else {
        for (u32 s = threadIdx.x; s < tile; s += blockDim.x) {
            size_t g = gbase | ((size_t)(s >> b) << lo) | (s & bmask);
            smem[pad_idx(s)] = g < n_in ? __ldg(src + g) : 0u;
        }
    }
    __syncthreads();

    // ---- radix-8 rounds over groups of <=3 layers
    const u32 n_rounds = (nl + 2) / 3;
    for (u32 rr = 0; rr < n_rounds; rr++) {
        const u32 r = INV ? rr : (n_rounds - 1 - rr);
        const u32 p = b + 3 * r;
        const u32 k = min(3u, nl - 3 * r);
        const u32 groups = tile >> k;
        const u32 hshift = nl - 3 * r - k;
        for (u32 q = threadIdx.x; q < groups; q += blockDim.x) {
            const u32 low = q & ((1u << p) - 1);
            const u32 high = q >> p;
            const u32 base = (high << (p + k)) | low;
            const u32 H = (hi << hshift) | high;
            const u32 layer0 = lo + 3 * r;
            if (k == 3) {
                u32 v[8];
                if (p == 0) {
                    uint4 a = *reinterpret_cast<uint4*>(&smem[pad_idx(base)]);
                    uint4 c = *reinterpret_cast<uint4*>(&smem[pad_idx(base) + 4]);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                    v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
                } else {
#pragma unroll
                    for (u32 j = 0; j < 8; j++) v[j] = smem[pad_idx(base + (j << p))];
                }
                if (INV) {
#pragma unroll
                    for (u32 jj = 0; jj < 4; jj++) {
                        u32 t = load_twiddle(tree, M, L, layer0, 4 * H + jj);
                        bfly<true>(v[2 * jj], v[2 * jj + 1], t);
                    }
#pragma unroll
                    for (u32 jj = 0; jj < 2; jj++) {
                        u32 t = load_twiddle(tree, M, L, layer0 + 1, 2 * H + jj);
                        bfly<true>(v[4 * jj], v[4 * jj + 2], t);
                        bfly<true>(v[4 * jj + 1], v[4 * jj + 3], t);
                    }
                    {
                        u32 t = load_twiddle(tree, M, L, layer0 + 2, H);
#pragma unroll
                        for (u32 jj = 0; jj < 4; jj++) bfly<true>(v[jj], v[jj + 4], t);
                    }
                } else {
                    {
                        u32 t = load_twiddle(tree, M, L, layer0 + 2, H);
#pragma unroll
                        for (u32 jj = 0; jj < 4; jj++) bfly<false>(v[jj], v[jj + 4], t);
                    }
#pragma unroll
                    for (u32 jj = 0; jj < 2; jj++) {
                        u32 t = load_twiddle(tree, M, L, layer0 + 1, 2 * H + jj);
                        bfly<false>(v[4 * jj], v[4 * jj + 2], t);
                        bfly<false>(v[4 * jj + 1], v[4 * jj + 3], t);
                    }
#pragma unroll
                    for (u32 jj = 0; jj < 4; jj++) {
                        u32 t = load_twiddle(tree, M, L, layer0, 4 * H + jj);
                        bfly<false>(v[2 * jj], v[2 * jj + 1], t);
                    }
                }
                if (p == 0) {
                    *reinterpret_cast<uint4*>(&smem[pad_idx(base)]) = make_uint4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<uint4*>(&smem[pad_idx(base) + 4]) = make_uint4(v[4], v[5], v[6], v[7]);
                } else {
#pragma unroll
                    for (u32 j = 0; j < 8; j++) smem[pad_idx(base + (j << p))] = v[j];
                }
            } else if (k == 2) {
                u32 v[4];
#pragma unroll
                for (u32 j = 0; j < 4; j++) v[j] = smem[pad_idx(base + (j << p))];
                if (INV) {
#pragma unroll
                    for (u32 jj = 0; jj < 2; jj++) {
                        u32 t = load_twiddle(tree, M, L, layer0, 2 * H + jj);
                        bfly<true>(v[2 * jj], v[2 * jj + 1], t);
                    }
                    u32 t = load_twiddle(tree, M, L, layer0 + 1, H);
                    bfly<true>(v[0], v[2], t);
                    bfly<true>(v[1], v[3], t);
                } else {
                    u32 t = load_twiddle(tree, M, L, layer0 + 1, H);
                    bfly<false>(v[0], v[2], t);
                    bfly<false>(v[1], v[3], t);
#pragma unroll
                    for (u32 jj = 0; jj < 2; jj++) {
                        u32 t2 = load_twiddle(tree, M, L, layer0, 2 * H + jj);
                        bfly<false>(v[2 * jj], v[2 * jj + 1], t2);
                    }
                }
#pragma unroll
                for (u32 j = 0; j < 4; j++) smem[pad_idx(base + (j << p))] = v[j];
            } else {
                u32 v0 = smem[pad_idx(base)], v1 = smem[pad_idx(base + (1u << p))];
                u32 t = load_twiddle(tree, M, L, layer0, H);
                bfly<INV>(v0, v1, t);
                smem[pad_idx(base)] = v0;
                smem[pad_idx(base + (1u << p))] = v1;
            }
        }
        __syncthreads();
    }

    // ---- store
    if (b == 0) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + gbase);
        for (u32 s = threadIdx.x * 4; s < tile; s += blockDim.x * 4) {
            uint4 v = *reinterpret_cast<uint4*>(&smem[pad_idx(s)]);
            if (scale != 1) {
                v.x = m31_mul(v.x, scale); v.y = m31_mul(v.y, scale);
                v.z = m31_mul(v.z, scale); v.w = m31_mul(v.w, scale);
            }
            d4[s >> 2] = v;
        }
    } else {
        for (u32 s = threadIdx.x; s < tile; s += blockDim.x) {
            size_t g = gbase | ((size_t)(s >> b) << lo) | (s & bmask);
            u32 v = smem[pad_idx(s)];
            if (scale != 1) v = m31_mul(v, scale);
            dst[g] = v;
        }
    }
}

// Domains of log size 1 and 2 (cpu/circle.rs:26-50, 107-124): one thread per column.
template <bool INV>
__global__ void fft_small_kernel(const u32* const* src_cols, u32* const* dst_cols, u32 n_cols, u32 L, u32 log_in,
                                 u32 px, u32 py) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const u32* src = src_cols[c];
    u32* dst = dst_cols[c];
    u32 n = 1u << L, n_in = 1u << log_in;
    u32 v[4];
    for (u32 i = 0; i < n; i++) v[i] = i < n_in ? src[i] : 0u;
    if (L == 1) {
        if (INV) {
            u32 y_inv = m31_inv(py);
            bfly<true>(v[0], v[1], y_inv);
            u32 n_inv = m31_inv(2);
            v[0] = m31_mul(v[0], n_inv);
            v[1] = m31_mul(v[1], n_inv);
        } else {
            bfly<false>(v[0], v[1], py);
        }
    } else {
        if (INV) {
            u32 x_inv = m31_inv(px), y_inv = m31_inv(py);
            bfly<true>(v[0], v[1], y_inv);
            bfly<true>(v[2], v[3], m31_neg(y_inv));
            bfly<true>(v[0], v[2], x_inv);
            bfly<true>(v[1], v[3], x_inv);
            u32 n_inv = m31_inv(4);
            for (u32 i = 0; i < 4; i++) v[i] = m31_mul(v[i], n_inv);
        } else {
            bfly<false>(v[0], v[2], px);
            bfly<false>(v[1], v[3], px);
            bfly<false>(v[0], v[1], py);
            bfly<false>(v[2], v[3], m31_neg(py));
        }
    }
    for (u32 i = 0; i < n; i++) dst[i] = v[i];
}

__global__ void bit_reverse_kernel(u32* col, u32 log_size) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << log_size)) return;
    u32 j = bit_reverse((u32)i, log_size);
    if (i < j) {
        u32 a = col[i], c = col[j];
        col[i] = c;
        col[j] = a;
    }
}

// --------------------------------------------------------------------------- eval_at_point
// f(P) = fold(coeffs, [pi^{L-2}(x), ..., pi(x), x, y])  (cpu/circle.rs:73-87, poly/utils.rs:44-55)
//      = sum_j c_j * prod_{bits i of j} factor_i,   factor_0 = y, factor_1 = x, factor_i = pi^{i-1}(x).
// Split j = (chunk | r | t): t = bits 0-7 (thread), r = bits 8-14 (per-thread loop), chunk = bits 15+.
//   stage 1 (one CTA per 2^15-coefficient chunk):  S_t = sum_r c[chunk, r, t] * mid[r]  is a dot product
//     of BASE-field coefficients with QM31 constants: 4 u64 multiply-accumulates per coefficient
//     (folded every 4 terms), coalesced 1 KB row reads;  partial = sum_t lo[t] * S_t  (one QM31
//     product per thread + a block reduction).
//   stage 2 (one CTA per polynomial): f(P) = sum_chunk hi[chunk] * partial[chunk].
// The monomial tables lo[256], mid[128], hi[<=512] per sample point are built on the host.
constexpr u32 EAP_LO_BITS = 8, EAP_MID_BITS = 7, EAP_CHUNK_LOG = EAP_LO_BITS + EAP_MID_BITS;
constexpr u32 EAP_MAX_LEVELS = 32;
constexpr u32 EAP_HI_MAX = 1024;                                            // polynomials up to 2^25 coefficients
constexpr u32 EAP_TABLE_STRIDE = (256 + 128 + EAP_HI_MAX) * 4;              // words per point

struct EapJob {
    const u32* coeffs;
    u32 log_size;
    u32 point;        // index into the monomial tables
    u32 partial_off;  // offset (in QM31s) of this poly's partials
    u32 first_chunk;  // index of this poly's first chunk in the flattened chunk list
};

__device__ __forceinline__ QM31 ld_qm(const u32* p) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    return qm_make(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ u64 eap_fold64(u64 x) { return (x & P) + (x >> 31); }

__device__ __forceinline__ QM31 block_sum_qm(QM31 v, QM31* sh) {
    const u32 t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (u32 stride = 128; stride > 0; stride >>= 1) {
        if (t < stride) sh[t] = qm_add(sh[t], sh[t + stride]);
        __syncthreads();
    }
    return sh[0];
}

__global__ void __launch_bounds__(256) eap_stage1_kernel(const EapJob* __restrict__ jobs, const u32* __restrict__ chunk_job,
                                                          const u32* __restrict__ tables, u32* __restrict__ partials) {
    __shared__ QM31 sh[256];
    const u32 chunk = blockIdx.x;
    const EapJob job = jobs[chunk_job[chunk]];
    const u32 local_chunk = chunk - job.first_chunk;
    const u32 L = job.log_size;
    const u32 lo_bits = min(L, EAP_LO_BITS);
    const u32 mid_bits = L > EAP_LO_BITS ? min(L - EAP_LO_BITS, EAP_MID_BITS) : 0;
    const u32* c = job.coeffs + ((size_t)local_chunk << EAP_CHUNK_LOG);
    const u32* tab = tables + (size_t)job.point * EAP_TABLE_STRIDE;
    const u32* mid = tab + 256 * 4;
    const u32 t = threadIdx.x;
    QM31 v = qm_zero();
    if (t < (1u << lo_bits)) {
        u64 a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        const u32 n_r = 1u << mid_bits;
        for (u32 r0 = 0; r0 < n_r; r0 += 4) {
#pragma unroll
            for (u32 j = 0; j < 4; j++) {
                const u32 r = r0 + j;
                if (r < n_r) {
                    const u64 f = __ldg(c + ((size_t)r << EAP_LO_BITS) + t);
                    const uint4 m = __ldg(reinterpret_cast<const uint4*>(mid) + r);
                    a0 += f * m.x;
                    a1 += f * m.y;
                    a2 += f * m.z;
                    a3 += f * m.w;
                }
            }
            a0 = eap_fold64(a0);
            a1 = eap_fold64(a1);
            a2 = eap_fold64(a2);
            a3 = eap_fold64(a3);
        }
        QM31 s = qm_make(m31_reduce64(a0), m31_reduce64(a1), m31_reduce64(a2), m31_reduce64(a3));
        v = qm_mul(ld_qm(tab + t * 4), s);
    }
    QM31 r = block_sum_qm(v, sh);
    if (t == 0) {
        u32* o = partials + ((size_t)job.partial_off + local_chunk) * 4;
        o[0] = r.a; o[1] = r.b; o[2] = r.c; o[3] = r.d;
    }
}

__global__ void __launch_bounds__(256) eap_stage2_kernel(const EapJob* __restrict__ jobs, const u32* __restrict__ tables,
                                                          const u32* __restrict__ partials, u32* __restrict__ out) {
    __shared__ QM31 sh[256];
    const EapJob job = jobs[blockIdx.x];
    const u32 n_part = job.log_size > EAP_CHUNK_LOG ? 1u << (job.log_size - EAP_CHUNK_LOG) : 1u;
    const u32* hi = tables + (size_t)job.point * EAP_TABLE_STRIDE + (256 + 128) * 4;
    const u32* p = partials + (size_t)job.partial_off * 4;
    QM31 v = qm_zero();
    for (u32 i = threadIdx.x; i < n_part; i += 256) v = qm_add(v, qm_mul(ld_qm(hi + i * 4), ld_qm(p + (size_t)i * 4)));
    QM31 r = block_sum_qm(v, sh);
    if (threadIdx.x == 0) {
        u32* o = out + (size_t)blockIdx.x * 4;
        o[0] = r.a; o[1] = r.b; o[2] = r.c; o[3] = r.d;
    }
}

// --------------------------------------------------------------------------- host planning
struct Pass {
    u32 lo, nl, b;
};
static void plan_passes(u32 L, std::vector<Pass>& out) {
    // first pass: contiguous tile of up to 2^12 words; the rest: strided passes of <= 12 layers
    // with 64-byte (b=4) or 32-byte (b=3, 12 layers) contiguous runs.
    const u32 FIRST_MAX = 12, REST_MAX = 12;
    out.clear();
    if (L <= FIRST_MAX) {
        out.push_back({0, L, 0});
        return;
    }
    u32 n_rest = (L - FIRST_MAX + REST_MAX - 1) / REST_MAX;
    u32 rest_total = L - FIRST_MAX;
    if (n_rest == 1 && rest_total < 6) rest_total = 6 < L / 2 ? 6 : L / 2;
    u32 first = L - rest_total;
    out.push_back({0, first, 0});
    u32 lo = first;
    u32 base = rest_total / n_rest, extra = rest_total % n_rest;
    for (u32 i = 0; i < n_rest; i++) {
        u32 nl = base + (i < extra ? 1 : 0);
        out.push_back({lo, nl, nl >= 12 ? 3u : 4u});
        lo += nl;
    }
}

// fft4.cu: four-columns-per-CTA passes for 2^12..2^24 points (returns 1 when it does not apply)
template <bool INV>
int run_fft4(const u32* const* src, u32* const* dst, size_t n_cols, u32 L, u32 log_in, const u32* tree, u32 M, u32 scale_last);

template <bool INV>
static int run_fft(const u32* const* src_host, u32* const* dst_host, size_t n_cols, u32 L, u32 log_in,
                   const cm31_twiddles* tw) {
    if (n_cols == 0) return 0;
    CM_REQUIRE(tw != nullptr, "fft: null twiddles");
    CM_REQUIRE(L >= 1 && L <= tw->log_size, "fft: domain larger than the twiddle tree");
    CM_REQUIRE(log_in <= L, "fft: more coefficients than domain points");
    DeviceTable dsrc, ddst;
    if (int e = dsrc.upload(src_host, n_cols * sizeof(void*))) return e;
    if (int e = ddst.upload(dst_host, n_cols * sizeof(void*))) return e;
    const u32* const* src = (const u32* const*)dsrc.d;
    u32* const* dst = (u32* const*)ddst.d;
    if (L <= 2) {
        CirclePointM31 p = CanonicCoset(L).half_coset().initial();
        ProfScope prof(INV ? "ifft_small" : "rfft_small", 4ull * n_cols * ((1ull << L) + (1ull << log_in)));
        fft_small_kernel<INV><<<(unsigned)((n_cols + 127) / 128), 128, 0, stream()>>>(src, dst, (u32)n_cols, L, log_in,
                                                                                     p.x, p.y);
        CM_LAUNCH_CHECK();
        return 0;
    }
    const u32* tree = INV ? tw->itw : tw->tw;
    u32 scale_last = 1;
    if (INV) scale_last = m31_inv((u32)(((u64)1 << L) % P));
    {
        int e = run_fft4<INV>(src, dst, n_cols, L, log_in, tree, tw->log_size, scale_last);
        if (e != 1) return e;
    }
    std::vector<Pass> passes;
    plan_passes(L, passes);
    size_t np = passes.size();
    for (size_t i = 0; i < np; i++) {
        const Pass& ps = INV ? passes[i] : passes[np - 1 - i];
        bool first = (i == 0), last = (i == np - 1);
        u32 tile_log = ps.nl + ps.b;
        u32 tile = 1u << tile_log;
        u32 threads = tile / 8;
        if (threads < 32) threads = 32;
        if (threads > 1024) threads = 1024;
        size_t smem = (size_t)(tile + ((tile >> 6) << 3) + 8) * 4;
        auto kern = fft_pass_kernel<INV>;
        if (smem > 48 * 1024) CM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        size_t n_blocks = n_cols * (((size_t)1 << L) >> tile_log);
        CM_REQUIRE(n_blocks < (1ull << 31), "fft: batch too large for one launch");
        // algorithmic bytes of the whole transform (read 2^log_in, write 2^L words per column) split evenly over its passes
        ProfScope prof(INV ? "ifft_pass" : "rfft_pass", 4ull * n_cols * ((1ull << L) + (1ull << log_in)) / np);
        prof_ops(n_cols * (3ull * ps.nl * (1ull << (L - 1)) + ((INV && last) ? (1ull << L) : 0)));
        kern<<<(unsigned)n_blocks, threads, smem, stream()>>>(first ? src : (const u32* const*)dst, dst, L,
                                                              first ? log_in : L, ps.lo, ps.nl, ps.b, tree,
                                                              tw->log_size, (INV && last) ? scale_last : 1u,
                                                              (u32)n_cols);
        CM_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace cm31

using namespace cm31;

extern "C" {

int cm31_twiddles_create(uint32_t log_size, cm31_twiddles** out) {
    CM_REQUIRE(out != nullptr, "twiddles_create: null out");
    CM_REQUIRE(log_size >= 3 && log_size <= 30, "twiddles_create: log_size must be in [3,30]");
    cm31_twiddles* tw = new cm31_twiddles();
    tw->log_size = log_size;
    u32 k = log_size - 1;
    size_t n = (size_t)1 << k;
    if (int e = cm31_malloc((void**)&tw->tw, n * 4)) return e;
    if (int e = cm31_malloc((void**)&tw->itw, n * 4)) return e;
    Coset root = CanonicCoset(log_size).half_coset();
    TwiddleTables tab;
    for (u32 b = 0; b < 32; b++) {
        tab.t[b] = cp_from_index((u32)(((u64)root.step_size << b) & 0x7fffffffu));
        tab.i[b] = cp_from_index((u32)(((u64)root.initial_index << b) & 0x7fffffffu));
    }
    ProfScope prof("twiddles", 8ull * n);
    twiddle_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, stream()>>>(tw->tw, tw->itw, k, tab);
    CM_LAUNCH_CHECK();
    *out = tw;
    return 0;
}

int cm31_twiddles_destroy(cm31_twiddles* tw) {
    if (!tw) return 0;
    cm31_free(tw->tw);
    cm31_free(tw->itw);
    delete tw;
    return 0;
}

int cm31_twiddles_buffers(const cm31_twiddles* tw, const uint32_t** twiddles, const uint32_t** itwiddles,
                          uint32_t* log_size) {
    CM_REQUIRE(tw != nullptr, "twiddles_buffers: null handle");
    if (twiddles) *twiddles = tw->tw;
    if (itwiddles) *itwiddles = tw->itw;
    if (log_size) *log_size = tw->log_size;
    return 0;
}

int cm31_interpolate_batch(uint32_t* const* cols, size_t n_cols, uint32_t log_size, const cm31_twiddles* tw) {
    return run_fft<true>((const u32* const*)cols, cols, n_cols, log_size, log_size, tw);
}

int cm31_interpolate_batch_to(const uint32_t* const* evals, uint32_t* const* coeffs_out, size_t n_cols, uint32_t log_size,
                              const cm31_twiddles* tw) {
    return run_fft<true>(evals, coeffs_out, n_cols, log_size, log_size, tw);
}

int cm31_evaluate_batch(const uint32_t* const* coeffs, uint32_t* const* out, size_t n_cols, uint32_t log_size,
                        uint32_t log_eval_size, const cm31_twiddles* tw) {
    CM_REQUIRE(log_eval_size >= log_size, "evaluate: domain smaller than the polynomial");
    return run_fft<false>(coeffs, out, n_cols, log_eval_size, log_size, tw);
}

int cm31_bit_reverse(uint32_t* col, uint32_t log_size) {
    size_t n = (size_t)1 << log_size;
    ProfScope prof("bit_reverse", 8ull * n);
    bit_reverse_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>(col, log_size);
    CM_LAUNCH_CHECK();
    return 0;
}

int cm31_eval_at_point_batch(const uint32_t* const* coeffs, const uint32_t* log_sizes_host, size_t n_polys,
                             const uint32_t* points_host, size_t n_points, const uint32_t* point_idx_host,
                             uint32_t* out_host) {
    if (n_polys == 0) return 0;
    // monomial tables per point: lo[t] over factors 0..7, mid[r] over 8..14, hi[k] over 15..
    u32 max_log = 0;
    for (size_t i = 0; i < n_polys; i++) max_log = std::max(max_log, log_sizes_host[i]);
    CM_REQUIRE(max_log <= EAP_CHUNK_LOG + 10, "eval_at_point: polynomial too large");
    const u32 n_hi = max_log > EAP_CHUNK_LOG ? 1u << (max_log - EAP_CHUNK_LOG) : 1u;
    std::vector<u32> factors(n_points * EAP_TABLE_STRIDE, 0);
    for (size_t k = 0; k < n_points; k++) {
        const u32* pt = points_host + k * 8;
        QM31 x = qm_make(pt[0], pt[1], pt[2], pt[3]);
        QM31 y = qm_make(pt[4], pt[5], pt[6], pt[7]);
        QM31 fac[EAP_MAX_LEVELS];
        fac[0] = y;
        for (u32 i = 1; i < EAP_MAX_LEVELS; i++) {
            fac[i] = x;
            x = qm_double_x(x);
        }
        u32* tab = &factors[k * EAP_TABLE_STRIDE];
        auto fill = [&](u32* dst, u32 count, u32 first_factor) {
            // dst[j] = prod over set bits b of j of fac[first_factor + b]; dst[j] from dst[j without top bit]
            std::vector<QM31> m(count);
            m[0] = qm_one();
            for (u32 j = 1; j < count; j++) {
                u32 top = 31 - __builtin_clz(j);
                m[j] = qm_mul(m[j & ~(1u << top)], fac[first_factor + top]);
            }
            for (u32 j = 0; j < count; j++) {
                dst[4 * j] = m[j].a; dst[4 * j + 1] = m[j].b; dst[4 * j + 2] = m[j].c; dst[4 * j + 3] = m[j].d;
            }
        };
        fill(tab, 256, 0);
        fill(tab + 256 * 4, 128, EAP_LO_BITS);
        fill(tab + (256 + 128) * 4, n_hi, EAP_CHUNK_LOG);
    }
    std::vector<EapJob> jobs(n_polys);
    std::vector<u32> chunk_job;
    u32 part_off = 0;
    for (size_t i = 0; i < n_polys; i++) {
        u32 L = log_sizes_host[i];
        CM_REQUIRE(point_idx_host[i] < n_points, "eval_at_point: point index out of range");
        u32 ch_log = L < EAP_CHUNK_LOG ? L : EAP_CHUNK_LOG;
        u32 n_chunks = 1u << (L - ch_log);
        jobs[i].coeffs = coeffs[i];
        jobs[i].log_size = L;
        jobs[i].point = point_idx_host[i];
        jobs[i].partial_off = part_off;
        jobs[i].first_chunk = (u32)chunk_job.size();
        for (u32 c = 0; c < n_chunks; c++) chunk_job.push_back((u32)i);
        part_off += n_chunks;
    }
    DeviceTable djobs, dchunk, dfac;
    if (int e = djobs.upload(jobs.data(), jobs.size() * sizeof(EapJob))) return e;
    if (int e = dchunk.upload(chunk_job.data(), chunk_job.size() * 4)) return e;
    if (int e = dfac.upload(factors.data(), factors.size() * 4)) return e;
    u32 *dpart = nullptr, *dout = nullptr;
    CM_CUDA(cudaMallocAsync(&dpart, (size_t)part_off * 16, stream()));
    CM_CUDA(cudaMallocAsync(&dout, n_polys * 16, stream()));
    {
        uint64_t coeff_bytes = 0;
        for (size_t i = 0; i < n_polys; i++) coeff_bytes += 4ull << log_sizes_host[i];
        ProfScope prof("eval_at_point", coeff_bytes, 2);
        prof_ops(coeff_bytes / 4 * 8);  // per coefficient: QM31 x M31 (4 mul) + QM31 add (4)
        eap_stage1_kernel<<<(unsigned)chunk_job.size(), 256, 0, stream()>>>((const EapJob*)djobs.d, (const u32*)dchunk.d,
                                                                           (const u32*)dfac.d, dpart);
        eap_stage2_kernel<<<(unsigned)n_polys, 256, 0, stream()>>>((const EapJob*)djobs.d, (const u32*)dfac.d, dpart, dout);
    }
    CM_LAUNCH_CHECK();
    CM_CUDA(cudaMemcpyAsync(out_host, dout, n_polys * 16, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaFreeAsync(dpart, stream()));
    CM_CUDA(cudaFreeAsync(dout, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

}  // extern "C"

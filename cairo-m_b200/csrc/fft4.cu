// Circle (I)FFT passes for large columns (2^12 .. 2^24 points): FOUR COLUMNS PER CTA.
//
// Replaces external/stwo/crates/prover/src/core/backend/simd/fft/{ifft,rfft}.rs (3 layers per
// pass over one column, twiddles re-read per column); bit-exact against the CpuBackend definition
// external/stwo/crates/prover/src/core/backend/cpu/circle.rs:18-135,190-229.
//
// Why four columns: the butterfly network (shared-memory indices, twiddle values, bounds) is the
// same for every column of a batch, and trace columns always come in batches.  A
// tile element is a uint4 = the same point of 4 columns, so every LDS/STS moves 4 columns at once
// (128-bit, conflict-free with one uint4 of padding per 8), twiddles are fetched once per radix-8
// group and the integer index math is amortised 4x.  The kernel is integer-issue bound
// (DESIGN.md §4): what matters is instructions per butterfly, not bytes.
//
// A pass covers layers [lo, lo+NL) on a tile of 2^NL strided rows x 2^B contiguous words
// (B = 0: the first, contiguous pass with lo = 0).  Tile = 2^TL points x 4 columns, TL = NL + B.
#include "circle.hpp"
#include "common.cuh"

namespace cm31 {

__device__ __forceinline__ u32 fft4_pad(u32 e) { return e + (e >> 3); }

__device__ __forceinline__ u32 fft4_twiddle(const u32* __restrict__ tree, u32 M, u32 L, u32 layer, u32 h) {
    u32 end = 1u << (M - 1);
    if (layer == 0) {  // circle layer: [y, -y, -x, x] from the first line layer (cpu/circle.rs:209-229)
        const u32* line0 = tree + (end - (1u << (L - 1)));
        uint2 xy = __ldg(reinterpret_cast<const uint2*>(line0) + (h >> 2));
        u32 sel = h & 3;
        u32 v = (sel < 2) ? xy.y : xy.x;
        return (sel == 1 || sel == 2) ? m31_neg(v) : v;
    }
    return __ldg(tree + (end - (1u << (L - layer))) + h);
}

// a * t mod P with the twiddle passed DOUBLED (t2 = 2t < 2^32), as stwo's SIMD backend does (simd/m31.rs:110-226
// `mul_doubled`): the 64-bit product 2*a*t has floor(a*t / 2^31) in its upper word and 2 * (a*t mod 2^31) in its lower
// word, so the split costs one shift instead of a mask and a 64-bit funnel shift -- one ALU-pipe instruction less per
// butterfly in a kernel that is bound by exactly that pipe (4 FMA-pipe + 5 ALU-pipe instructions per butterfly before).
__device__ __forceinline__ u32 m31_mul_dbl(u32 a, u32 t2) {
    const u64 p = (u64)a * t2;
    const u32 s = ((u32)p >> 1) + (u32)(p >> 32);  // < 2^31 + 2^31
    const u32 r = s - P;
    return r < s ? r : s;
}
template <bool INV>
__device__ __forceinline__ void bfly1(u32& v0, u32& v1, u32 t) {  // t: the DOUBLED twiddle
    if (INV) {  // ibutterfly (core/fft.rs:14-21)
        u32 tmp = v0;
        v0 = m31_add(tmp, v1);
        v1 = m31_mul_dbl(m31_sub(tmp, v1), t);
    } else {  // butterfly (core/fft.rs:5-12)
        u32 tmp = m31_mul_dbl(v1, t);
        v1 = m31_sub(v0, tmp);
        v0 = m31_add(v0, tmp);
    }
}
template <bool INV>
__device__ __forceinline__ void bfly4(uint4& a, uint4& b, u32 t) {
    bfly1<INV>(a.x, b.x, t);
    bfly1<INV>(a.y, b.y, t);
    bfly1<INV>(a.z, b.z, t);
    bfly1<INV>(a.w, b.w, t);
}

// K layers (layer0 .. layer0+K-1) on the 2^K uint4 values of one group.  Layer l pairs values
// (j, j + 2^l); its twiddle index is (H << (K-1-l)) | (j >> (l+1)), H = index of the group in the
// top layer of the round.
// The doubled twiddles of the whole tile are staged in shared memory once, next to the tile (tws): layer li of the pass
// (li = 0 .. NL-1) owns 2^(NL-1-li) entries at offset 2^NL - 2^(NL-li), indexed by the tile-local group number.  A radix-8
// round then reads its 4 + 2 + 1 twiddles as one LDS.128, one LDS.64 and one LDS.32 (shared-memory latency, broadcast inside a
// warp) instead of 7 global loads whose L2 latency every warp of the CTA paid right after each barrier.
// ZT (forward only): the round's top layer pairs a value with a coefficient that is zero by construction (the upper half of
// a zero-padded low-degree extension), so its butterfly v0 +- t*0 is a copy: no twiddle, no multiply.
template <bool INV, int K, bool ZT = false>
__device__ __forceinline__ void radix_round(uint4 (&v)[1 << K], const u32* __restrict__ tws, u32 NLr, u32 high) {
    // tws: first entry of the round's lowest layer; NLr = layers of the pass at and above it (NL - 3r): layer l of the round
    // starts 2^NLr - 2^(NLr-l) entries further
    u32 tw[(1 << K) - 1];
    if (K == 3) {
        const uint4 t0 = *reinterpret_cast<const uint4*>(tws + 4 * high);
        tw[0] = t0.x; tw[1] = t0.y; tw[2] = t0.z; tw[3] = t0.w;
        const uint2 t1 = *reinterpret_cast<const uint2*>(tws + ((1u << NLr) - (1u << (NLr - 1))) + 2 * high);
        tw[4] = t1.x; tw[5] = t1.y;
        if (!ZT) tw[6] = tws[((1u << NLr) - (1u << (NLr - 2))) + high];
    } else if (K == 2) {
        const uint2 t0 = *reinterpret_cast<const uint2*>(tws + 2 * high);
        tw[0] = t0.x; tw[1] = t0.y;
        if (!ZT) tw[2] = tws[((1u << NLr) - (1u << (NLr - 1))) + high];
    } else {
        if (!ZT) tw[0] = tws[high];
    }
#pragma unroll
    for (int ll = 0; ll < K; ll++) {
        const int l = INV ? ll : K - 1 - ll;
#pragma unroll
        for (int j = 0; j < (1 << K); j++)
            if (!(j & (1 << l))) {
                if (ZT && l == K - 1)
                    v[j | (1 << l)] = v[j];
                else
                    bfly4<INV>(v[j], v[j | (1 << l)], tw[((1 << K) - (1 << (K - l))) + (j >> (l + 1))]);
            }
    }
}

template <bool INV, int NL, int B, int THREADS, bool ZTOP = false>
__global__ void __launch_bounds__(THREADS, (THREADS <= 512 ? 2 : 1)) fft4_pass_kernel(const u32* const* __restrict__ src_cols, u32* const* __restrict__ dst_cols,
                                                            u32 L, u32 log_in, u32 lo, const u32* __restrict__ tree, u32 M, u32 scale,
                                                            u32 n_cols) {
    extern __shared__ uint4 sm4[];
    constexpr u32 TL = NL + B;
    constexpr u32 TILE = 1u << TL;
    constexpr u32 BMASK = (1u << B) - 1;
    const u32 tid = threadIdx.x;
    const u32 n_quads = (n_cols + 3) >> 2;
    const u32 quad = blockIdx.x % n_quads;  // column quad fastest: concurrent CTAs share twiddles through L2
    const u32 tile_id = blockIdx.x / n_quads;
    const u32* src[4];
    u32* dst[4];
    bool valid[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        u32 ci = quad * 4 + c;
        valid[c] = ci < n_cols;
        ci = valid[c] ? ci : n_cols - 1;
        src[c] = src_cols[ci];
        dst[c] = dst_cols[ci];
    }
    const u32 lo_hi = tile_id & ((1u << (lo - B)) - 1);
    const u32 hi = tile_id >> (lo - B);
    const u32 n_in = 1u << log_in;
    const size_t gbase = ((size_t)hi << (lo + NL)) | ((size_t)lo_hi << B);

    // ---- twiddles of the tile (all NL layers), doubled: entry e of layer li = tree twiddle (hi << (NL-1-li)) | e
    u32* tws = reinterpret_cast<u32*>(sm4 + (TILE + (TILE >> 3) + 1));
    for (u32 e = tid; e < (1u << NL) - 1; e += THREADS) {
        const u32 hb = 31 - __clz((1u << NL) - 1 - e);  // = NL - 1 - li
        const u32 li = NL - 1 - hb;
        const u32 idx = e - ((1u << NL) - (2u << hb));
        tws[e] = fft4_twiddle(tree, M, L, lo + li, (hi << hb) | idx) << 1;
    }

    // ---- load: tile point s of the 4 columns -> one uint4
    if (B == 0) {
        // contiguous tile: 128-bit loads along each column, 4x4 transpose in registers
        for (u32 s = tid * 4; s < TILE; s += THREADS * 4) {
            uint4 r[4];
#pragma unroll
            for (int c = 0; c < 4; c++) r[c] = (gbase + s < n_in) ? __ldg(reinterpret_cast<const uint4*>(src[c] + gbase + s)) : make_uint4(0, 0, 0, 0);
            sm4[fft4_pad(s)] = make_uint4(r[0].x, r[1].x, r[2].x, r[3].x);
            sm4[fft4_pad(s + 1)] = make_uint4(r[0].y, r[1].y, r[2].y, r[3].y);
            sm4[fft4_pad(s + 2)] = make_uint4(r[0].z, r[1].z, r[2].z, r[3].z);
            sm4[fft4_pad(s + 3)] = make_uint4(r[0].w, r[1].w, r[2].w, r[3].w);
        }
    } else {
        for (u32 s = tid; s < TILE; s += THREADS) {
            const size_t g = gbase | ((size_t)(s >> B) << lo) | (s & BMASK);
            uint4 v = make_uint4(0, 0, 0, 0);
            if (g < n_in) v = make_uint4(__ldg(src[0] + g), __ldg(src[1] + g), __ldg(src[2] + g), __ldg(src[3] + g));
            sm4[fft4_pad(s)] = v;
        }
    }
    __syncthreads();

    // ---- radix-8 register rounds (the last round takes the 1 or 2 leftover layers)
    constexpr int NR = (NL + 2) / 3;
#pragma unroll
    for (int rr = 0; rr < NR; rr++) {
        const int r = INV ? rr : NR - 1 - rr;
        const int p = B + 3 * r;
        const int k = (NL - 3 * r) < 3 ? (NL - 3 * r) : 3;
        const u32 groups = TILE >> k;
        const u32* tws_r = tws + ((1u << NL) - (1u << (NL - 3 * r)));  // first twiddle of layer 3r of the pass
        for (u32 q = tid; q < groups; q += THREADS) {
            const u32 low = q & ((1u << p) - 1);
            const u32 high = q >> p;
            const u32 base = (high << (p + k)) | low;
            if (k == 3) {
                uint4 v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = sm4[fft4_pad(base + (j << p))];
                if (ZTOP && rr == 0)
                    radix_round<INV, 3, true>(v, tws_r, NL - 3 * r, high);
                else
                    radix_round<INV, 3, false>(v, tws_r, NL - 3 * r, high);
#pragma unroll
                for (int j = 0; j < 8; j++) sm4[fft4_pad(base + (j << p))] = v[j];
            } else if (k == 2) {
                uint4 v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = sm4[fft4_pad(base + (j << p))];
                if (ZTOP && rr == 0)
                    radix_round<INV, 2, true>(v, tws_r, NL - 3 * r, high);
                else
                    radix_round<INV, 2, false>(v, tws_r, NL - 3 * r, high);
#pragma unroll
                for (int j = 0; j < 4; j++) sm4[fft4_pad(base + (j << p))] = v[j];
            } else {
                uint4 v[2];
#pragma unroll
                for (int j = 0; j < 2; j++) v[j] = sm4[fft4_pad(base + (j << p))];
                if (ZTOP && rr == 0)
                    radix_round<INV, 1, true>(v, tws_r, NL - 3 * r, high);
                else
                    radix_round<INV, 1, false>(v, tws_r, NL - 3 * r, high);
#pragma unroll
                for (int j = 0; j < 2; j++) sm4[fft4_pad(base + (j << p))] = v[j];
            }
        }
        __syncthreads();
    }

    // ---- store
    if (B == 0) {
        for (u32 s = tid * 4; s < TILE; s += THREADS * 4) {
            uint4 e[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                e[i] = sm4[fft4_pad(s + i)];
                if (scale != 1) {
                    e[i].x = m31_mul(e[i].x, scale);
                    e[i].y = m31_mul(e[i].y, scale);
                    e[i].z = m31_mul(e[i].z, scale);
                    e[i].w = m31_mul(e[i].w, scale);
                }
            }
            if (valid[0]) *reinterpret_cast<uint4*>(dst[0] + gbase + s) = make_uint4(e[0].x, e[1].x, e[2].x, e[3].x);
            if (valid[1]) *reinterpret_cast<uint4*>(dst[1] + gbase + s) = make_uint4(e[0].y, e[1].y, e[2].y, e[3].y);
            if (valid[2]) *reinterpret_cast<uint4*>(dst[2] + gbase + s) = make_uint4(e[0].z, e[1].z, e[2].z, e[3].z);
            if (valid[3]) *reinterpret_cast<uint4*>(dst[3] + gbase + s) = make_uint4(e[0].w, e[1].w, e[2].w, e[3].w);
        }
    } else {
        for (u32 s = tid; s < TILE; s += THREADS) {
            const size_t g = gbase | ((size_t)(s >> B) << lo) | (s & BMASK);
            uint4 v = sm4[fft4_pad(s)];
            if (scale != 1) {
                v.x = m31_mul(v.x, scale);
                v.y = m31_mul(v.y, scale);
                v.z = m31_mul(v.z, scale);
                v.w = m31_mul(v.w, scale);
            }
            if (valid[0]) dst[0][g] = v.x;
            if (valid[1]) dst[1][g] = v.y;
            if (valid[2]) dst[2][g] = v.z;
            if (valid[3]) dst[3][g] = v.w;
        }
    }
}

template <bool INV, int NL, int B, int THREADS, bool ZTOP = false>
static int launch_fft4(const u32* const* src, u32* const* dst, u32 L, u32 log_in, u32 lo, const u32* tree, u32 M, u32 scale,
                       size_t n_cols, size_t n_passes) {
    constexpr u32 TL = NL + B;
    constexpr size_t TILE = (size_t)1 << TL;
    const size_t smem = (TILE + (TILE >> 3) + 1) * 16 + ((size_t)4 << NL);  // tile (padded) + the tile's twiddles
    auto kern = fft4_pass_kernel<INV, NL, B, THREADS, ZTOP>;
    static bool attr_set = false;  // one per template instantiation
    if (!attr_set) {
        CM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const size_t n_quads = (n_cols + 3) / 4;
    const size_t n_blocks = n_quads * (((size_t)1 << L) >> TL);
    CM_REQUIRE(n_blocks < (1ull << 31), "fft: batch too large for one launch");
    ProfScope prof(INV ? "ifft_pass" : "rfft_pass", 4ull * n_cols * ((1ull << L) + (1ull << log_in)) / n_passes);
    prof_ops(n_cols * (3ull * NL * (1ull << (L - 1)) + ((INV && scale != 1) ? (1ull << L) : 0)));  // NL layers of 2^(L-1) butterflies (mul + add + sub)
    kern<<<(unsigned)n_blocks, THREADS, smem, stream()>>>(src, dst, L, log_in, lo, tree, M, scale, (u32)n_cols);
    CM_LAUNCH_CHECK();
    return 0;
}

// pass 2 of a two-pass transform: NL2 = L - NL1 layers starting at lo = NL1, tile 2^TL points
template <bool INV, int TL, int THREADS, bool ZTOP = false>
static int launch_second(u32 nl2, const u32* const* src, u32* const* dst, u32 L, u32 log_in, u32 lo, const u32* tree, u32 M,
                         u32 scale, size_t n_cols) {
#define CM_F4(N)                                                                                                          \
    case N:                                                                                                               \
        if constexpr (N >= 1 && (TL == 12 ? N <= 9 : (N >= 9 && N <= 11))) return launch_fft4<INV, N, TL - N, THREADS, ZTOP>(src, dst, L, log_in, lo, tree, M, scale, n_cols, 2); \
        break;
    switch (nl2) {
        CM_F4(1) CM_F4(2) CM_F4(3) CM_F4(4) CM_F4(5) CM_F4(6) CM_F4(7) CM_F4(8) CM_F4(9) CM_F4(10) CM_F4(11)
    }
#undef CM_F4
    set_error("fft4: unsupported second-pass shape");
    return -1;
}

// Returns 1 if this path does not cover (L, n_cols); 0 on success; <0 / cuda error otherwise.
template <bool INV>
int run_fft4(const u32* const* src, u32* const* dst, size_t n_cols, u32 L, u32 log_in, const u32* tree, u32 M, u32 scale_last) {
    if (L < 12 || L > 26 || log_in < 2) return 1;
    if (L >= 25) {
        // three passes: 13 contiguous layers, then two strided passes of 6-7 layers on 2^13-point tiles
        // with 256-512-byte runs (cfg5 sizes: 2^25, 2^26 points)
        const u32 nl2 = L - 19;  // 6 or 7
        auto p1 = [&](const u32* const* s_, u32 lin, u32 scale) { return launch_fft4<INV, 13, 0, 1024>(s_, dst, L, lin, 0, tree, M, scale, n_cols, 3); };
        auto p2 = [&](const u32* const* s_, u32 lin, u32 scale) {
            return nl2 == 6 ? launch_fft4<INV, 6, 7, 1024>(s_, dst, L, lin, 13, tree, M, scale, n_cols, 3)
                            : launch_fft4<INV, 7, 6, 1024>(s_, dst, L, lin, 13, tree, M, scale, n_cols, 3);
        };
        auto p3 = [&](const u32* const* s_, u32 lin, u32 scale) {
            if constexpr (!INV)
                if (lin < L) return launch_fft4<INV, 6, 7, 1024, true>(s_, dst, L, lin, 13 + nl2, tree, M, scale, n_cols, 3);
            return launch_fft4<INV, 6, 7, 1024>(s_, dst, L, lin, 13 + nl2, tree, M, scale, n_cols, 3);
        };
        const u32* const* d = (const u32* const*)dst;
        if (INV) {
            if (int e = p1(src, log_in, 1)) return e;
            if (int e = p2(d, L, 1)) return e;
            return p3(d, L, scale_last);
        }
        if (int e = p3(src, log_in, 1)) return e;
        if (int e = p2(d, L, 1)) return e;
        return p1(d, L, 1);
    }
    if (L == 12) {
        if constexpr (!INV)
            if (log_in < L) return launch_fft4<INV, 12, 0, 512, true>(src, dst, L, log_in, 0, tree, M, scale_last, n_cols, 1);
        return launch_fft4<INV, 12, 0, 512>(src, dst, L, log_in, 0, tree, M, scale_last, n_cols, 1);
    }
    // two passes: the contiguous one covers layers [0, NL1), the strided one [NL1, L)
    const bool big = L > 21;  // second pass would fall below 32-byte runs with 2^12-point tiles
    const u32 nl1 = big ? 13 : 12;
    const u32 nl2 = L - nl1;
    const u32 min_run = big ? 2 : 3;
    if (nl2 > (big ? 13u : 12u) - min_run) return 1;
    auto first = [&](const u32* const* s, u32 lin, u32 scale) {
        return big ? launch_fft4<INV, 13, 0, 1024>(s, dst, L, lin, 0, tree, M, scale, n_cols, 2)
                   : launch_fft4<INV, 12, 0, 512>(s, dst, L, lin, 0, tree, M, scale, n_cols, 2);
    };
    auto second = [&](const u32* const* s, u32 lin, u32 scale) {
        if constexpr (!INV)
            if (lin < L)  // low-degree extension: the upper half of the coefficients is zero padding, the top layer is a copy
                return big ? launch_second<INV, 13, 1024, true>(nl2, s, dst, L, lin, nl1, tree, M, scale, n_cols)
                           : launch_second<INV, 12, 512, true>(nl2, s, dst, L, lin, nl1, tree, M, scale, n_cols);
        return big ? launch_second<INV, 13, 1024>(nl2, s, dst, L, lin, nl1, tree, M, scale, n_cols)
                   : launch_second<INV, 12, 512>(nl2, s, dst, L, lin, nl1, tree, M, scale, n_cols);
    };
    if (INV) {  // layers ascending: contiguous pass first, scale folded into the last store
        if (int e = first(src, log_in, 1)) return e;
        return second((const u32* const*)dst, L, scale_last);
    }
    // forward: layers descending, strided pass first (reads the zero-padded coefficients)
    if (int e = second(src, log_in, 1)) return e;
    return first((const u32* const*)dst, L, 1);
}

template int run_fft4<true>(const u32* const*, u32* const*, size_t, u32, u32, const u32*, u32, u32);
template int run_fft4<false>(const u32* const*, u32* const*, size_t, u32, u32, const u32*, u32, u32);

}  // namespace cm31

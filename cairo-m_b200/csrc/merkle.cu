// MerkleOps<Blake2sMerkleHasher>::commit_on_layer and GrindOps on sm_100a.
//
// Replaces external/stwo/crates/prover/src/core/backend/simd/blake2s.rs:60-142 (commit_on_layer),
// simd/grind.rs:21-71; defined by cpu/blake2s.rs:9-23 + vcs/blake2_merkle.rs:14-30 and
// cpu/grind.rs:5-16.
//
// One thread per Merkle node: consecutive threads read consecutive words of every column
// (coalesced 128 B per warp per column) and write 32 contiguous bytes each.  The message block
// lives in 16 registers and is refilled 16 columns at a time.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "blake2s.cuh"
#include "circle.hpp"
#include "common.cuh"

namespace cm31 {

// PREV_IN_FLIGHT: `prev` was written earlier in the SAME launch by other threads of this CTA (fused layers): its loads
// must stay coherent (ld.global.cg), never the read-only path.
template <bool PREV_IN_FLIGHT = false>
__device__ __forceinline__ void hash_node(size_t i, const u32* prev, const u32* const* __restrict__ cols, u32 n_cols, u32* out) {
    Blake2sState st;
    blake2s_init(st);
    u32 m[16];
    const u64 total = (prev ? 64ull : 0ull) + 4ull * n_cols;
    u64 done = 0;
    if (prev) {
        const uint4* p4 = reinterpret_cast<const uint4*>(prev + i * 16);
        uint4 a, b, c, d;
        if (PREV_IN_FLIGHT) {
            a = __ldcg(p4);
            b = __ldcg(p4 + 1);
            c = __ldcg(p4 + 2);
            d = __ldcg(p4 + 3);
        } else {
            a = p4[0];
            b = p4[1];
            c = p4[2];
            d = p4[3];
        }
        m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w;
        m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
        m[8] = c.x; m[9] = c.y; m[10] = c.z; m[11] = c.w;
        m[12] = d.x; m[13] = d.y; m[14] = d.z; m[15] = d.w;
        done = 64;
        blake2s_compress(st, m, done, done == total);
    }
    for (u32 c0 = 0; c0 < n_cols || (total == 0 && c0 == 0); c0 += 16) {
        u32 nb = min(16u, n_cols - c0);
#pragma unroll
        for (u32 k = 0; k < 16; k++) m[k] = (k < nb) ? __ldg(cols[c0 + k] + i) : 0u;
        done += 4ull * nb;
        blake2s_compress(st, m, done, done == total);
        if (total == 0) break;
    }
    uint4* o4 = reinterpret_cast<uint4*>(out + i * 8);
    o4[0] = make_uint4(st.h[0], st.h[1], st.h[2], st.h[3]);
    o4[1] = make_uint4(st.h[4], st.h[5], st.h[6], st.h[7]);
}

__global__ void __launch_bounds__(256) merkle_layer_kernel(u32 log_size, const u32* __restrict__ prev,
                                                           const u32* const* __restrict__ cols, u32 n_cols,
                                                           u32* __restrict__ out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << log_size)) return;
    hash_node(i, prev, cols, n_cols, out);
}

// ---- one node hashed by FOUR lanes.  The layers near the root that receive the columns of the 16-row padding components
// (~500 trace / ~1000 interaction columns land in the 32-node layer) are one Merkle-Damgard chain of 30-65 dependent
// compressions per node: with a thread per node a single warp issues ~1300 dependent-ish instructions per compression
// (profiles/full_merkle_top_kernel_r01h.md: 219-264 us, one warp active).  Here lane q of a quad holds column q of the 4x4
// Blake2s state (v[q], v[4+q], v[8+q], v[12+q]): the four column G's of a round run in parallel across the quad, three
// shuffles rotate b, c, d into the diagonals and back, the message block sits in shared memory (16 words per node) and the
// next block's column words are loaded while the current one is compressed.
__constant__ uint8_t CM_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
__constant__ u32 CM_BLAKE_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};

// Lane q's four message indices of round r, packed 4 bits each (column G: x, y; diagonal G: x, y).  The table is built once
// per CTA in shared memory (a constant-memory table indexed by q serialises the quad's lanes on every access).
__device__ __forceinline__ void quad_sigma_init(u32 (*tab)[10]) {  // first 40 threads of the CTA; followed by __syncthreads
    if (threadIdx.x < 40) {
        const u32 q = threadIdx.x / 10, r = threadIdx.x % 10;
        const uint8_t* sg = CM_SIGMA[r];
        tab[q][r] = (u32)sg[2 * q] | ((u32)sg[2 * q + 1] << 4) | ((u32)sg[8 + 2 * q] << 8) | ((u32)sg[9 + 2 * q] << 12);
    }
}
__device__ __forceinline__ void quad_compress(u32& h_lo, u32& h_hi, const u32* msg, u32 q, const u32* sig, u64 t, bool last, u32 mask) {
    u32 a = h_lo, b = h_hi, c = CM_BLAKE_IV[q], d = CM_BLAKE_IV[4 + q];
    if (q == 0) d ^= (u32)t;
    else if (q == 1) d ^= (u32)(t >> 32);
    else if (q == 2 && last) d = ~d;
    const u32 l1 = (q + 1) & 3, l2 = (q + 2) & 3, l3 = (q + 3) & 3;
    // the round's four message words are read from shared memory one round ahead (volatile: the compiler must not hoist all
    // forty reads to the top and spill them -- the kernel is capped at 64 registers by its 1024-thread launch bound)
    const volatile u32* vm = msg;
    u32 sg = sig[0];
    u32 x0 = vm[sg & 15u], y0 = vm[(sg >> 4) & 15u], x1 = vm[(sg >> 8) & 15u], y1 = vm[(sg >> 12) & 15u];
#pragma unroll
    for (int r = 0; r < 10; r++) {
        u32 nx0 = 0, ny0 = 0, nx1 = 0, ny1 = 0;
        if (r < 9) {
            sg = sig[r + 1];
            nx0 = vm[sg & 15u], ny0 = vm[(sg >> 4) & 15u], nx1 = vm[(sg >> 8) & 15u], ny1 = vm[(sg >> 12) & 15u];
        }
        CM_G(a, b, c, d, x0, y0)
        b = __shfl_sync(mask, b, l1, 4);
        c = __shfl_sync(mask, c, l2, 4);
        d = __shfl_sync(mask, d, l3, 4);
        CM_G(a, b, c, d, x1, y1)
        b = __shfl_sync(mask, b, l3, 4);
        c = __shfl_sync(mask, c, l2, 4);
        d = __shfl_sync(mask, d, l1, 4);
        x0 = nx0, y0 = ny0, x1 = nx1, y1 = ny1;
    }
    h_lo ^= a ^ c;
    h_hi ^= b ^ d;
}
// node `i` by the quad of lanes (q = lane & 3); `mask` = the lanes of this warp that take part (whole quads).
// msg = this node's 16 shared-memory words.  Column words are prefetched one block ahead and the column POINTERS two
// blocks ahead, so neither of the two dependent global loads is ever waited for between compressions.
__device__ __noinline__ void hash_node_quad(size_t i, const u32* prev, const u32* const* __restrict__ cols, u32 n_cols, u32* out,
                                            u32* msg, u32 q, const u32 (*sigtab)[10], u32 mask) {
    const u32* sig = sigtab[q];  // shared memory, read one round ahead
    u32 h_lo = CM_BLAKE_IV[q] ^ (q == 0 ? 0x01010020u : 0u), h_hi = CM_BLAKE_IV[4 + q];
    const u64 total = (prev ? 64ull : 0ull) + 4ull * n_cols;
    u64 done = 0;
    u32 nxt[4];
    const u32* nptr[4];  // pointers of the columns of the block after the one in `nxt`
    auto load_ptrs = [&](u32 c0) {
#pragma unroll
        for (u32 k = 0; k < 4; k++) {
            const u32 col = c0 + 4 * q + k;
            nptr[k] = col < n_cols ? cols[col] : nullptr;
        }
    };
    auto load_words = [&]() {
#pragma unroll
        for (u32 k = 0; k < 4; k++) nxt[k] = nptr[k] ? __ldg(nptr[k] + i) : 0u;
    };
    u32 c0;    // first column of the block whose pointers sit in nptr
    u32 held;  // bytes of the block in nxt
    if (prev) {
        const uint4 w = __ldcg(reinterpret_cast<const uint4*>(prev + i * 16) + q);
        nxt[0] = w.x, nxt[1] = w.y, nxt[2] = w.z, nxt[3] = w.w;
        held = 64;
        c0 = 0;
        load_ptrs(0);
    } else {
        load_ptrs(0);
        load_words();
        held = 4u * min(16u, n_cols);
        c0 = 16;
        load_ptrs(16);
    }
    while (true) {
        __syncwarp(mask);  // the quad has finished reading the previous block
#pragma unroll
        for (u32 k = 0; k < 4; k++) msg[4 * q + k] = nxt[k];
        __syncwarp(mask);
        done += held;
        const bool last = done == total;
        if (!last) {  // words of the next block, pointers of the one after: both in flight while this block is compressed
            load_words();
            held = 4u * min(16u, n_cols - c0);
            c0 += 16;
            load_ptrs(c0);
        }
        quad_compress(h_lo, h_hi, msg, q, sig, done, last, mask);
        if (last) break;
    }
    out[i * 8 + q] = h_lo;
    out[i * 8 + 4 + q] = h_hi;
}

// nodes [first, first + count) of a layer: a rank's row range of a sharded tree (columns may be peer pointers)
__global__ void __launch_bounds__(256) merkle_layer_range_kernel(const u32* __restrict__ prev, const u32* const* __restrict__ cols, u32 n_cols,
                                                                 u32* __restrict__ out, size_t first, size_t count) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    hash_node(first + i, prev, cols, n_cols, out);
}

// The top of a tree (layers top_log .. 0, at most 2^10 nodes wide) in ONE single-CTA launch:
// these layers are pure latency (<= 1024 hashes each), a launch per layer costs more than the work.
// layer_out[l] = output of layer l; cols of layer l = cols[col_start[l] .. col_start[l+1]).
struct MerkleTopArgs {
    u32* layer_out[11];
    u32 col_start[12];
};
constexpr u32 QUAD_MAX_LOG = 8;  // 4 lanes per node, <= 1024 threads
__global__ void __launch_bounds__(1024) merkle_top_kernel(u32 top_log, const u32* prev, const u32* const* cols, MerkleTopArgs args, int use_quads) {
    __shared__ u32 quad_msg[(1u << QUAD_MAX_LOG) * 16];
    __shared__ u32 quad_sig[4][10];
    quad_sigma_init(quad_sig);
    __syncthreads();
    for (int l = (int)top_log; l >= 0; l--) {
        const u32 n_cols = args.col_start[l + 1] - args.col_start[l];
        if (use_quads && l <= (int)QUAD_MAX_LOG && (4u << l) <= blockDim.x && (prev != nullptr || n_cols != 0)) {
            if (threadIdx.x < (4u << l)) {
                const u32 mask = (4u << l) >= 32u ? 0xffffffffu : ((1u << (4u << l)) - 1u);  // partial warp near the root
                hash_node_quad(threadIdx.x >> 2, prev, cols + args.col_start[l], n_cols, args.layer_out[l], quad_msg + (threadIdx.x >> 2) * 16,
                               threadIdx.x & 3u, quad_sig, mask);
            }
        } else if (threadIdx.x < (1u << l)) {
            if (l == (int)top_log) hash_node<false>(threadIdx.x, prev, cols + args.col_start[l], n_cols, args.layer_out[l]);
            else hash_node<true>(threadIdx.x, prev, cols + args.col_start[l], n_cols, args.layer_out[l]);
        }
        prev = args.layer_out[l];
        __syncthreads();  // block-scope visibility of the layer just written
    }
}

// ------------------------------------------------------------------ FRI tail: the small inner layers in ONE launch
// FriProver::commit_inner_layers (fri.rs:226-266) below 2^FRI_TAIL_LOG points is a chain of tiny, strictly dependent steps:
// Merkle tree of the layer (blake2_merkle.rs), root -> channel.mix_root, channel.draw_secure_felt -> alpha (channel/blake2s.rs),
// fold_line (fri.rs:1132-1157) [+ fold_circle_into_line of a column that joins at this size, fri.rs:1159-1189].  Issued from the
// host that is 3-4 launches and one device->host round trip PER LAYER (~60-85 us each, ~1 ms per proof for the last dozen
// layers).  Here a single CTA runs the whole chain, Fiat-Shamir included: the Blake2s channel lives in shared memory, the host
// gets the roots back and replays mix_root / draw on its own channel (microseconds) so both transcripts stay identical.
struct FriTailLayerDev {
    const u32* in[4];
    u32* out[4];
    u32* lvl[16];         // lvl[j] = hash layer with 2^j nodes of this layer's tree
    const u32* circ[4];   // circle evaluation (log = this layer's log) folded into `out`, or null
    const u32* itw_line;  // 1/x twiddles of the line domain
    const u32* itw_circ;  // first line layer pairs the 1/y twiddles are derived from (circle log > 2)
    u32 log, has_circ, iy0, iy1;
};
__device__ __forceinline__ QM31 ldcg4(const u32* const* c, size_t i) { return qm_make(__ldcg(c[0] + i), __ldcg(c[1] + i), __ldcg(c[2] + i), __ldcg(c[3] + i)); }

__global__ void __launch_bounds__(1024) fri_tail_kernel(const FriTailLayerDev* __restrict__ layers, u32 n_layers, const u32* __restrict__ digest_in,
                                                        u32* __restrict__ roots_out) {
    __shared__ u32 sh_digest[8];
    __shared__ u32 sh_alpha[4];
    __shared__ u32 quad_msg[(1u << QUAD_MAX_LOG) * 16];
    __shared__ u32 quad_sig[4][10];
    const u32 tid = threadIdx.x;
    if (tid < 8) sh_digest[tid] = digest_in[tid];
    quad_sigma_init(quad_sig);
    __syncthreads();
    for (u32 li = 0; li < n_layers; li++) {
        const FriTailLayerDev& L = layers[li];
        const u32 k = L.log;
        // ---- leaves: Blake2s of the 4 coordinate words (one 16-byte block)
        for (u32 n = tid; n < (1u << k); n += blockDim.x) {
            Blake2sState st;
            blake2s_init(st);
            u32 m[16];
#pragma unroll
            for (int c = 0; c < 4; c++) m[c] = __ldcg(L.in[c] + n);
#pragma unroll
            for (int c = 4; c < 16; c++) m[c] = 0;
            blake2s_compress(st, m, 16, true);
            uint4* o = reinterpret_cast<uint4*>(L.lvl[k] + (size_t)n * 8);
            o[0] = make_uint4(st.h[0], st.h[1], st.h[2], st.h[3]);
            o[1] = make_uint4(st.h[4], st.h[5], st.h[6], st.h[7]);
        }
        __syncthreads();
        // ---- inner levels
        for (int j = (int)k - 1; j >= 0; j--) {
            if (j <= (int)QUAD_MAX_LOG) {  // few nodes: 4 lanes per node (latency of one compression / ~3), as in merkle_top_kernel
                if (tid < (4u << j)) {
                    const u32 mask = (4u << j) >= 32u ? 0xffffffffu : ((1u << (4u << j)) - 1u);
                    hash_node_quad(tid >> 2, L.lvl[j + 1], nullptr, 0, L.lvl[j], quad_msg + (tid >> 2) * 16, tid & 3u, quad_sig, mask);
                }
            } else {
                for (u32 n = tid; n < (1u << j); n += blockDim.x) hash_node<true>(n, L.lvl[j + 1], nullptr, 0, L.lvl[j]);
            }
            __syncthreads();
        }
        // ---- Fiat-Shamir: mix_root, draw_secure_felt (channel/blake2s.rs:60-116)
        if (tid == 0) {
            u32 m[16];
            Blake2sState st;
            blake2s_init(st);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                m[i] = sh_digest[i];
                const u32 r = __ldcg(L.lvl[0] + i);
                m[8 + i] = r;
                roots_out[li * 8 + i] = r;
            }
            blake2s_compress(st, m, 64, true);
#pragma unroll
            for (int i = 0; i < 8; i++) sh_digest[i] = st.h[i];
            for (u32 n_sent = 0;; n_sent++) {  // draw_base_felts: retry until all 8 words are below 2P
                Blake2sState d;
                blake2s_init(d);
#pragma unroll
                for (int i = 0; i < 8; i++) m[i] = sh_digest[i];
                m[8] = n_sent;
#pragma unroll
                for (int i = 9; i < 16; i++) m[i] = 0;
                blake2s_compress(d, m, 36, true);
                bool ok = true;
#pragma unroll
                for (int i = 0; i < 8; i++) ok = ok && d.h[i] < 2 * P;
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 4; i++) sh_alpha[i] = d.h[i] >= P ? d.h[i] - P : d.h[i];
                    break;
                }
            }
        }
        __syncthreads();
        // ---- fold_line (+ fold_circle_into_line with the same alpha)
        const QM31 alpha = qm_make(sh_alpha[0], sh_alpha[1], sh_alpha[2], sh_alpha[3]);
        const QM31 alpha_sq = qm_sqr(alpha);
        for (u32 n = tid; n < (1u << (k - 1)); n += blockDim.x) {
            const QM31 f0 = ldcg4(L.in, 2 * (size_t)n), f1 = ldcg4(L.in, 2 * (size_t)n + 1);
            const u32 it = __ldg(L.itw_line + n);
            QM31 v = qm_add(qm_add(f0, f1), qm_mul(alpha, qm_mul_m31(qm_sub(f0, f1), it)));
            if (L.has_circ) {
                const QM31 c0 = qm_make(__ldg(L.circ[0] + 2 * n), __ldg(L.circ[1] + 2 * n), __ldg(L.circ[2] + 2 * n), __ldg(L.circ[3] + 2 * n));
                const QM31 c1 = qm_make(__ldg(L.circ[0] + 2 * n + 1), __ldg(L.circ[1] + 2 * n + 1), __ldg(L.circ[2] + 2 * n + 1), __ldg(L.circ[3] + 2 * n + 1));
                u32 ity;
                if (k <= 2) {
                    ity = n == 0 ? L.iy0 : L.iy1;
                } else {
                    const u32 x = __ldg(L.itw_circ + 2 * (n >> 2)), y = __ldg(L.itw_circ + 2 * (n >> 2) + 1);
                    const u32 sel = n & 3;
                    const u32 w = sel < 2 ? y : x;
                    ity = (sel == 1 || sel == 2) ? m31_neg(w) : w;
                }
                const QM31 fp = qm_add(qm_mul(alpha, qm_mul_m31(qm_sub(c0, c1), ity)), qm_add(c0, c1));
                v = qm_add(qm_mul(v, alpha_sq), fp);
            }
            L.out[0][n] = v.a;
            L.out[1][n] = v.b;
            L.out[2][n] = v.c;
            L.out[3][n] = v.d;
        }
        __syncthreads();
    }
}

// Layers log_size, log_size-1, .., log_size-n_levels+1 in ONE launch, when only the first of them carries columns (FRI
// layer trees, the composition tree, the gaps between column sizes of a trace tree): a CTA hashes 256 nodes of the
// first layer and keeps halving inside the block, so a tree costs a few launches instead of one per layer.
struct MerkleMultiArgs {
    u32* layer_out[9];
};
// `first`: first node of the launch's range in the first layer (a multiple of 256; 0 for a whole layer)
__global__ void __launch_bounds__(256) merkle_multi_kernel(u32 log_size, const u32* prev, const u32* const* cols, u32 n_cols, u32 n_levels,
                                                           MerkleMultiArgs args, size_t first) {
    const size_t i = first + blockIdx.x * (size_t)256 + threadIdx.x;
    if (i < ((size_t)1 << log_size)) hash_node<false>(i, prev, cols, n_cols, args.layer_out[0]);
    for (u32 lvl = 1; lvl < n_levels; lvl++) {
        __syncthreads();  // the children written by this CTA are visible to it
        const u32 cnt = 256u >> lvl;
        if (threadIdx.x < cnt) hash_node<true>((first >> lvl) + blockIdx.x * (size_t)cnt + threadIdx.x, args.layer_out[lvl - 1], nullptr, 0, args.layer_out[lvl]);
    }
}

// The same fusion for LARGE first layers (>= 2^19 nodes), without block barriers: a WARP owns 32 * PER_LANE consecutive nodes of
// the first layer and the whole subtree above them that keeps all 32 lanes busy (PER_LANE = 8: 256 -> 128 -> 64 -> 32 nodes,
// 15 full-warp compressions for 480 nodes).  Lanes exchange children through the layer buffers themselves (they are
// outputs anyway): __syncwarp orders the stores, the re-reads bypass L1.  ncu r01d on the CTA-wide form: 55 % of the
// stall samples were `barrier` -- after the first level 128, 64, 32, .. of 256 threads hash while the CTA waits; here no
// warp ever waits for another, so the schedulers always find an eligible warp among the resident ones.
template <int PER_LANE>
__global__ void __launch_bounds__(256) merkle_warp_kernel(u32 log_size, const u32* prev, const u32* const* cols, u32 n_cols, u32 n_levels,
                                                          MerkleMultiArgs args, size_t first, size_t count) {
    const u32 lane = threadIdx.x & 31u;
    size_t base = ((blockIdx.x * (size_t)256 + threadIdx.x) >> 5) * (32u * PER_LANE);
    if (base >= count) return;  // whole warps only: the range is a multiple of 32 * PER_LANE
    base += first;
#pragma unroll 1
    for (u32 j = 0; j < (u32)PER_LANE; j++) hash_node<false>(base + j * 32u + lane, prev, cols, n_cols, args.layer_out[0]);
    u32 cnt = 32u * PER_LANE;
#pragma unroll 1
    for (u32 lvl = 1; lvl < n_levels; lvl++) {
        __syncwarp();  // the children this warp wrote are visible to all of its lanes
        cnt >>= 1;
        base >>= 1;
#pragma unroll 1
        for (u32 k = lane; k < cnt; k += 32u) hash_node<true>(base + k, args.layer_out[lvl - 1], nullptr, 0, args.layer_out[lvl]);
    }
}

// grind: thread t tests nonce = base + t; result = atomicMin over matching nonces.
__global__ void grind_kernel(const u32* __restrict__ digest, u32 pow_bits, u64 base, unsigned long long* best) {
    u64 nonce = base + blockIdx.x * (u64)blockDim.x + threadIdx.x;
    Blake2sState st;
    blake2s_init(st);
    u32 m[16];
#pragma unroll
    for (int k = 0; k < 8; k++) m[k] = digest[k];
    m[8] = (u32)nonce;
    m[9] = (u32)(nonce >> 32);
#pragma unroll
    for (int k = 10; k < 16; k++) m[k] = 0;
    blake2s_compress(st, m, 40, true);
    // trailing zeros of the digest read as a little-endian u128 (channel/blake2s.rs:57-59)
    u32 tz = 0;
    if (st.h[0]) tz = __ffs(st.h[0]) - 1;
    else if (st.h[1]) tz = 32 + __ffs(st.h[1]) - 1;
    else if (st.h[2]) tz = 64 + __ffs(st.h[2]) - 1;
    else if (st.h[3]) tz = 96 + __ffs(st.h[3]) - 1;
    else tz = 128;
    if (tz >= pow_bits) atomicMin(best, (unsigned long long)nonce);
}

}  // namespace cm31

using namespace cm31;

extern "C" {

int cm31_blake2s_commit_layer(uint32_t log_size, const uint32_t* prev_layer, const uint32_t* const* cols,
                              size_t n_cols, uint32_t* out_layer) {
    CM_REQUIRE(out_layer != nullptr, "commit_layer: null output");
    CM_REQUIRE(log_size <= 30, "commit_layer: layer too large");
    DeviceTable dcols;
    if (n_cols != 0)  // inner layers carry no columns: nothing to upload
        if (int e = dcols.upload(cols, n_cols * sizeof(void*))) return e;
    size_t n = (size_t)1 << log_size;
    unsigned threads = n < 256 ? (unsigned)((n + 31) / 32 * 32) : 256;
    ProfScope prof(n_cols ? "merkle_leaf_layer" : "merkle_inner_layer", (4ull * n_cols + 32ull + (prev_layer ? 64ull : 0ull)) * n);
    merkle_layer_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream()>>>(
        log_size, prev_layer, (const u32* const*)dcols.d, (u32)n_cols, out_layer);
    CM_LAUNCH_CHECK();
    return 0;
}

int cm31_blake2s_commit_layer_range(uint32_t log_size, const uint32_t* prev_layer, const uint32_t* const* cols, size_t n_cols,
                                    uint32_t* out_layer, size_t first_node, size_t n_nodes) {
    CM_REQUIRE(out_layer != nullptr && log_size <= 30, "commit_layer_range: bad arguments");
    CM_REQUIRE(first_node + n_nodes <= ((size_t)1 << log_size), "commit_layer_range: range outside the layer");
    if (n_nodes == 0) return 0;
    DeviceTable dcols;
    if (n_cols != 0)
        if (int e = dcols.upload(cols, n_cols * sizeof(void*))) return e;
    ProfScope prof(n_cols ? "merkle_leaf_layer" : "merkle_inner_layer", (4ull * n_cols + 32ull + (prev_layer ? 64ull : 0ull)) * n_nodes);
    merkle_layer_range_kernel<<<(unsigned)((n_nodes + 255) / 256), 256, 0, stream()>>>(prev_layer, (const u32* const*)dcols.d, (u32)n_cols, out_layer,
                                                                                       first_node, n_nodes);
    CM_LAUNCH_CHECK();
    return 0;
}

int cm31_blake2s_commit_multi(uint32_t log_size, const uint32_t* prev_layer, const uint32_t* const* cols, size_t n_cols,
                              uint32_t n_levels, uint32_t* const* out_layers) {
    return cm31_blake2s_commit_multi_range(log_size, prev_layer, cols, n_cols, n_levels, out_layers, 0, (size_t)1 << log_size);
}

// nodes [first_node, first_node + n_nodes) of layer log_size and the whole subtree above them in the next n_levels - 1
// layers (a rank's share of a striped tree; the whole layer for first_node = 0, n_nodes = 2^log_size)
int cm31_blake2s_commit_multi_range(uint32_t log_size, const uint32_t* prev_layer, const uint32_t* const* cols, size_t n_cols,
                                    uint32_t n_levels, uint32_t* const* out_layers, size_t first_node, size_t n_nodes) {
    CM_REQUIRE(out_layers != nullptr && n_levels >= 1 && n_levels <= 9, "commit_multi: 1..9 levels");
    CM_REQUIRE(log_size <= 30 && log_size >= 8 && log_size + 1 >= n_levels, "commit_multi: first layer needs at least 256 nodes");
    CM_REQUIRE(n_nodes >= 256 && (n_nodes & (n_nodes - 1)) == 0 && first_node % n_nodes == 0 && first_node + n_nodes <= ((size_t)1 << log_size),
               "commit_multi: the node range must be an aligned power of two of at least 256 nodes");
    MerkleMultiArgs args;
    for (u32 l = 0; l < 9; l++) args.layer_out[l] = l < n_levels ? out_layers[l] : nullptr;
    DeviceTable dcols;
    if (n_cols != 0)
        if (int e = dcols.upload(cols, n_cols * sizeof(void*))) return e;
    const size_t n = n_nodes;
    uint64_t bytes = (4ull * n_cols + 32ull + (prev_layer ? 64ull : 0ull)) * n;
    for (u32 l = 1; l < n_levels; l++) bytes += 96ull * (n >> l);
    if (n >= ((size_t)1 << 19) && n_levels >= 2 && !getenv("CM31_MERKLE_CTA_FUSION")) {
        // large layers: barrier-free warp subtrees, 4 (3) levels per launch, the remaining levels by further launches
        const u32 per_lane = n >= ((size_t)1 << 21) ? 8 : 4;
        const u32 lv = std::min<u32>(n_levels, per_lane == 8 ? 4 : 3);
        uint64_t b0 = (4ull * n_cols + 32ull + (prev_layer ? 64ull : 0ull)) * n;
        for (u32 l = 1; l < lv; l++) b0 += 96ull * (n >> l);
        {
            ProfScope prof(n_cols ? "merkle_leaf_layer" : "merkle_inner_layer", b0);
            const unsigned blocks = (unsigned)((n / (32 * per_lane) * 32 + 255) / 256);
            if (per_lane == 8) merkle_warp_kernel<8><<<blocks, 256, 0, stream()>>>(log_size, prev_layer, (const u32* const*)dcols.d, (u32)n_cols, lv, args, first_node, n);
            else merkle_warp_kernel<4><<<blocks, 256, 0, stream()>>>(log_size, prev_layer, (const u32* const*)dcols.d, (u32)n_cols, lv, args, first_node, n);
            CM_LAUNCH_CHECK();
        }
        if (lv == n_levels) return 0;
        if (n_levels - lv == 1) return cm31_blake2s_commit_layer_range(log_size - lv, out_layers[lv - 1], nullptr, 0, out_layers[lv], first_node >> lv, n >> lv);
        return cm31_blake2s_commit_multi_range(log_size - lv, out_layers[lv - 1], nullptr, 0, n_levels - lv, out_layers + lv, first_node >> lv, n >> lv);
    }
    ProfScope prof(n_cols ? "merkle_leaf_layer" : "merkle_inner_layer", bytes);
    merkle_multi_kernel<<<(unsigned)(n / 256), 256, 0, stream()>>>(log_size, prev_layer, (const u32* const*)dcols.d, (u32)n_cols, n_levels, args, first_node);
    CM_LAUNCH_CHECK();
    return 0;
}

int cm31_blake2s_commit_top(uint32_t top_log_size, const uint32_t* prev_layer, const uint32_t* const* cols,
                            const uint32_t* col_start_host, uint32_t* const* out_layers) {
    CM_REQUIRE(top_log_size <= 10, "commit_top: at most 2^10 nodes in the widest layer");
    CM_REQUIRE(out_layers != nullptr && col_start_host != nullptr, "commit_top: null argument");
    MerkleTopArgs args;
    for (u32 l = 0; l <= top_log_size; l++) args.layer_out[l] = out_layers[l];
    for (u32 l = 0; l <= top_log_size + 1; l++) args.col_start[l] = col_start_host[l];
    size_t n_cols = col_start_host[top_log_size + 1];
    DeviceTable dcols;
    if (int e = dcols.upload(cols, n_cols * sizeof(void*))) return e;
    uint64_t bytes = 0;
    for (u32 l = 0; l <= top_log_size; l++) bytes += (4ull * (args.col_start[l + 1] - args.col_start[l]) + 96ull) << l;
    ProfScope prof("merkle_top_layers", bytes);
    static const int use_quads = getenv("CM31_MERKLE_NO_QUADS") ? 0 : 1;
    unsigned threads = 1u << top_log_size < 32 ? 32 : 1u << top_log_size;
    if (use_quads) threads = std::max(threads, std::min(1024u, 4u << std::min(top_log_size, QUAD_MAX_LOG)));
    if (threads < 64) threads = 64;  // the sigma table is built by the first 40 threads
    merkle_top_kernel<<<1, threads, 0, stream()>>>(top_log_size, prev_layer, (const u32* const*)dcols.d, args, use_quads);
    CM_LAUNCH_CHECK();
    return 0;
}

int cm31_fri_tail(const uint32_t digest_in[8], const cm31_fri_tail_layer* layers, size_t n_layers, const cm31_twiddles* tw,
                  uint32_t* roots_out_host) {
    if (n_layers == 0) return 0;
    CM_REQUIRE(digest_in != nullptr && layers != nullptr && tw != nullptr && roots_out_host != nullptr, "fri_tail: null argument");
    CM_REQUIRE(n_layers <= 16, "fri_tail: too many layers");
    static thread_local std::vector<FriTailLayerDev> host;
    host.assign(n_layers, FriTailLayerDev());
    uint64_t bytes = 0;
    for (size_t i = 0; i < n_layers; i++) {
        const cm31_fri_tail_layer& l = layers[i];
        CM_REQUIRE(l.log_size >= 1 && l.log_size <= 15 && l.log_size + 1 <= tw->log_size, "fri_tail: bad layer size");
        CM_REQUIRE(i == 0 || l.log_size + 1 == layers[i - 1].log_size, "fri_tail: layers must halve");
        CM_REQUIRE(l.tree_levels != nullptr, "fri_tail: null tree level table");
        for (int k = 0; k < 4; k++) CM_REQUIRE(l.ev_in[k] != nullptr && l.ev_out[k] != nullptr, "fri_tail: null evaluation column");
        for (u32 j = 0; j <= l.log_size; j++) CM_REQUIRE(l.tree_levels[j] != nullptr, "fri_tail: null tree level");
        CM_REQUIRE((l.circle[0] == nullptr) == (l.circle[1] == nullptr) && (l.circle[0] == nullptr) == (l.circle[2] == nullptr) &&
                       (l.circle[0] == nullptr) == (l.circle[3] == nullptr),
                   "fri_tail: a joining circle evaluation has 4 coordinate columns");
        FriTailLayerDev& d = host[i];
        for (int k = 0; k < 4; k++) {
            d.in[k] = l.ev_in[k];
            d.out[k] = l.ev_out[k];
            d.circ[k] = l.circle[k];
        }
        for (u32 j = 0; j <= l.log_size; j++) d.lvl[j] = l.tree_levels[j];
        d.log = l.log_size;
        d.has_circ = l.circle[0] != nullptr;
        // twiddle tree levels as in cm31_fold_line / cm31_fold_circle_into_line (fri.cu): the inverse-twiddle level whose coset
        // is half_odds(c) starts at 2^(M-1) - 2^c
        d.itw_line = tw->itw + (((size_t)1 << (tw->log_size - 1)) - ((size_t)1 << l.log_size));
        d.itw_circ = nullptr;
        d.iy0 = d.iy1 = 0;
        if (d.has_circ) {  // the circle evaluation has log size l.log_size and folds into 2^(log_size-1) line points
            if (l.log_size <= 2) {
                CircleDomain dom = CanonicCoset(l.log_size).circle_domain();
                d.iy0 = m31_inv(dom.at(bit_reverse(0, l.log_size)).y);
                d.iy1 = l.log_size == 2 ? m31_inv(dom.at(bit_reverse(2, l.log_size)).y) : 0;
            } else {
                d.itw_circ = tw->itw + (((size_t)1 << (tw->log_size - 1)) - ((size_t)1 << (l.log_size - 1)));
            }
        }
        bytes += (16ull + 96ull + 24ull) << l.log_size;
    }
    DeviceTable dl, dd;
    if (int e = dl.upload(host.data(), n_layers * sizeof(FriTailLayerDev))) return e;
    if (int e = dd.upload(digest_in, 32)) return e;
    u32* droots = nullptr;
    CM_CUDA(cudaMallocAsync(&droots, n_layers * 32, stream()));
    {
        ProfScope prof("fri_tail", bytes);
        fri_tail_kernel<<<1, 1024, 0, stream()>>>((const FriTailLayerDev*)dl.d, (u32)n_layers, (const u32*)dd.d, droots);
        CM_LAUNCH_CHECK();
    }
    CM_CUDA(cudaMemcpyAsync(roots_out_host, droots, n_layers * 32, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaFreeAsync(droots, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

int cm31_grind_blake2s(const uint8_t digest[32], uint32_t pow_bits, uint64_t* nonce_out) {
    CM_REQUIRE(nonce_out != nullptr, "grind: null output");
    CM_REQUIRE(pow_bits <= 64, "grind: pow_bits too large");
    DeviceTable dd;
    if (int e = dd.upload(digest, 32)) return e;
    unsigned long long* dbest = nullptr;
    CM_CUDA(cudaMallocAsync(&dbest, 8, stream()));
    const unsigned long long none = ~0ull;
    // batches of nonces in ascending order; the first batch with a hit contains the minimum.
    u64 batch = 1ull << (pow_bits + 2 < 16 ? 16 : (pow_bits + 2 > 26 ? 26 : pow_bits + 2));
    for (u64 base = 0;; base += batch) {
        CM_CUDA(cudaMemcpyAsync(dbest, &none, 8, cudaMemcpyHostToDevice, stream()));
        {
            ProfScope prof("grind", 0);
            grind_kernel<<<(unsigned)(batch / 256), 256, 0, stream()>>>((const u32*)dd.d, pow_bits, base, dbest);
        }
        CM_LAUNCH_CHECK();
        unsigned long long got = none;
        CM_CUDA(cudaMemcpyAsync(&got, dbest, 8, cudaMemcpyDeviceToHost, stream()));
        CM_CUDA(cudaStreamSynchronize(stream()));
        if (got != none) {
            *nonce_out = got;
            break;
        }
    }
    CM_CUDA(cudaFreeAsync(dbest, stream()));
    return 0;
}

}  // extern "C"

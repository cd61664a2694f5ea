// MerkleOps<Blake2sMerkleHasher>::commit_on_layer and GrindOps on sm_100a.
//
// Replaces external/stwo/crates/prover/src/core/backend/simd/blake2s.rs:60-142 (commit_on_layer),
// simd/grind.rs:21-71; defined by cpu/blake2s.rs:9-23 + vcs/blake2_merkle.rs:14-30 and
// cpu/grind.rs:5-16.
//
// One thread per Merkle node: consecutive threads read consecutive words of every column
// (coalesced 128 B per warp per column) and write 32 contiguous bytes each.  The message block
// lives in 16 registers and is refilled 16 columns at a time.
#include "blake2s.cuh"
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

// The top of a tree (layers top_log .. 0, at most 2^10 nodes wide) in ONE single-CTA launch:
// these layers are pure latency (<= 1024 hashes each), a launch per layer costs more than the work.
// layer_out[l] = output of layer l; cols of layer l = cols[col_start[l] .. col_start[l+1]).
struct MerkleTopArgs {
    u32* layer_out[11];
    u32 col_start[12];
};
__global__ void __launch_bounds__(1024) merkle_top_kernel(u32 top_log, const u32* prev, const u32* const* cols, MerkleTopArgs args) {
    for (int l = (int)top_log; l >= 0; l--) {
        if (threadIdx.x < (1u << l)) {
            if (l == (int)top_log) hash_node<false>(threadIdx.x, prev, cols + args.col_start[l], args.col_start[l + 1] - args.col_start[l], args.layer_out[l]);
            else hash_node<true>(threadIdx.x, prev, cols + args.col_start[l], args.col_start[l + 1] - args.col_start[l], args.layer_out[l]);
        }
        prev = args.layer_out[l];
        __syncthreads();  // block-scope visibility of the layer just written
    }
}

// Layers log_size, log_size-1, .., log_size-n_levels+1 in ONE launch, when only the first of them carries columns (FRI
// layer trees, the composition tree, the gaps between column sizes of a trace tree): a CTA hashes 256 nodes of the
// first layer and keeps halving inside the block, so a tree costs a few launches instead of one per layer.
struct MerkleMultiArgs {
    u32* layer_out[9];
};
__global__ void __launch_bounds__(256) merkle_multi_kernel(u32 log_size, const u32* prev, const u32* const* cols, u32 n_cols, u32 n_levels,
                                                           MerkleMultiArgs args) {
    const size_t i = blockIdx.x * (size_t)256 + threadIdx.x;
    if (i < ((size_t)1 << log_size)) hash_node<false>(i, prev, cols, n_cols, args.layer_out[0]);
    for (u32 lvl = 1; lvl < n_levels; lvl++) {
        __syncthreads();  // the children written by this CTA are visible to it
        const u32 cnt = 256u >> lvl;
        if (threadIdx.x < cnt) hash_node<true>(blockIdx.x * (size_t)cnt + threadIdx.x, args.layer_out[lvl - 1], nullptr, 0, args.layer_out[lvl]);
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

int cm31_blake2s_commit_multi(uint32_t log_size, const uint32_t* prev_layer, const uint32_t* const* cols, size_t n_cols,
                              uint32_t n_levels, uint32_t* const* out_layers) {
    CM_REQUIRE(out_layers != nullptr && n_levels >= 1 && n_levels <= 9, "commit_multi: 1..9 levels");
    CM_REQUIRE(log_size <= 30 && log_size >= 8 && log_size + 1 >= n_levels, "commit_multi: first layer needs at least 256 nodes");
    MerkleMultiArgs args;
    for (u32 l = 0; l < 9; l++) args.layer_out[l] = l < n_levels ? out_layers[l] : nullptr;
    DeviceTable dcols;
    if (n_cols != 0)
        if (int e = dcols.upload(cols, n_cols * sizeof(void*))) return e;
    size_t n = (size_t)1 << log_size;
    uint64_t bytes = (4ull * n_cols + 32ull + (prev_layer ? 64ull : 0ull)) * n;
    for (u32 l = 1; l < n_levels; l++) bytes += 96ull * (n >> l);
    ProfScope prof(n_cols ? "merkle_leaf_layer" : "merkle_inner_layer", bytes);
    merkle_multi_kernel<<<(unsigned)(n / 256), 256, 0, stream()>>>(log_size, prev_layer, (const u32* const*)dcols.d, (u32)n_cols, n_levels, args);
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
    merkle_top_kernel<<<1, 1u << top_log_size < 32 ? 32 : 1u << top_log_size, 0, stream()>>>(top_log_size, prev_layer, (const u32* const*)dcols.d, args);
    CM_LAUNCH_CHECK();
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

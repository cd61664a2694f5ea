// Single-proof sharding across the GPUs of one node (BASELINE north star: "columns shard by
// component/column-range across the 8 GPUs with NCCL over NVLink only for the single all-gather at Merkle-root/FRI-commit").
//
// One process per GPU; every rank runs the same protocol driver (same transcript, same allocation sequence) and owns a
// subset of the components.  The pieces here:
//   * an ARENA: one cudaMalloc'ed region per rank, exported with cudaIpc and mapped by every peer.  While sharding is on,
//     cm31_malloc bump-allocates from it (cm31_free is a no-op, the arena is reset per proof).  Every rank performs the same
//     allocation sequence, so a buffer lives at the SAME OFFSET on every rank: the peer's copy of any buffer is
//     peer_base[r] + (p - my_base) -- no handle or offset exchange per buffer.  Kernels (Merkle leaves, DEEP quotients,
//     gathers) simply receive peer pointers for columns another rank owns and read them over NVLink/NVSwitch while they
//     compute.
//   * an NCCL communicator for the small exchanges (barriers, 32-byte-per-node layer joins, claimed sums, OODS values,
//     multiplicity bins) and the in-place all-gathers of row-sharded results.
//   * reduce_m31: mod-P sum of the composition accumulators over ranks as reduce-scatter (peer loads of my row range) +
//     all-gather -- NCCL's integer sum would overflow u32 (8 x 2^31).
// Reference anchors for what is exchanged: vcs/prover.rs:50-64 (a Merkle leaf hashes one row of ALL columns),
// air/accumulation.rs:49-58 (per-log-size sums), backend/cpu/quotients.rs:28-38 (a quotient row reads all columns).
#include <nccl.h>

#include <cstring>

#include "common.cuh"

namespace cm31 {

struct ShardState {
    bool on = false;
    bool arena_active = false;  // cm31_malloc draws from the arena (between cm31_shard_arena_reset and cm31_shard_arena_suspend)
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0, arena_pos = 0, arena_peak = 0;
    uint8_t* peer_base[16] = {};
    uint64_t bytes_gathered = 0, bytes_reduced = 0, n_collectives = 0;
};
static ShardState g_sh;

#define CM_NCCL(expr)                                                                              \
    do {                                                                                           \
        ncclResult_t _r = (expr);                                                                  \
        if (_r != ncclSuccess) {                                                                   \
            cm31::set_error(std::string(#expr) + ": " + ncclGetErrorString(_r));                   \
            return -3;                                                                             \
        }                                                                                          \
    } while (0)

// called by cm31_malloc / cm31_free (runtime.cu)
bool shard_arena_alloc(void** out, size_t bytes, int* status) {
    if (!g_sh.on || !g_sh.arena_active) return false;
    size_t need = (bytes + 255) & ~(size_t)255;
    if (g_sh.arena_pos + need > g_sh.arena_bytes) {
        set_error("cm31: shard arena exhausted (" + std::to_string(g_sh.arena_bytes >> 20) + " MiB): raise the arena size given to cm31_shard_init");
        *status = -1;
        return true;
    }
    *out = g_sh.arena + g_sh.arena_pos;
    g_sh.arena_pos += need;
    if (g_sh.arena_pos > g_sh.arena_peak) g_sh.arena_peak = g_sh.arena_pos;
    *status = 0;
    return true;
}
bool shard_arena_owns(const void* p) {
    return g_sh.on && (const uint8_t*)p >= g_sh.arena && (const uint8_t*)p < g_sh.arena + g_sh.arena_bytes;
}

// dst[k][row] = sum over ranks of src_r[k][row] mod P, rows [first, first + count)
struct PeerCols4 {
    const u32* p[16][4];
};
__global__ void __launch_bounds__(256) reduce_m31_kernel(PeerCols4 src, int world, u32* d0, u32* d1, u32* d2, u32* d3, size_t first, size_t count) {
    size_t i = blockIdx.x * (size_t)256 + threadIdx.x;
    if (i >= count) return;
    const size_t row = first + i;
    u32 acc[4] = {0, 0, 0, 0};
    for (int r = 0; r < world; r++) {
#pragma unroll
        for (int k = 0; k < 4; k++) acc[k] = m31_add(acc[k], src.p[r][k][row]);
    }
    d0[row] = acc[0];
    d1[row] = acc[1];
    d2[row] = acc[2];
    d3[row] = acc[3];
}

}  // namespace cm31

using namespace cm31;

extern "C" {

int cm31_shard_unique_id(uint8_t out[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    CM_REQUIRE(out != nullptr, "shard_unique_id: null output");
    ncclUniqueId id;
    CM_NCCL(ncclGetUniqueId(&id));
    memcpy(out, &id, 128);
    return 0;
}

int cm31_shard_init(int rank, int world, const uint8_t unique_id[128], size_t arena_bytes) {
    CM_REQUIRE(!g_sh.on, "shard_init: already initialised");
    CM_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world, "shard_init: bad rank / world size");
    CM_REQUIRE((world & (world - 1)) == 0, "shard_init: the world size must be a power of two (row ranges of 2^k-sized layers)");
    CM_REQUIRE(unique_id != nullptr && arena_bytes >= ((size_t)1 << 20), "shard_init: bad arguments");
    ncclUniqueId id;
    memcpy(&id, unique_id, 128);
    CM_NCCL(ncclCommInitRank(&g_sh.comm, world, id, rank));
    g_sh.rank = rank;
    g_sh.world = world;
    CM_CUDA(cudaMalloc((void**)&g_sh.arena, arena_bytes));
    g_sh.arena_bytes = arena_bytes;
    g_sh.arena_pos = 0;
    // exchange the arena's IPC handle: all-gather of 64 bytes per rank
    cudaIpcMemHandle_t mine;
    CM_CUDA(cudaIpcGetMemHandle(&mine, g_sh.arena));
    uint8_t* dh = nullptr;
    CM_CUDA(cudaMalloc((void**)&dh, 64 * (size_t)world));
    CM_CUDA(cudaMemcpyAsync(dh + 64 * (size_t)rank, &mine, 64, cudaMemcpyHostToDevice, stream()));
    CM_NCCL(ncclAllGather(dh + 64 * (size_t)rank, dh, 64, ncclUint8, g_sh.comm, stream()));
    std::vector<uint8_t> handles(64 * (size_t)world);
    CM_CUDA(cudaMemcpyAsync(handles.data(), dh, handles.size(), cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    CM_CUDA(cudaFree(dh));
    for (int r = 0; r < world; r++) {
        if (r == rank) {
            g_sh.peer_base[r] = g_sh.arena;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles.data() + 64 * (size_t)r, 64);
        CM_CUDA(cudaIpcOpenMemHandle((void**)&g_sh.peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    g_sh.on = true;
    return 0;
}

int cm31_shard_finalize(void) {
    if (!g_sh.on) return 0;
    CM_CUDA(cudaDeviceSynchronize());
    for (int r = 0; r < g_sh.world; r++)
        if (r != g_sh.rank && g_sh.peer_base[r]) cudaIpcCloseMemHandle(g_sh.peer_base[r]);
    ncclCommDestroy(g_sh.comm);
    cudaFree(g_sh.arena);
    g_sh = ShardState();
    return 0;
}

int cm31_shard_info(int* rank, int* world, uint64_t stats[4]) {
    if (rank) *rank = g_sh.rank;
    if (world) *world = g_sh.on ? g_sh.world : 0;
    if (stats) {
        stats[0] = g_sh.arena_peak;
        stats[1] = g_sh.bytes_gathered;
        stats[2] = g_sh.bytes_reduced;
        stats[3] = g_sh.n_collectives;
    }
    return 0;
}

// Stream-ordered barrier over the ranks: later work on this rank's stream starts only after every rank's earlier work on
// ITS stream has completed (a 4-byte all-reduce).
int cm31_shard_barrier(void) {
    CM_REQUIRE(g_sh.on, "shard: not initialised");
    static u32* token = nullptr;
    if (!token) {
        CM_CUDA(cudaMalloc((void**)&token, 4));
        CM_CUDA(cudaMemsetAsync(token, 0, 4, stream()));
    }
    CM_NCCL(ncclAllReduce(token, token, 1, ncclUint32, ncclSum, g_sh.comm, stream()));
    g_sh.n_collectives++;
    return 0;
}

// New proof: every rank has finished reading its peers (barrier + host synchronisation), the arena restarts at offset 0.
int cm31_shard_arena_reset(void) {
    CM_REQUIRE(g_sh.on, "shard: not initialised");
    if (int e = cm31_lanes_join()) return e;
    if (int e = cm31_shard_barrier()) return e;
    CM_CUDA(cudaStreamSynchronize(stream()));
    g_sh.arena_pos = 0;
    g_sh.arena_active = true;
    return 0;
}
int cm31_shard_arena_suspend(void) {
    g_sh.arena_active = false;
    return 0;
}

// buf holds world * bytes_per_rank bytes; rank r contributed the r-th slice (in place).
int cm31_shard_allgather(void* buf, size_t bytes_per_rank) {
    CM_REQUIRE(g_sh.on && buf != nullptr, "shard_allgather: not initialised / null buffer");
    if (g_sh.world == 1 || bytes_per_rank == 0) return 0;
    CM_NCCL(ncclAllGather((uint8_t*)buf + bytes_per_rank * (size_t)g_sh.rank, buf, bytes_per_rank, ncclUint8, g_sh.comm, stream()));
    g_sh.bytes_gathered += bytes_per_rank * (size_t)(g_sh.world - 1);
    g_sh.n_collectives++;
    return 0;
}

// several in-place all-gathers as ONE NCCL group (the sharded layers of a Merkle tree, the four coordinate columns)
int cm31_shard_allgather_many(void* const* bufs, const size_t* bytes_per_rank, size_t n) {
    CM_REQUIRE(g_sh.on, "shard_allgather_many: not initialised");
    if (g_sh.world == 1 || n == 0) return 0;
    CM_NCCL(ncclGroupStart());
    for (size_t i = 0; i < n; i++) {
        if (bytes_per_rank[i] == 0) continue;
        ncclResult_t r = ncclAllGather((uint8_t*)bufs[i] + bytes_per_rank[i] * (size_t)g_sh.rank, bufs[i], bytes_per_rank[i], ncclUint8, g_sh.comm, stream());
        if (r != ncclSuccess) {
            ncclGroupEnd();
            set_error(std::string("ncclAllGather: ") + ncclGetErrorString(r));
            return -3;
        }
        g_sh.bytes_gathered += bytes_per_rank[i] * (size_t)(g_sh.world - 1);
    }
    CM_NCCL(ncclGroupEnd());
    g_sh.n_collectives++;
    return 0;
}

// element-wise u32 sum over ranks of a DEVICE buffer (multiplicity bins: counts stay far below 2^32)
int cm31_shard_allreduce_u32(uint32_t* buf, size_t n) {
    CM_REQUIRE(g_sh.on && buf != nullptr, "shard_allreduce: not initialised / null buffer");
    if (g_sh.world == 1 || n == 0) return 0;
    CM_NCCL(ncclAllReduce(buf, buf, n, ncclUint32, ncclSum, g_sh.comm, stream()));
    g_sh.bytes_reduced += 4 * n;
    g_sh.n_collectives++;
    return 0;
}

// the same for a small HOST table in which every entry is non-zero on exactly one rank (claimed sums, sampled values):
// upload, all-reduce, download, synchronise
int cm31_shard_allreduce_host_u32(uint32_t* host, size_t n) {
    CM_REQUIRE(g_sh.on && host != nullptr, "shard_allreduce_host: not initialised / null buffer");
    if (g_sh.world == 1 || n == 0) return 0;
    u32* d = nullptr;
    CM_CUDA(cudaMallocAsync(&d, 4 * n, stream()));
    CM_CUDA(cudaMemcpyAsync(d, host, 4 * n, cudaMemcpyHostToDevice, stream()));
    CM_NCCL(ncclAllReduce(d, d, n, ncclUint32, ncclSum, g_sh.comm, stream()));
    CM_CUDA(cudaMemcpyAsync(host, d, 4 * n, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaFreeAsync(d, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    g_sh.n_collectives++;
    return 0;
}

// the peer's copy of an arena buffer (same offset in rank `owner`'s arena); identity when owner < 0 or owner == my rank
int cm31_shard_peer(const void* p, int owner, const void** out) {
    CM_REQUIRE(out != nullptr, "shard_peer: null output");
    if (!g_sh.on || owner < 0 || owner == g_sh.rank || p == nullptr) {
        *out = p;
        return 0;
    }
    CM_REQUIRE(owner < g_sh.world, "shard_peer: owner out of range");
    CM_REQUIRE(shard_arena_owns(p), "shard_peer: the buffer does not live in the shard arena");
    *out = g_sh.peer_base[owner] + ((const uint8_t*)p - g_sh.arena);
    return 0;
}

// dst4 = sum over ranks of src4 (mod P), n elements per column: every rank sums ITS row range from all peers' src4 (peer
// loads), then the ranges are all-gathered in place.  src4 must be complete on every rank: a barrier is issued first.
// dst4 must not alias src4 (peers are still reading src4 while dst4 is written).
int cm31_shard_reduce_m31(uint32_t* const dst4[4], const uint32_t* const src4[4], size_t n) {
    CM_REQUIRE(g_sh.on, "shard_reduce_m31: not initialised");
    CM_REQUIRE(n % (size_t)g_sh.world == 0, "shard_reduce_m31: length not divisible by the world size");
    if (int e = cm31_shard_barrier()) return e;
    PeerCols4 pc;
    for (int r = 0; r < g_sh.world; r++)
        for (int k = 0; k < 4; k++) {
            const void* q = nullptr;
            if (int e = cm31_shard_peer(src4[k], r, &q)) return e;
            pc.p[r][k] = (const u32*)q;
        }
    const size_t count = n / (size_t)g_sh.world, first = count * (size_t)g_sh.rank;
    {
        ProfScope prof("shard_reduce_m31", 16ull * count * (size_t)(g_sh.world + 1));
        reduce_m31_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream()>>>(pc, g_sh.world, dst4[0], dst4[1], dst4[2], dst4[3], first, count);
        CM_LAUNCH_CHECK();
    }
    void* bufs[4] = {dst4[0], dst4[1], dst4[2], dst4[3]};
    size_t per[4] = {4 * count, 4 * count, 4 * count, 4 * count};
    return cm31_shard_allgather_many(bufs, per, 4);
}

}  // extern "C"

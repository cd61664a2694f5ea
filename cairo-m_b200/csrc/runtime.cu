// libcm31 runtime: errors, stream, stream-ordered device memory, gathers.
// Replaces Column<T> storage management of the reference backends
// (external/stwo/crates/prover/src/core/backend/mod.rs:46-65).
#include <algorithm>
#include <cstring>
#include <mutex>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace cm31 {

static thread_local std::string g_err;
static cudaStream_t g_stream = 0;  // the stream every cm31_* call issues on (= the current lane)
// Lanes: the proof's own work lives on the caller's stream (lane 0); groups that are independent of
// each other until the next join -- the per-component witness / logup / constraint programs and the
// per-size FFT and quotient batches of SMALL components -- may be issued on a side stream (lane 1),
// so their latency-bound launches overlap the large components' kernels.
static cudaStream_t g_main = 0, g_side = nullptr;
static cudaEvent_t g_ev_fork = nullptr, g_ev_join = nullptr;
static bool g_forked = false;
static int g_lanes_on = -1;
static bool g_pool_ready[64] = {false};

void set_error(const std::string& msg) { g_err = msg; }
cudaStream_t stream() { return g_stream; }

static int ensure_pool() {
    int dev = 0;
    CM_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !g_pool_ready[dev]) {
        cudaMemPool_t pool;
        CM_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t thresh = UINT64_MAX;  // keep freed blocks cached: proving reuses the same sizes
        CM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
        g_pool_ready[dev] = true;
    }
    return 0;
}

// ------------------------------------------------------------------ per-kernel timing
struct ProfRec {
    const char* name;
    uint64_t bytes;
    cudaEvent_t a, b;
    uint64_t ops = 0;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_free;
static uint64_t g_launches = 0;  // every bracketed launch, counted even when timing is off

static cudaEvent_t prof_event() {
    if (!g_prof_free.empty()) {
        cudaEvent_t e = g_prof_free.back();
        g_prof_free.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void prof_begin(const char* name, uint64_t alg_bytes, unsigned n_kernels) {
    g_launches += n_kernels;
    if (!g_prof_on) return;
    ProfRec r{name, alg_bytes, prof_event(), prof_event()};
    cudaEventRecord(r.a, stream());
    g_prof.push_back(r);
}
void prof_ops(uint64_t m31_ops) {
    if (g_prof_on && !g_prof.empty()) g_prof.back().ops += m31_ops;
}
bool prof_enabled() { return g_prof_on; }
void prof_end() {
    if (!g_prof_on) return;
    cudaEventRecord(g_prof.back().b, stream());
}

// Small host tables (pointer arrays, bytecode, constants) go through a page-locked staging ring so
// the H2D copy is a true async DMA (a pageable source costs a driver-side staging copy + ~10 us per
// call, which dominated the gaps between short kernels).  A slot is reused only after the ring
// wraps, at which point the stream is synchronised.
static uint8_t* g_stage = nullptr;      // pinned host ring
static uint8_t* g_stage_dev = nullptr;  // device ring: a table lives at the SAME offset in both
static size_t g_stage_pos = 0;
constexpr size_t STAGE_BYTES = 16u << 20;

int DeviceTable::upload(const void* host, size_t bytes) {
    if (int e = reserve(bytes)) return e;
    return fill(host, bytes == 0 ? 0 : bytes);
}
int DeviceTable::reserve(size_t bytes) {
    release();
    if (bytes == 0) bytes = 4;
    if (bytes > STAGE_BYTES / 4) {  // large: own allocation, plain copy (the runtime stages pageable sources itself)
        if (int e = ensure_pool()) return e;
        CM_CUDA(cudaMallocAsync(&d, bytes, stream()));
        owned = true;
        return 0;
    }
    if (!g_stage) {
        CM_CUDA(cudaHostAlloc((void**)&g_stage, STAGE_BYTES, cudaHostAllocDefault));
        CM_CUDA(cudaMalloc((void**)&g_stage_dev, STAGE_BYTES));
    }
    size_t need = (bytes + 255) & ~(size_t)255;
    if (g_stage_pos + need > STAGE_BYTES) {
        // wrap: every table issued so far (on either lane) must be consumed before its slot is reused
        CM_CUDA(cudaStreamSynchronize(g_main));
        if (g_side) CM_CUDA(cudaStreamSynchronize(g_side));
        g_stage_pos = 0;
    }
    d = g_stage_dev + g_stage_pos;
    ring_at = g_stage_pos;
    owned = false;
    g_stage_pos += need;
    return 0;
}
int DeviceTable::fill(const void* host, size_t bytes) {
    if (!host || bytes == 0) return 0;
    if (owned) {
        CM_CUDA(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, stream()));
        return 0;
    }
    memcpy(g_stage + ring_at, host, bytes);
    CM_CUDA(cudaMemcpyAsync(d, g_stage + ring_at, bytes, cudaMemcpyHostToDevice, stream()));
    return 0;
}
void DeviceTable::release() {
    if (d && owned) cudaFreeAsync(d, stream());
    d = nullptr;
    owned = false;
}

__global__ void gather_u32_kernel(const u32* const* cols, size_t n_cols, const u32* idx, size_t n_idx, u32* out) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n_cols * n_idx) return;
    size_t c = t / n_idx, q = t % n_idx;
    out[t] = cols[c][idx[q]];
}
__global__ void gather_words_kernel(const u32* const* srcs, const u32* src_id, const u32* word, size_t n, u32* out) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t < n) out[t] = srcs[src_id[t]][word[t]];
}
__global__ void gather_runs_kernel(const u32* const* srcs, const u32* src_id, const u32* word, const u32* out_off, size_t n, u32* out) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const u32* s = srcs[src_id[t]] + word[t];
    u32 o = out_off[t], cnt = out_off[t + 1] - o;
    for (u32 j = 0; j < cnt; j++) out[o + j] = s[j];
}
__global__ void gather_runs2_kernel(const u32* const* srcs, const u32* src_id, const u32* word, const u32* out_off, const u32* cnt, size_t n,
                                    u32* out) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const u32* s = srcs[src_id[t]] + word[t];
    u32 o = out_off[t], c = cnt[t];
    for (u32 j = 0; j < c; j++) out[o + j] = s[j];
}
// one CTA row (blockIdx.y) per grid request; thread t -> (row k = t / n_cols, column c = t % n_cols)
__global__ void gather_grid_kernel(const u32* const* srcs, const u32* desc, const u32* cols, const u32* rows, u32* out) {
    const u32* d = desc + 5 * blockIdx.y;
    const u32 col_off = d[0], n_cols = d[1], row_off = d[2], n_rows = d[3], base = d[4];
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < (size_t)n_cols * n_rows; t += (size_t)gridDim.x * blockDim.x) {
        u32 k = (u32)(t / n_cols), c = (u32)(t % n_cols);
        out[base + t] = __ldg(srcs[cols[col_off + c]] + rows[row_off + k]);
    }
}
__global__ void gather_hash_kernel(const u32* layer, const u32* idx, size_t n_idx, u32* out) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n_idx * 8) return;
    out[t] = layer[(size_t)idx[t >> 3] * 8 + (t & 7)];
}

}  // namespace cm31

using namespace cm31;

extern "C" {

const char* cm31_last_error(void) { return g_err.c_str(); }

int cm31_device_count(int* out) {
    CM_CUDA(cudaGetDeviceCount(out));
    return 0;
}
// The runtime state of this library (streams, staging rings, lane events, the sharding arena) belongs to ONE device and ONE
// host thread per process -- the deployment model is one process per GPU (SURVEY.md §8b "Threading", §8e).  The device is
// fixed by the first call that touches it; asking for another one afterwards is refused instead of leaving kernel-argument
// tables in the first device's memory.
static int g_bound_device = -1;
int cm31_set_device(int ordinal) {
    CM_REQUIRE(g_bound_device < 0 || g_bound_device == ordinal || g_stage == nullptr,
               "set_device: this process already runs on another device (one process per GPU; the staging rings live on the first device)");
    CM_CUDA(cudaSetDevice(ordinal));
    g_bound_device = ordinal;
    return 0;
}
int cm31_set_stream(void* s) {
    g_main = (cudaStream_t)s;
    g_stream = g_main;
    return 0;
}
// lane 1: the side stream, ordered after everything issued on lane 0 up to the FIRST switch since the
// last join (fork); lane 0: the caller's stream.  Work issued on one lane between a fork and the next
// cm31_lanes_join() must not depend on work issued on the other lane in that window, and buffers
// must be freed on the lane that last used them (or after the join).  CM31_NO_LANES=1 keeps
// everything on lane 0.
int cm31_lane(int lane) {
    if (g_lanes_on < 0) g_lanes_on = getenv("CM31_NO_LANES") ? 0 : 1;
    if (!g_lanes_on || lane == 0) {
        g_stream = g_main;
        return 0;
    }
    if (!g_side) {
        // highest priority: the side lane's few-CTA kernels are dispatched as soon as an SM slot frees up instead of
        // queueing behind every CTA of a large lane-0 kernel
        int prio_low = 0, prio_high = 0;
        CM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
        CM_CUDA(cudaStreamCreateWithPriority(&g_side, cudaStreamNonBlocking, prio_high));
        CM_CUDA(cudaEventCreateWithFlags(&g_ev_fork, cudaEventDisableTiming));
        CM_CUDA(cudaEventCreateWithFlags(&g_ev_join, cudaEventDisableTiming));
    }
    if (!g_forked) {
        CM_CUDA(cudaEventRecord(g_ev_fork, g_main));
        CM_CUDA(cudaStreamWaitEvent(g_side, g_ev_fork, 0));
        g_forked = true;
    }
    g_stream = g_side;
    return 0;
}
int cm31_lanes_join(void) {
    g_stream = g_main;
    if (g_forked) {
        CM_CUDA(cudaEventRecord(g_ev_join, g_side));
        CM_CUDA(cudaStreamWaitEvent(g_main, g_ev_join, 0));
        g_forked = false;
    }
    return 0;
}
// NVTX ranges (no-ops unless a profiler is attached): the prover marks its phases with the span names of the reference's
// `tracing` instrumentation (stark.hpp Span<B>)
static const bool g_nvtx_on = getenv("CM31_NO_NVTX") == nullptr;
int cm31_range_push(const char* name) {
    if (g_nvtx_on) nvtxRangePushA(name ? name : "");
    return 0;
}
int cm31_range_pop(void) {
    if (g_nvtx_on) nvtxRangePop();
    return 0;
}
int cm31_sync(void) {
    CM_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}
int cm31_profile_enable(int on) {
    g_prof_on = on != 0;
    return 0;
}
int cm31_profile_reset(void) {
    for (ProfRec& r : g_prof) {
        g_prof_free.push_back(r.a);
        g_prof_free.push_back(r.b);
    }
    g_prof.clear();
    g_launches = 0;
    return 0;
}
int cm31_profile_launches(uint64_t* out) {
    *out = g_launches;
    return 0;
}
int cm31_profile_report(char* buf, size_t cap, size_t* len) {
    CM_CUDA(cudaStreamSynchronize(stream()));
    struct Agg {
        const char* name;
        double ms = 0;
        uint64_t launches = 0, bytes = 0, ops = 0;
    };
    std::vector<Agg> aggs;
    for (ProfRec& r : g_prof) {
        float ms = 0;
        CM_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        Agg* a = nullptr;
        for (Agg& x : aggs)
            if (strcmp(x.name, r.name) == 0) a = &x;
        if (!a) {
            aggs.push_back(Agg());
            a = &aggs.back();
            a->name = r.name;
        }
        a->ms += ms;
        a->launches++;
        a->bytes += r.bytes;
        a->ops += r.ops;
    }
    std::string s = "[";
    for (size_t i = 0; i < aggs.size(); i++) {
        char line[256];
        snprintf(line, sizeof line, "%s{\"kernel\": \"%s\", \"ms\": %.6f, \"launches\": %llu, \"alg_bytes\": %llu, \"m31_ops\": %llu}",
                 i ? ", " : "", aggs[i].name, aggs[i].ms, (unsigned long long)aggs[i].launches, (unsigned long long)aggs[i].bytes,
                 (unsigned long long)aggs[i].ops);
        s += line;
    }
    s += "]";
    if (len) *len = s.size();
    CM_REQUIRE(buf == nullptr || s.size() + 1 <= cap, "profile_report: buffer too small");
    if (buf) memcpy(buf, s.c_str(), s.size() + 1);
    return 0;
}

// CSV "kernel,start_ms,dur_ms" of every bracketed launch since the last reset (start relative to the first one):
// the gaps between rows are host/driver time (uploads, allocations, synchronisations, transcript).
int cm31_profile_trace(char* buf, size_t cap, size_t* len) {
    CM_CUDA(cudaStreamSynchronize(stream()));
    std::string s;
    for (ProfRec& r : g_prof) {
        float t0 = 0, d = 0;
        CM_CUDA(cudaEventElapsedTime(&t0, g_prof[0].a, r.a));
        CM_CUDA(cudaEventElapsedTime(&d, r.a, r.b));
        char line[160];
        snprintf(line, sizeof line, "%s,%.4f,%.4f\n", r.name, t0, d);
        s += line;
    }
    if (len) *len = s.size();
    CM_REQUIRE(buf == nullptr || s.size() + 1 <= cap, "profile_trace: buffer too small");
    if (buf) memcpy(buf, s.c_str(), s.size() + 1);
    return 0;
}

// Head-room for the stream-ordered pool.  Growing the pool (cuMemCreate + map) while kernels are running stalls the caller for
// 30-240 ms on some boxes (round 1: the first step with two prover inputs in flight; round 2: proofs whose buffers outlive the
// call -- cm31_prove_cairo_m_async keeps the trees of proof i until proof i+1 has gathered from them -- make the free list
// differ from proof to proof, so a request occasionally fits no cached block).  cm31_pool_reserve_headroom(factor) makes the pool
// hold factor x its high-water mark of bytes in use: one allocation + free while the GPU is idle, after which later requests
// are served from cached memory.  Best effort: failures (not enough free memory) are ignored.
int cm31_pool_reserve_headroom(double factor) {
    int dev = 0;
    CM_CUDA(cudaGetDevice(&dev));
    cudaMemPool_t pool;
    CM_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t used_high = 0, reserved = 0;
    CM_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &used_high));
    CM_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved));
    const uint64_t want = (uint64_t)((double)used_high * factor);
    if (want <= reserved) return 0;
    size_t free_b = 0, total_b = 0;
    CM_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t extra = want - reserved;
    if (extra + (8ull << 30) > free_b) return 0;  // keep 8 GB clear of the pool
    CM_CUDA(cudaStreamSynchronize(g_main));
    if (g_side) CM_CUDA(cudaStreamSynchronize(g_side));
    void* p = nullptr;
    if (cudaMallocAsync(&p, extra, g_main) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    CM_CUDA(cudaFreeAsync(p, g_main));
    CM_CUDA(cudaStreamSynchronize(g_main));
    return 0;
}

int cm31_malloc(void** out, size_t bytes) {
    CM_REQUIRE(out != nullptr, "malloc: null out");
    if (bytes == 0) bytes = 4;
    int st = 0;
    if (cm31::shard_arena_alloc(out, bytes, &st)) return st;
    if (int e = ensure_pool()) return e;
    CM_CUDA(cudaMallocAsync(out, bytes, stream()));
    return 0;
}
int cm31_free(void* p) {
    if (p && cm31::shard_arena_owns(p)) return 0;  // arena buffers live until the per-proof reset
    if (p) CM_CUDA(cudaFreeAsync(p, stream()));
    return 0;
}
int cm31_memset0(void* p, size_t bytes) {
    CM_CUDA(cudaMemsetAsync(p, 0, bytes, stream()));
    return 0;
}
int cm31_h2d(void* dst, const void* src, size_t bytes) {
    CM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream()));
    return 0;
}
// Background host->device copies (prover-input staging): issued on a second stream so the DMA of
// the execution bundles / access log overlaps the input-independent start of the proof
// (twiddles, preprocessed tree).  cm31_h2d_bg orders the copy after everything already enqueued on
// the main stream (the destination was just allocated there); cm31_bg_fence makes the main stream
// wait for all background copies issued so far.
static cudaStream_t g_copy_stream = nullptr;
static cudaEvent_t g_copy_event = nullptr, g_main_event = nullptr;
static int ensure_copy_stream() {
    if (!g_copy_stream) {
        CM_CUDA(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
        CM_CUDA(cudaEventCreateWithFlags(&g_copy_event, cudaEventDisableTiming));
        CM_CUDA(cudaEventCreateWithFlags(&g_main_event, cudaEventDisableTiming));
    }
    return 0;
}
int cm31_h2d_bg(void* dst, const void* src_host, size_t bytes) {
    if (int e = ensure_copy_stream()) return e;
    CM_CUDA(cudaEventRecord(g_main_event, stream()));
    CM_CUDA(cudaStreamWaitEvent(g_copy_stream, g_main_event, 0));
    CM_CUDA(cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, g_copy_stream));
    return 0;
}
// Bracketed form: cm31_bg_begin orders the copy stream after the main stream ONCE (all destinations allocated before it),
// cm31_h2d_bg_ordered then only enqueues the copy (2 driver calls fewer per copy).
// Deferred form (cm31_bg_defer / cm31_bg_release): a prefetch of the NEXT segment's input is usually issued right before the
// proof of the current one, and a 327 MB DMA that runs during the launch-bound start of that proof (preprocessed tree, trace
// fill: hundreds of short kernels and small tables) delays it -- measured: preprocessed + trace phases 4.7 -> 7.7 ms while the
// copy is in flight, the long-kernel phases unaffected.  While deferral is on, background copies and their marks are only
// RECORDED (tagged with a ticket); the prover releases them at a point of its own choosing (before a commitment: seconds of
// FFT / Merkle kernels, few launches), or at the latest when the input they belong to is about to be consumed.
struct DeferredBg {
    uint64_t ticket;
    void* dst;
    const void* src;
    size_t bytes;
    uint32_t mark;
    bool is_mark;
};
static std::vector<DeferredBg> g_bg_deferred;
static uint64_t g_bg_ticket = 0;  // 0 = not deferring
static uint64_t g_bg_next_ticket = 1;
int cm31_bg_defer(int on, uint64_t* ticket_out) {
    g_bg_ticket = on ? g_bg_next_ticket++ : 0;
    if (ticket_out) *ticket_out = g_bg_ticket;
    return 0;
}
int cm31_bg_begin(void) {
    if (g_bg_ticket) return 0;
    if (int e = ensure_copy_stream()) return e;
    CM_CUDA(cudaEventRecord(g_main_event, stream()));
    CM_CUDA(cudaStreamWaitEvent(g_copy_stream, g_main_event, 0));
    return 0;
}
int cm31_h2d_bg_ordered(void* dst, const void* src_host, size_t bytes) {
    if (int e = ensure_copy_stream()) return e;
    if (g_bg_ticket) {
        g_bg_deferred.push_back(DeferredBg{g_bg_ticket, dst, src_host, bytes, 0, false});
        return 0;
    }
    CM_CUDA(cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, g_copy_stream));
    return 0;
}
// Marks: an event after the background copies issued so far; cm31_bg_wait(mark) makes the CURRENT lane wait for it, so a
// component's trace fill starts as soon as ITS rows have landed instead of after the whole input (256 recyclable events:
// a recycled mark only over-waits).
static cudaEvent_t g_marks[256] = {nullptr};
static uint32_t g_next_mark = 0;
int cm31_bg_mark(uint32_t* mark_out) {
    CM_REQUIRE(mark_out != nullptr, "bg_mark: null output");
    if (int e = ensure_copy_stream()) return e;
    uint32_t m = g_next_mark++ % 256;
    if (!g_marks[m]) CM_CUDA(cudaEventCreateWithFlags(&g_marks[m], cudaEventDisableTiming));
    *mark_out = m;
    if (g_bg_ticket) {
        g_bg_deferred.push_back(DeferredBg{g_bg_ticket, nullptr, nullptr, 0, m, true});
        return 0;
    }
    CM_CUDA(cudaEventRecord(g_marks[m], g_copy_stream));
    return 0;
}
// Issues the recorded copies with a ticket <= upto (0 = all), oldest first, ordered after the CURRENT point of the main stream.
// A mark must be released before anything waits on it: cm31_prove_cairo_m releases the tickets of the input it is about to
// consume before its first cm31_bg_wait.
int cm31_bg_cancel(uint64_t ticket) {  // the destination buffers are about to be freed: forget the recorded copies
    g_bg_deferred.erase(std::remove_if(g_bg_deferred.begin(), g_bg_deferred.end(), [&](const DeferredBg& d) { return d.ticket == ticket; }),
                        g_bg_deferred.end());
    return 0;
}
// Throttled form of a released copy: a few CTAs read the page-locked host buffer through the unified address space, so the
// PCIe reads in flight are bounded (16 CTAs x 256 threads x 16 B = 64 KB, ~25-30 GB/s) and the link is never saturated.  Measured
// on a 2^22-step proof with the next segment's 327 MB travelling meanwhile: copy-engine DMA (55 GB/s for 6 ms) stretches the
// phases it overlaps by 3.0 ms whichever phases they are; the throttled kernel copy (~13 ms long) by about 1 ms in total.
__global__ void bg_copy_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
static int bg_copy(void* dst, const void* src, size_t bytes, int throttle_ctas) {
    if (throttle_ctas > 0 && bytes >= (1u << 20) && ((uintptr_t)dst & 15) == 0) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.devicePointer != nullptr && ((uintptr_t)at.devicePointer & 15) == 0 &&
            (at.type == cudaMemoryTypeHost)) {
            const size_t body = bytes & ~(size_t)15;  // 16-byte words by the kernel, the last <= 15 bytes by a plain copy
            bg_copy_kernel<<<throttle_ctas, 256, 0, g_copy_stream>>>((uint4*)dst, (const uint4*)at.devicePointer, body / 16);
            CM_LAUNCH_CHECK();
            if (body < bytes)
                CM_CUDA(cudaMemcpyAsync((uint8_t*)dst + body, (const uint8_t*)src + body, bytes - body, cudaMemcpyHostToDevice, g_copy_stream));
            return 0;
        }
        cudaGetLastError();
    }
    CM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_copy_stream));
    return 0;
}
int cm31_bg_release_throttled(uint64_t upto, int throttle_ctas);
int cm31_bg_release(uint64_t upto) { return cm31_bg_release_throttled(upto, 0); }
int cm31_bg_release_throttled(uint64_t upto, int throttle_ctas) {
    size_t n = 0;
    while (n < g_bg_deferred.size() && (upto == 0 || g_bg_deferred[n].ticket <= upto)) n++;
    if (n == 0) return 0;
    if (int e = ensure_copy_stream()) return e;
    CM_CUDA(cudaEventRecord(g_main_event, g_main));
    CM_CUDA(cudaStreamWaitEvent(g_copy_stream, g_main_event, 0));
    for (size_t i = 0; i < n; i++) {
        const DeferredBg& d = g_bg_deferred[i];
        if (d.is_mark) CM_CUDA(cudaEventRecord(g_marks[d.mark], g_copy_stream));
        else if (int e = bg_copy(d.dst, d.src, d.bytes, throttle_ctas)) return e;
    }
    g_bg_deferred.erase(g_bg_deferred.begin(), g_bg_deferred.begin() + n);
    return 0;
}
int cm31_bg_wait(uint32_t mark) {
    CM_REQUIRE(mark < 256, "bg_wait: bad mark");
    if (!g_marks[mark]) return 0;
    CM_CUDA(cudaStreamWaitEvent(stream(), g_marks[mark], 0));
    return 0;
}
int cm31_bg_fence(void) {
    if (!g_copy_stream) return 0;
    CM_CUDA(cudaEventRecord(g_copy_event, g_copy_stream));
    CM_CUDA(cudaStreamWaitEvent(stream(), g_copy_event, 0));
    return 0;
}

int cm31_d2h(void* dst, const void* src, size_t bytes) {
    CM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}
int cm31_d2d(void* dst, const void* src, size_t bytes) {
    CM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream()));
    return 0;
}

// ---- peer memory (NVLink): buffers other ranks' kernels read directly (sharded commitment)
int cm31_ipc_alloc(size_t bytes, void** out, uint8_t handle_out[64]) {
    CM_REQUIRE(out != nullptr && handle_out != nullptr, "ipc_alloc: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CM_CUDA(cudaMalloc(out, bytes == 0 ? 4 : bytes));  // IPC handles need a cudaMalloc allocation, not the async pool
    cudaIpcMemHandle_t h;
    CM_CUDA(cudaIpcGetMemHandle(&h, *out));
    memcpy(handle_out, &h, 64);
    return 0;
}
int cm31_ipc_free(void* p) {
    if (p) CM_CUDA(cudaFree(p));
    return 0;
}
int cm31_ipc_open(const uint8_t handle[64], void** out) {
    CM_REQUIRE(out != nullptr && handle != nullptr, "ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CM_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int cm31_ipc_close(void* p) {
    if (p) CM_CUDA(cudaIpcCloseMemHandle(p));
    return 0;
}

int cm31_gather_u32(const uint32_t* const* cols, size_t n_cols, const uint32_t* idx_host, size_t n_idx,
                    uint32_t* out_host) {
    if (n_cols == 0 || n_idx == 0) return 0;
    DeviceTable dcols, didx;
    if (int e = dcols.upload(cols, n_cols * sizeof(void*))) return e;
    if (int e = didx.upload(idx_host, n_idx * 4)) return e;
    u32* dout = nullptr;
    size_t total = n_cols * n_idx;
    CM_CUDA(cudaMallocAsync(&dout, total * 4, stream()));
    ProfScope prof("gather_u32", 8ull * total);
    gather_u32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream()>>>((const u32* const*)dcols.d, n_cols,
                                                                            (const u32*)didx.d, n_idx, dout);
    CM_LAUNCH_CHECK();
    CM_CUDA(cudaMemcpyAsync(out_host, dout, total * 4, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaFreeAsync(dout, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

int cm31_gather_words(const uint32_t* const* srcs, size_t n_srcs, const uint32_t* src_id_host, const uint32_t* word_idx_host,
                      size_t n, uint32_t* out_host) {
    if (n == 0) return 0;
    for (size_t k = 0; k < n; k++) CM_REQUIRE(src_id_host[k] < n_srcs, "gather_words: source id out of range");
    DeviceTable dsrcs, dsid, dword;
    if (int e = dsrcs.upload(srcs, n_srcs * sizeof(void*))) return e;
    if (int e = dsid.upload(src_id_host, n * 4)) return e;
    if (int e = dword.upload(word_idx_host, n * 4)) return e;
    u32* dout = nullptr;
    CM_CUDA(cudaMallocAsync(&dout, n * 4, stream()));
    {
        ProfScope prof("gather_words", 8ull * n);
        gather_words_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>((const u32* const*)dsrcs.d, (const u32*)dsid.d,
                                                                              (const u32*)dword.d, n, dout);
    }
    CM_LAUNCH_CHECK();
    CM_CUDA(cudaMemcpyAsync(out_host, dout, n * 4, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaFreeAsync(dout, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

int cm31_gather_runs(const uint32_t* const* srcs, size_t n_srcs, const uint32_t* src_id_host, const uint32_t* word_idx_host,
                     const uint32_t* out_off_host, size_t n, uint32_t* out_host) {
    if (n == 0) return 0;
    for (size_t k = 0; k < n; k++) CM_REQUIRE(src_id_host[k] < n_srcs && out_off_host[k] <= out_off_host[k + 1], "gather_runs: bad request");
    size_t total = out_off_host[n];
    DeviceTable dsrcs, dsid, dword, doff;
    if (int e = dsrcs.upload(srcs, n_srcs * sizeof(void*))) return e;
    if (int e = dsid.upload(src_id_host, n * 4)) return e;
    if (int e = dword.upload(word_idx_host, n * 4)) return e;
    if (int e = doff.upload(out_off_host, (n + 1) * 4)) return e;
    u32* dout = nullptr;
    CM_CUDA(cudaMallocAsync(&dout, std::max<size_t>(4, total * 4), stream()));
    {
        ProfScope prof("gather_runs", 8ull * total);
        gather_runs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>((const u32* const*)dsrcs.d, (const u32*)dsid.d,
                                                                             (const u32*)dword.d, (const u32*)doff.d, n, dout);
    }
    CM_LAUNCH_CHECK();
    CM_CUDA(cudaMemcpyAsync(out_host, dout, total * 4, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaFreeAsync(dout, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

// Asynchronous form: the gather kernels and the device->host copy of the result are enqueued, *result_out names the page-locked
// buffer the words land in, and cm31_gather_wait() blocks until they have.  The prover uses it to leave the assembly and the
// serialisation of a proof (host work, ~1.4 ms) for a moment when the NEXT proof has the GPU busy (cm31_prove_cairo_m_async).
static cudaEvent_t g_gather_event = nullptr;
int cm31_gather_batch_async(const uint32_t* const* srcs, size_t n_srcs, const uint32_t* src_id_host, const uint32_t* word_idx_host,
                            const uint32_t* out_off_host, const uint32_t* cnt_host, size_t n_runs, const uint32_t* grid_desc_host,
                            size_t n_grids, const uint32_t* grid_cols_host, size_t n_grid_cols, const uint32_t* grid_rows_host,
                            size_t n_grid_rows, size_t total_words, const uint32_t** result_out) {
    CM_REQUIRE(result_out != nullptr, "gather_batch_async: null result pointer");
    *result_out = nullptr;
    if (total_words == 0) return 0;
    for (size_t k = 0; k < n_runs; k++)
        CM_REQUIRE(src_id_host[k] < n_srcs && (size_t)out_off_host[k] + cnt_host[k] <= total_words, "gather_batch: bad run request");
    size_t max_work = 0;
    for (size_t g = 0; g < n_grids; g++) {
        const uint32_t* d = grid_desc_host + 5 * g;
        CM_REQUIRE((size_t)d[0] + d[1] <= n_grid_cols && (size_t)d[2] + d[3] <= n_grid_rows && (size_t)d[4] + (size_t)d[1] * d[3] <= total_words,
                   "gather_batch: bad grid request");
        max_work = std::max(max_work, (size_t)d[1] * d[3]);
    }
    for (size_t c = 0; c < n_grid_cols; c++) CM_REQUIRE(grid_cols_host[c] < n_srcs, "gather_batch: bad grid column");
    // page-locked landing buffers (grow-only, two of them alternate): the result copy is a true DMA, and the caller may read
    // the result of one gather while the next proof is already running (cm31_gather_batch_async / cm31_gather_wait)
    static u32* pinned[2] = {nullptr, nullptr};
    static size_t pinned_words[2] = {0, 0};
    static int which = 0;
    which ^= 1;
    if (pinned_words[which] < total_words) {
        if (pinned[which]) cudaFreeHost(pinned[which]);
        pinned_words[which] = std::max<size_t>(total_words * 2, (size_t)1 << 18);
        CM_CUDA(cudaHostAlloc((void**)&pinned[which], pinned_words[which] * 4, cudaHostAllocDefault));
    }
    DeviceTable dsrcs, dsid, dword, doff, dcnt, ddesc, dcols, drows;
    if (int e = dsrcs.upload(srcs, n_srcs * sizeof(void*))) return e;
    u32* dout = nullptr;
    CM_CUDA(cudaMallocAsync(&dout, total_words * 4, stream()));
    ProfScope prof("gather_runs", 8ull * total_words, (n_runs ? 1 : 0) + (n_grids ? 1 : 0));
    if (n_runs) {
        if (int e = dsid.upload(src_id_host, n_runs * 4)) return e;
        if (int e = dword.upload(word_idx_host, n_runs * 4)) return e;
        if (int e = doff.upload(out_off_host, n_runs * 4)) return e;
        if (int e = dcnt.upload(cnt_host, n_runs * 4)) return e;
        gather_runs2_kernel<<<(unsigned)((n_runs + 255) / 256), 256, 0, stream()>>>((const u32* const*)dsrcs.d, (const u32*)dsid.d, (const u32*)dword.d,
                                                                                   (const u32*)doff.d, (const u32*)dcnt.d, n_runs, dout);
        CM_LAUNCH_CHECK();
    }
    if (n_grids) {
        if (int e = ddesc.upload(grid_desc_host, n_grids * 20)) return e;
        if (int e = dcols.upload(grid_cols_host, n_grid_cols * 4)) return e;
        if (int e = drows.upload(grid_rows_host, n_grid_rows * 4)) return e;
        dim3 grid((unsigned)std::min<size_t>((max_work + 255) / 256, 64), (unsigned)n_grids);
        gather_grid_kernel<<<grid, 256, 0, stream()>>>((const u32* const*)dsrcs.d, (const u32*)ddesc.d, (const u32*)dcols.d, (const u32*)drows.d, dout);
        CM_LAUNCH_CHECK();
    }
    CM_CUDA(cudaMemcpyAsync(pinned[which], dout, total_words * 4, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaFreeAsync(dout, stream()));
    if (!g_gather_event) CM_CUDA(cudaEventCreateWithFlags(&g_gather_event, cudaEventDisableTiming));
    CM_CUDA(cudaEventRecord(g_gather_event, stream()));
    *result_out = pinned[which];
    return 0;
}
// the words requested by the last cm31_gather_batch_async have landed in its result buffer
int cm31_gather_wait(void) {
    if (g_gather_event) CM_CUDA(cudaEventSynchronize(g_gather_event));
    return 0;
}
int cm31_gather_batch(const uint32_t* const* srcs, size_t n_srcs, const uint32_t* src_id_host, const uint32_t* word_idx_host,
                      const uint32_t* out_off_host, const uint32_t* cnt_host, size_t n_runs, const uint32_t* grid_desc_host,
                      size_t n_grids, const uint32_t* grid_cols_host, size_t n_grid_cols, const uint32_t* grid_rows_host,
                      size_t n_grid_rows, size_t total_words, uint32_t* out_host) {
    if (total_words == 0) return 0;
    const uint32_t* res = nullptr;
    if (int e = cm31_gather_batch_async(srcs, n_srcs, src_id_host, word_idx_host, out_off_host, cnt_host, n_runs, grid_desc_host, n_grids,
                                        grid_cols_host, n_grid_cols, grid_rows_host, n_grid_rows, total_words, &res))
        return e;
    if (int e = cm31_gather_wait()) return e;
    memcpy(out_host, res, total_words * 4);
    return 0;
}

int cm31_gather_hash(const uint32_t* layer, const uint32_t* idx_host, size_t n_idx, uint32_t* out_host) {
    if (n_idx == 0) return 0;
    DeviceTable didx;
    if (int e = didx.upload(idx_host, n_idx * 4)) return e;
    u32* dout = nullptr;
    CM_CUDA(cudaMallocAsync(&dout, n_idx * 32, stream()));
    ProfScope prof("gather_hash", 64ull * n_idx);
    gather_hash_kernel<<<(unsigned)((n_idx * 8 + 255) / 256), 256, 0, stream()>>>(layer, (const u32*)didx.d, n_idx,
                                                                                 dout);
    CM_LAUNCH_CHECK();
    CM_CUDA(cudaMemcpyAsync(out_host, dout, n_idx * 32, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaFreeAsync(dout, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

}  // extern "C"

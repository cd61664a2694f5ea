// Device adapter on sm_100a: the runner's execution trace + memory-access log -> the prover input
// (per-opcode ExecutionBundles, the global data-access log, clock-update rows, the distinct touched cells)
// resident in HBM, so the serial host step BEFORE the proving hot path is no longer
// the bottleneck of a 27 ms proof.
//
// Replaces (the host restatement of the same functions is csrc/cairo/vm.hpp::import_from_vm):
//   import_internal / import_from_runner_output   crates/prover/src/adapter/mod.rs:97-193, 233-…
//   ExecutionBundleIterator::next                 crates/prover/src/adapter/memory.rs:264-403
//   Memory::push (prev clock / prev value, clock-update splitting at RC20_LIMIT, initial/final cells)
//                                                 crates/prover/src/adapter/memory.rs:470-…
//   IoTraceEntry {fp, pc}, IoMemoryEntry {address, value[4]}   crates/prover/src/adapter/io.rs:38-60
//
// The reference walks the two logs serially through a HashMap.  Here:
//   1. one thread per STEP reads the opcode of the instruction at pc (preloaded program cell) -> number of log
//      entries the step consumes (1-2 instruction words + its data accesses); an exclusive scan gives every
//      step its offset into the memory log (the fetched word is then checked against the log itself);
//   2. the log is sorted by address, stable in access order (LSD radix sort over the address bits in use) with
//      {access index, clock, value} as payload: the neighbouring element of the sorted order IS Memory::push's
//      previous (clock, value) of that cell;
//   3. one thread per sorted position resolves prev_clock / prev_value and the clock-update count
//      (delta / RC20_LIMIT); a scan in ACCESS order places the clock-update rows where the reference pushes them;
//   4. steps are stably partitioned by opcode (radix sort on the 6-bit opcode) = states_by_opcodes in execution
//      order; bundles are written straight into the per-component row buffers the trace fill consumes.
// Sorting and prefix sums use CUB (toolkit library, plumbing); the resolution kernels are hand-written.
// Only O(distinct cells) data returns to the host (boundary memory rows + the partial Poseidon2 Merkle trees
// are built there by the same code as the host adapter).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "cairo/vm.hpp"
#include "common.cuh"

namespace cm31 {

enum : u32 { ADAPTER_ERR_PC = 1, ADAPTER_ERR_OPCODE = 2, ADAPTER_ERR_FETCH = 4, ADAPTER_ERR_END = 8 };
constexpr u32 ADAPTER_MAX_OPCODE = 64;

struct OpcodeTables {
    uint8_t n_acc[ADAPTER_MAX_OPCODE];   // data accesses of the opcode, 0xFF = not an opcode
    uint8_t n_inst[ADAPTER_MAX_OPCODE];  // QM31 words fetched for the instruction (1 or 2)
    uint8_t size[ADAPTER_MAX_OPCODE];    // instruction size in M31 words
};
__constant__ OpcodeTables c_ops;

// ---- 1. per step: opcode, log entries consumed, data accesses; per-opcode step counts
__global__ void __launch_bounds__(256) adapter_count_kernel(const uint2* __restrict__ trace, u32 n_steps, const uint4* __restrict__ init,
                                                            u32 n_init, u32* __restrict__ opcode, u32* __restrict__ cnt,
                                                            u32* __restrict__ ndata, u32* __restrict__ hist, u32* __restrict__ err) {
    __shared__ u32 sh[ADAPTER_MAX_OPCODE];
    if (threadIdx.x < ADAPTER_MAX_OPCODE) sh[threadIdx.x] = 0;
    __syncthreads();
    u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_steps) {
        u32 pc = trace[s].y;  // IoTraceEntry {fp, pc}
        u32 op = ADAPTER_MAX_OPCODE;
        if (pc < n_init) op = __ldg(&init[pc]).x;
        else atomicOr(err, ADAPTER_ERR_PC);
        if (op >= ADAPTER_MAX_OPCODE || c_ops.n_acc[op] == 0xFF) {
            if (pc < n_init) atomicOr(err, ADAPTER_ERR_OPCODE);
            opcode[s] = 0;
            cnt[s] = 0;
            ndata[s] = 0;
        } else {
            opcode[s] = op;
            cnt[s] = (u32)c_ops.n_inst[op] + c_ops.n_acc[op];
            ndata[s] = c_ops.n_acc[op];
            unsigned peers = __match_any_sync(__activemask(), op);
            if ((threadIdx.x & 31) == (u32)(__ffs(peers) - 1)) atomicAdd(&sh[op], (u32)__popc(peers));
        }
    }
    __syncthreads();
    if (threadIdx.x < ADAPTER_MAX_OPCODE && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// ---- per step: check the fetched instruction against the log, emit sort keys (address) and the clock of every entry
__global__ void __launch_bounds__(256) adapter_keys_kernel(const uint2* __restrict__ trace, u32 n_steps, const u32* __restrict__ mem, u32 n_mem,
                                                           const u32* __restrict__ opcode, const u32* __restrict__ off,
                                                           const u32* __restrict__ cnt, u32* __restrict__ keys, uint4* __restrict__ pay,
                                                           u32* __restrict__ max_addr, u32* __restrict__ err) {
    u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    u32 mx = 0;
    if (s < n_steps) {
        u32 o = off[s], c = cnt[s];
        if ((size_t)o + c > n_mem) {
            atomicOr(err, ADAPTER_ERR_END);  // "Unexpected end of trace while reading multi-word instruction or operand"
        } else if (c) {
            u32 pc = trace[s].y, op = opcode[s];
            const u32* e = mem + 5 * (size_t)o;
            bool ok = e[0] == pc && e[1] == op;  // UnexpectedMemoryAccess {expected: pc, found}
            if (c_ops.n_inst[op] == 2) ok = ok && e[5] == pc + 1;
            if (!ok) atomicOr(err, ADAPTER_ERR_FETCH);
            for (u32 k = 0; k < c; k++) {
                u32 a = e[5 * k];
                keys[o + k] = a;
                pay[o + k] = make_uint4(o + k, s + 1, e[5 * k + 1], 0u);  // access index, clock (0 is reserved for preloaded values), value limb 0
                mx = max(mx, a);
            }
        }
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(max_addr, mx);
}

__global__ void adapter_iota_kernel(u32* p, u32 n) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// ---- 3. one thread per position of the address-sorted log: Memory::push's (prev_clock, prev_value) and the number of
// clock-update rows; head[p] = 1 on the first access of a cell.  The sort carries {access index, clock, value limb 0} as its
// payload, so the predecessor is the neighbouring (coalesced) element: nothing is gathered through the permutation, and the
// only scattered traffic is ONE 8-byte store per access (the first version gathered 4-byte words through the permutation
// and scattered three: 4.0 GB of DRAM sectors for 0.6 GB of algorithmic bytes, profiles/full_adapter_resolve_kernel_r01e.md).
__device__ __forceinline__ void adapter_prev(u32 p, const u32* addr_sorted, const uint4* pay, const uint4* init, u32 n_init, bool& head,
                                             u32& pclk, u32& pv0) {
    u32 a = addr_sorted[p];
    head = p == 0 || addr_sorted[p - 1] != a;
    if (!head) {
        uint4 q = pay[p - 1];
        pclk = q.y;
        pv0 = q.z;
    } else if (a < n_init) {  // preloaded cell: (value, clock 0)
        pclk = 0;
        pv0 = __ldg(&init[a]).x;
    } else {  // first touch of a fresh cell: its own value at clock 0
        pclk = 0;
        pv0 = pay[p].z;
    }
}
__global__ void __launch_bounds__(256) adapter_resolve_kernel(const u32* __restrict__ addr_sorted, const uint4* __restrict__ pay, u32 n_mem,
                                                              const uint4* __restrict__ init, u32 n_init, uint2* __restrict__ prev,
                                                              u32* __restrict__ n_cu, u32* __restrict__ head_flag) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_mem) return;
    bool head;
    u32 pclk, pv0;
    adapter_prev(p, addr_sorted, pay, init, n_init, head, pclk, pv0);
    uint4 me = pay[p];
    u32 i = me.x, clock = me.y, steps = 0;
    if (clock > pclk && clock - pclk > RC20_LIMIT) steps = (clock - pclk) / RC20_LIMIT;
    prev[i] = make_uint2(pclk + steps * RC20_LIMIT, pv0);  // < clock < P: no reduction needed
    if (steps) n_cu[i] = steps;                            // n_cu is zero-filled: rows are rare
    head_flag[p] = head ? 1u : 0u;
}

// ---- clock-update rows {address, prev_clk, initial value[4]} at the position the reference pushes them (access order)
__global__ void __launch_bounds__(256) adapter_clock_update_kernel(const u32* __restrict__ addr_sorted, const uint4* __restrict__ pay, u32 n_mem,
                                                                   const u32* __restrict__ mem, const uint4* __restrict__ init, u32 n_init,
                                                                   const u32* __restrict__ n_cu, const u32* __restrict__ cu_off,
                                                                   u32* __restrict__ rows) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_mem) return;
    u32 i = pay[p].x, steps = n_cu[i];
    if (!steps) return;
    bool head;
    u32 pclk, pv0;
    adapter_prev(p, addr_sorted, pay, init, n_init, head, pclk, pv0);
    u32 a = addr_sorted[p];
    uint4 v;
    if (a < n_init) v = __ldg(&init[a]);
    else {  // first access of the cell = lower bound of `a` in the sorted keys
        u32 lo = 0, hi = p;
        while (lo < hi) {
            u32 mid = (lo + hi) >> 1;
            if (addr_sorted[mid] < a) lo = mid + 1;
            else hi = mid;
        }
        const u32* e = mem + 5 * (size_t)pay[lo].x;
        v = make_uint4(e[1], e[2], e[3], e[4]);
    }
    u32* r = rows + 6 * (size_t)cu_off[i];
    for (u32 t = 0; t < steps; t++, r += 6) {
        r[0] = a;
        r[1] = pclk + t * RC20_LIMIT;
        r[2] = v.x; r[3] = v.y; r[4] = v.z; r[5] = v.w;
    }
}

// ---- distinct cells: {address, first value[4], last value[4], last clock}
__global__ void __launch_bounds__(256) adapter_cells_kernel(const u32* __restrict__ addr_sorted, const uint4* __restrict__ pay, u32 n_mem,
                                                            const u32* __restrict__ mem, const u32* __restrict__ head_incl,
                                                            u32* __restrict__ cells) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_mem) return;
    u32 a = addr_sorted[p];
    bool head = p == 0 || addr_sorted[p - 1] != a, tail = p + 1 == n_mem || addr_sorted[p + 1] != a;
    if (!head && !tail) return;
    u32* c = cells + 10 * (size_t)(head_incl[p] - 1);
    uint4 me = pay[p];
    const u32* e = mem + 5 * (size_t)me.x;
    if (head) {
        c[0] = a;
        c[1] = e[1]; c[2] = e[2]; c[3] = e[3]; c[4] = e[4];
    }
    if (tail) {
        c[5] = e[1]; c[6] = e[2]; c[7] = e[3]; c[8] = e[4];
        c[9] = me.y;
    }
}

// ---- 4. bundles, in opcode-sorted step order, into the per-component row buffers (12 words each)
struct AdapterDest {
    u32* base[ADAPTER_MAX_OPCODE];   // where the opcode's group starts (device words), null = no steps
    u32 start[ADAPTER_MAX_OPCODE];   // first position of the opcode's group in the sorted step order
};
__global__ void __launch_bounds__(256) adapter_bundles_kernel(const u32* __restrict__ steps_sorted, u32 n_steps, const uint2* __restrict__ trace,
                                                              const u32* __restrict__ opcode, const u32* __restrict__ off,
                                                              const u32* __restrict__ dstart, const u32* __restrict__ mem,
                                                              const uint2* __restrict__ prev, const AdapterDest* __restrict__ dest) {
    u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_steps) return;
    u32 s = steps_sorted[q], op = opcode[s], o = off[s];
    uint2 r = trace[s];
    const u32* e = mem + 5 * (size_t)o;
    u32 size = c_ops.size[op];
    u32 inst[6];
#pragma unroll
    for (u32 k = 0; k < 4; k++) inst[k] = k < size ? e[1 + k] : 0u;
    inst[4] = size > 4 ? e[6] : 0u;  // second QM31 word of the instruction, fetched at the same clock (memory.rs:317-339)
    inst[5] = size > 5 ? e[7] : 0u;
    uint4* out = (uint4*)(dest->base[op] + 12 * (size_t)(q - dest->start[op]));
    out[0] = make_uint4(r.y, r.x, s + 1, prev[o].x);  // pc, fp, clock, inst_prev_clock
    out[1] = make_uint4(inst[0], inst[1], inst[2], inst[3]);
    out[2] = make_uint4(inst[4], inst[5], dstart[s], (u32)c_ops.n_acc[op]);
}
// the global data-access log {address, prev_clock, prev_value, value}, in execution order (instruction fetches excluded)
__global__ void __launch_bounds__(256) adapter_accesses_kernel(u32 n_steps, const u32* __restrict__ opcode, const u32* __restrict__ off,
                                                               const u32* __restrict__ dstart, const u32* __restrict__ mem,
                                                               const uint2* __restrict__ prev, uint4* __restrict__ accesses) {
    u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_steps) return;
    u32 op = opcode[s], n = c_ops.n_acc[op], i = off[s] + c_ops.n_inst[op], d = dstart[s];
    for (u32 k = 0; k < n; k++, i++) {
        const u32* e = mem + 5 * (size_t)i;
        uint2 pr = prev[i];
        accesses[d + k] = make_uint4(e[0], pr.x, pr.y, e[1]);
    }
}

struct AdapterPlan {
    u32 n_steps = 0, n_mem = 0, n_init = 0;
    std::vector<void*> owned;
    uint2* trace = nullptr;
    u32 *mem = nullptr, *opcode = nullptr, *cnt = nullptr, *off = nullptr, *ndata = nullptr, *dstart = nullptr;
    uint4* init = nullptr;
    u32 *keys = nullptr, *addr_sorted = nullptr, *iota = nullptr, *steps_sorted = nullptr, *op_sorted = nullptr;
    uint4 *pay = nullptr, *pay_sorted = nullptr;  // {access index, clock, value limb 0, -}: the sort's payload
    uint2* prev = nullptr;                        // per access: {prev_clock, prev_value limb 0}
    u32 *n_cu = nullptr, *cu_off = nullptr, *head = nullptr, *head_incl = nullptr;
    u32* small = nullptr;  // [0..63] hist, [64] err, [65] max address
    void* temp = nullptr;
    size_t temp_bytes = 0;
    u32 n_cells = 0, n_cu_rows = 0, n_data = 0;
    u32 upload_mark = 0;       // background upload: the logs have landed once this staging mark is reached
    bool wait_upload = false;
    ~AdapterPlan() {
        for (void* p : owned) cudaFreeAsync(p, stream());
    }
    template <class T>
    int alloc(T*& p, size_t count) {
        void* q = nullptr;
        if (int e = cm31_malloc(&q, std::max<size_t>(count, 4) * sizeof(T))) return e;
        owned.push_back(q);
        p = (T*)q;
        return 0;
    }
};

static int grid_for(size_t n) { return (int)((n + 255) / 256); }

static int upload_opcode_tables() {
    static bool done[64] = {false};
    int dev = 0;
    CM_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && done[dev]) return 0;
    OpcodeTables t;
    for (u32 op = 0; op < ADAPTER_MAX_OPCODE; op++) {
        int n = opcode_memory_accesses(op);
        int size = opcode_size_in_m31s(op);
        t.n_acc[op] = n < 0 ? 0xFF : (uint8_t)n;
        t.size[op] = (uint8_t)size;
        t.n_inst[op] = size > 4 ? 2 : 1;
    }
    CM_CUDA(cudaMemcpyToSymbol(c_ops, &t, sizeof(t)));
    if (dev < 64) done[dev] = true;
    return 0;
}

static const char* adapter_error_text(u32 err) {
    if (err & ADAPTER_ERR_PC) return "adapter: pc outside the preloaded program memory";
    if (err & ADAPTER_ERR_OPCODE) return "adapter: invalid opcode";
    if (err & ADAPTER_ERR_END) return "adapter: unexpected end of the memory trace";
    if (err & ADAPTER_ERR_FETCH) return "adapter: unexpected memory access (instruction fetch does not match pc)";
    return "adapter: error";
}

}  // namespace cm31

using namespace cm31;

extern "C" {

// Phase 1: upload the logs, resolve everything whose size is data dependent.
//   trace_host : IoTraceEntry {fp, pc} x n_trace (one per step + the final state)
//   mem_host   : IoMemoryEntry {address, value[4]} x n_mem, access order
//   init_host  : preloaded memory, QM31 x n_init, address = index
// counts_out[0..63] = steps per opcode, [64] = data accesses, [65] = clock-update rows, [66] = distinct cells touched
// Phase 0: device buffers + the upload of the logs.  background != 0 issues the three copies on the background copy stream
// (ordered after the work already queued on the main stream) and returns at once: with page-locked logs the upload of
// segment i+1 overlaps the adapter kernels and the proof of segment i; cm31_adapter_scan_staged waits for it on the device.
int cm31_adapter_stage_logs(const uint32_t* trace_host, size_t n_trace, const uint32_t* mem_host, size_t n_mem, const uint32_t* init_host,
                            size_t n_init, int background, void** plan_out) {
    CM_REQUIRE(plan_out && trace_host && mem_host && init_host, "adapter_scan: null argument");
    CM_REQUIRE(n_trace >= 2, "adapter: empty trace");
    CM_REQUIRE(n_trace - 1 < (1u << 31) - 1 && n_mem < (1u << 31) - 1 && n_mem >= 1 && n_init >= 1 && n_init < (1u << 31), "adapter_scan: sizes out of range");
    if (int e = upload_opcode_tables()) return e;
    std::unique_ptr<AdapterPlan> pl(new AdapterPlan());
    AdapterPlan& P_ = *pl;
    u32 n = P_.n_steps = (u32)(n_trace - 1), M = P_.n_mem = (u32)n_mem;
    P_.n_init = (u32)n_init;
    int e = 0;
    if ((e = P_.alloc(P_.trace, n_trace)) || (e = P_.alloc(P_.mem, 5 * n_mem)) || (e = P_.alloc(P_.init, n_init)) || (e = P_.alloc(P_.opcode, n)) ||
        (e = P_.alloc(P_.cnt, (size_t)n + 1)) || (e = P_.alloc(P_.off, (size_t)n + 1)) || (e = P_.alloc(P_.ndata, (size_t)n + 1)) ||
        (e = P_.alloc(P_.dstart, (size_t)n + 1)) || (e = P_.alloc(P_.keys, M)) || (e = P_.alloc(P_.pay, M)) || (e = P_.alloc(P_.addr_sorted, M)) ||
        (e = P_.alloc(P_.iota, n)) || (e = P_.alloc(P_.pay_sorted, M)) || (e = P_.alloc(P_.steps_sorted, n)) || (e = P_.alloc(P_.op_sorted, n)) ||
        (e = P_.alloc(P_.prev, M)) || (e = P_.alloc(P_.n_cu, (size_t)M + 1)) ||
        (e = P_.alloc(P_.cu_off, (size_t)M + 1)) || (e = P_.alloc(P_.head, M)) || (e = P_.alloc(P_.head_incl, M)) || (e = P_.alloc(P_.small, 68)))
        return e;
    // temp storage for the CUB calls below (the largest request)
    size_t need = 0, t = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t, P_.cnt, P_.off, (int)(std::max(n, M) + 1), stream());
    need = std::max(need, t);
    cub::DeviceScan::InclusiveSum(nullptr, t, P_.head, P_.head_incl, (int)M, stream());
    need = std::max(need, t);
    cub::DeviceRadixSort::SortPairs(nullptr, t, P_.keys, P_.addr_sorted, P_.pay, P_.pay_sorted, (int)M, 0, 32, stream());
    need = std::max(need, t);
    cub::DeviceRadixSort::SortPairs(nullptr, t, P_.opcode, P_.op_sorted, P_.iota, P_.steps_sorted, (int)n, 0, 6, stream());
    need = std::max(need, t);
    P_.temp_bytes = need;
    if ((e = cm31_malloc(&P_.temp, need))) return e;
    P_.owned.push_back(P_.temp);

    if (background) {
        if ((e = cm31_bg_begin()) || (e = cm31_h2d_bg_ordered(P_.trace, trace_host, n_trace * 8)) || (e = cm31_h2d_bg_ordered(P_.init, init_host, n_init * 16)) ||
            (e = cm31_h2d_bg_ordered(P_.mem, mem_host, n_mem * 20)) || (e = cm31_bg_mark(&P_.upload_mark))) {
            cm31_bg_fence();  // copies already issued on the background stream still target these buffers: the plan's destructor
                              // frees them on the main stream, which must be ordered after those copies
            return e;
        }
        P_.wait_upload = true;
    } else {
        CM_CUDA(cudaMemcpyAsync(P_.trace, trace_host, n_trace * 8, cudaMemcpyHostToDevice, stream()));
        CM_CUDA(cudaMemcpyAsync(P_.mem, mem_host, n_mem * 20, cudaMemcpyHostToDevice, stream()));
        CM_CUDA(cudaMemcpyAsync(P_.init, init_host, n_init * 16, cudaMemcpyHostToDevice, stream()));
    }
    *plan_out = pl.release();
    return 0;
}

// Phase 1: resolve everything whose size is data dependent (see cm31_adapter_scan).
int cm31_adapter_scan_staged(void* plan, uint64_t counts_out[67]) {
    CM_REQUIRE(plan && counts_out, "adapter_scan: null argument");
    AdapterPlan& P_ = *(AdapterPlan*)plan;
    u32 n = P_.n_steps, M = P_.n_mem;
    if (P_.wait_upload) {
        if (int e = cm31_bg_wait(P_.upload_mark)) return e;
        P_.wait_upload = false;
    }
    CM_CUDA(cudaMemsetAsync(P_.small, 0, 68 * 4, stream()));
    CM_CUDA(cudaMemsetAsync(P_.cnt + n, 0, 4, stream()));
    CM_CUDA(cudaMemsetAsync(P_.ndata + n, 0, 4, stream()));
    CM_CUDA(cudaMemsetAsync(P_.n_cu, 0, ((size_t)M + 1) * 4, stream()));
    {
        ProfScope prof("adapter_count", 8ull * n + 12ull * n);
        adapter_count_kernel<<<grid_for(n), 256, 0, stream()>>>(P_.trace, n, P_.init, P_.n_init, P_.opcode, P_.cnt, P_.ndata, P_.small, P_.small + 64);
        CM_LAUNCH_CHECK();
    }
    {
        ProfScope prof("adapter_scan", 32ull * n, 0);
        size_t tb = P_.temp_bytes;
        CM_CUDA(cub::DeviceScan::ExclusiveSum(P_.temp, tb, P_.cnt, P_.off, (int)(n + 1), stream()));
        tb = P_.temp_bytes;
        CM_CUDA(cub::DeviceScan::ExclusiveSum(P_.temp, tb, P_.ndata, P_.dstart, (int)(n + 1), stream()));
    }
    {
        ProfScope prof("adapter_keys", 20ull * n + 8ull * M + 20ull * M, 2);
        adapter_keys_kernel<<<grid_for(n), 256, 0, stream()>>>(P_.trace, n, P_.mem, M, P_.opcode, P_.off, P_.cnt, P_.keys, P_.pay, P_.small + 65,
                                                               P_.small + 64);
        CM_LAUNCH_CHECK();
        adapter_iota_kernel<<<grid_for(n), 256, 0, stream()>>>(P_.iota, n);
        CM_LAUNCH_CHECK();
    }
    u32 small[68], totals[2];
    CM_CUDA(cudaMemcpyAsync(small, P_.small, 68 * 4, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaMemcpyAsync(&totals[0], P_.off + n, 4, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaMemcpyAsync(&totals[1], P_.dstart + n, 4, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    CM_REQUIRE(small[64] == 0, adapter_error_text(small[64]));
    CM_REQUIRE(totals[0] == M, "adapter: the memory trace does not end with the execution trace");
    P_.n_data = totals[1];
    u32 addr_bits = 1;
    while (addr_bits < 32 && (small[65] >> addr_bits)) addr_bits++;
    {
        ProfScope prof("adapter_sort", 40ull * M * ((addr_bits + 7) / 8) + 16ull * n, 0);
        size_t tb = P_.temp_bytes;
        CM_CUDA(cub::DeviceRadixSort::SortPairs(P_.temp, tb, P_.keys, P_.addr_sorted, P_.pay, P_.pay_sorted, (int)M, 0, (int)addr_bits, stream()));
        tb = P_.temp_bytes;
        CM_CUDA(cub::DeviceRadixSort::SortPairs(P_.temp, tb, P_.opcode, P_.op_sorted, P_.iota, P_.steps_sorted, (int)n, 0, 6, stream()));
    }
    {
        ProfScope prof("adapter_resolve", 20ull * M + 12ull * M);
        adapter_resolve_kernel<<<grid_for(M), 256, 0, stream()>>>(P_.addr_sorted, P_.pay_sorted, M, P_.init, P_.n_init, P_.prev, P_.n_cu, P_.head);
        CM_LAUNCH_CHECK();
    }
    {
        ProfScope prof("adapter_scan", 16ull * M, 0);
        size_t tb = P_.temp_bytes;
        CM_CUDA(cub::DeviceScan::ExclusiveSum(P_.temp, tb, P_.n_cu, P_.cu_off, (int)(M + 1), stream()));
        tb = P_.temp_bytes;
        CM_CUDA(cub::DeviceScan::InclusiveSum(P_.temp, tb, P_.head, P_.head_incl, (int)M, stream()));
    }
    CM_CUDA(cudaMemcpyAsync(&totals[0], P_.cu_off + M, 4, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaMemcpyAsync(&totals[1], P_.head_incl + (M - 1), 4, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    P_.n_cu_rows = totals[0];
    P_.n_cells = totals[1];
    for (u32 op = 0; op < ADAPTER_MAX_OPCODE; op++) counts_out[op] = small[op];
    counts_out[64] = P_.n_data;
    counts_out[65] = P_.n_cu_rows;
    counts_out[66] = P_.n_cells;
    return 0;
}

int cm31_adapter_free(void* plan);
int cm31_adapter_scan(const uint32_t* trace_host, size_t n_trace, const uint32_t* mem_host, size_t n_mem, const uint32_t* init_host,
                      size_t n_init, void** plan_out, uint64_t counts_out[67]) {
    CM_REQUIRE(plan_out && counts_out, "adapter_scan: null argument");
    void* plan = nullptr;
    if (int e = cm31_adapter_stage_logs(trace_host, n_trace, mem_host, n_mem, init_host, n_init, 0, &plan)) return e;
    if (int e = cm31_adapter_scan_staged(plan, counts_out)) {
        cm31_adapter_free(plan);
        return e;
    }
    *plan_out = plan;
    return 0;
}

// Phase 2: write the outputs into buffers sized from phase 1.
//   opcode_rows_dev[op] : device buffer where the bundles of opcode `op` go (12 words per step), null iff it has no steps
//   accesses_dev        : 4 words per data access;  clock_update_dev : 6 words per row
//   cells_host          : 10 words per distinct cell {address, first value[4], last value[4], last clock}, ascending address
int cm31_adapter_emit(void* plan, uint32_t* const opcode_rows_dev[64], const uint64_t counts[67], uint32_t* accesses_dev,
                      uint32_t* clock_update_dev, uint32_t* cells_host) {
    CM_REQUIRE(plan && opcode_rows_dev && counts && accesses_dev && cells_host, "adapter_emit: null argument");
    AdapterPlan& P_ = *(AdapterPlan*)plan;
    u32 n = P_.n_steps, M = P_.n_mem;
    AdapterDest dest;
    u32 at = 0;
    for (u32 op = 0; op < ADAPTER_MAX_OPCODE; op++) {
        dest.base[op] = opcode_rows_dev[op];
        dest.start[op] = at;
        CM_REQUIRE(counts[op] == 0 || opcode_rows_dev[op] != nullptr, "adapter: unimplemented opcode (no component takes its steps)");
        at += (u32)counts[op];
    }
    CM_REQUIRE(at == n, "adapter_emit: opcode counts do not add up to the number of steps");
    DeviceTable dt;
    if (int e = dt.upload(&dest, sizeof(dest))) return e;
    {
        ProfScope prof("adapter_bundles", 8ull * n + 20ull * n + 48ull * n);
        adapter_bundles_kernel<<<grid_for(n), 256, 0, stream()>>>(P_.steps_sorted, n, P_.trace, P_.opcode, P_.off, P_.dstart, P_.mem, P_.prev,
                                                                  (const AdapterDest*)dt.d);
        CM_LAUNCH_CHECK();
    }
    {
        ProfScope prof("adapter_accesses", 12ull * n + 32ull * P_.n_data);
        adapter_accesses_kernel<<<grid_for(n), 256, 0, stream()>>>(n, P_.opcode, P_.off, P_.dstart, P_.mem, P_.prev, (uint4*)accesses_dev);
        CM_LAUNCH_CHECK();
    }
    if (P_.n_cu_rows) {
        CM_REQUIRE(clock_update_dev != nullptr, "adapter_emit: null clock-update buffer");
        ProfScope prof("adapter_clock_update", 8ull * M + 24ull * P_.n_cu_rows);
        adapter_clock_update_kernel<<<grid_for(M), 256, 0, stream()>>>(P_.addr_sorted, P_.pay_sorted, M, P_.mem, P_.init, P_.n_init, P_.n_cu,
                                                                       P_.cu_off, clock_update_dev);
        CM_LAUNCH_CHECK();
    }
    u32* cells = nullptr;
    if (int e = P_.alloc(cells, 10 * (size_t)P_.n_cells)) return e;
    {
        ProfScope prof("adapter_cells", 8ull * M + 40ull * P_.n_cells);
        adapter_cells_kernel<<<grid_for(M), 256, 0, stream()>>>(P_.addr_sorted, P_.pay_sorted, M, P_.mem, P_.head_incl, cells);
        CM_LAUNCH_CHECK();
    }
    CM_CUDA(cudaMemcpyAsync(cells_host, cells, 40 * (size_t)P_.n_cells, cudaMemcpyDeviceToHost, stream()));
    CM_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

int cm31_adapter_free(void* plan) {
    delete (AdapterPlan*)plan;
    return 0;
}

}  // extern "C"

// Shared runtime pieces of libcm31: error reporting, stream, device scratch for pointer tables.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <string>
#include <vector>

#include "../../include/cm31.h"
#include "field.cuh"

namespace cm31 {

void set_error(const std::string& msg);
cudaStream_t stream();
// shard.cu: while a sharded proof is being made cm31_malloc bump-allocates from the peer-mapped arena
bool shard_arena_alloc(void** out, size_t bytes, int* status);
bool shard_arena_owns(const void* p);

#define CM_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            cm31::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));               \
            return (int)_e;                                                                    \
        }                                                                                      \
    } while (0)

#define CM_REQUIRE(cond, msg)                                    \
    do {                                                         \
        if (!(cond)) {                                           \
            cm31::set_error(std::string("cm31: ") + (msg));      \
            return -1;                                           \
        }                                                        \
    } while (0)

#define CM_LAUNCH_CHECK() CM_CUDA(cudaGetLastError())

// Per-kernel device timing (cm31_profile_*): when enabled, a ProfScope brackets one launch with
// two CUDA events on the launch stream and books `alg_bytes` ALGORITHMIC bytes
// to the kernel's name.  Disabled (default) it costs one branch.
void prof_begin(const char* name, uint64_t alg_bytes, unsigned n_kernels);
void prof_end();
// books ALGORITHMIC M31 operations (SURVEY.md §8d "ALGORITHMIC ops": butterfly = 3, QM31 mul = 31, QM31 x M31 = 4, ...) to the
// launch bracketed by the innermost open ProfScope; a no-op when timing is off
void prof_ops(uint64_t m31_ops);
bool prof_enabled();
struct ProfScope {
    ProfScope(const char* name, uint64_t alg_bytes, unsigned n_kernels = 1) { prof_begin(name, alg_bytes, n_kernels); }
    ~ProfScope() { prof_end(); }
};

// Uploads a small host table (pointer arrays, programs, constants): bump-allocated from a device
// ring mirrored by a pinned host ring (one async DMA per table, no allocation calls); slots are
// recycled when the ring wraps, after a stream synchronise.
struct DeviceTable {
    void* d = nullptr;
    bool owned = false;
    int upload(const void* host, size_t bytes);
    // two-step form for tables that embed their own device address: reserve() fixes `d`, fill() copies the bytes
    int reserve(size_t bytes);
    int fill(const void* host, size_t bytes);
    size_t ring_at = 0;
    void release();
    ~DeviceTable() { release(); }
};

struct cm31_twiddles_impl {
    uint32_t log_size;  // max circle-domain log size served
    uint32_t* tw;       // 2^(log_size-1) words
    uint32_t* itw;
};

}  // namespace cm31

struct cm31_twiddles : cm31::cm31_twiddles_impl {};

// Trace-fill input staging on sm_100a: AoS prover inputs (ExecutionBundles, data-access log,
// boundary-memory rows) -> SoA input columns consumed by the trace-fill AIR programs.
//
// Replaces the per-16-row packing of the reference
//   Pack::pack for ExecutionBundle   crates/prover/src/utils/execution_bundle.rs:31-75
//   get_access_field                 crates/prover/src/utils/data_accesses.rs:10-28
//   padding with ExecutionBundle::default() (a Ret)  crates/prover/src/adapter/memory.rs:112-124,
//                                                    components/opcodes/store_fp_fp.rs:169
// One thread per (row); bundle words are read with 128-bit loads, the access log is gathered
// through the per-step span, every output column is written coalesced.
#include <cstdlib>

#include "air/cairo_components.hpp"
#include "common.cuh"

namespace cm31 {

// the output column pointers travel as a kernel parameter (no table upload: this kernel runs once per opcode component, most of
// them 16-row paddings whose cost is the host's launch path)
struct UnpackOut {
    u32* p[N_BUNDLE_INPUTS];
};
__global__ void __launch_bounds__(256) unpack_bundles_kernel(const uint4* __restrict__ bundles, u32 n_real, u32 log_size,
                                                             const uint4* __restrict__ accesses, u32 n_accesses,
                                                             const __grid_constant__ UnpackOut outs, u32 n_slots) {
    u32* const* out = outs.p;
    u32 row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (1u << log_size)) return;
    u32 w[12];
    if (row < n_real) {
        uint4 a = __ldg(bundles + 3 * (size_t)row), b = __ldg(bundles + 3 * (size_t)row + 1), c = __ldg(bundles + 3 * (size_t)row + 2);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
    } else {
#pragma unroll
        for (int k = 0; k < 12; k++) w[k] = 0;
        w[4] = OP_RET;  // ExecutionBundle::default(): a Ret with zero registers/clock and an empty span
    }
#pragma unroll
    for (int k = 0; k < 10; k++) out[k][row] = w[k];
    u32 start = w[10], len = w[11];
#pragma unroll
    for (int k = 0; k < MAX_ACCESSES; k++) {
        if ((u32)k >= n_slots) break;  // access slots the component's trace program never reads are not written
        uint4 acc = make_uint4(0, 0, 0, 0);
        if ((u32)k < len && start + k < n_accesses) acc = __ldg(accesses + start + k);
        out[IN_ACC_BASE + 4 * k + ACC_ADDRESS][row] = acc.x;
        out[IN_ACC_BASE + 4 * k + ACC_PREV_CLOCK][row] = acc.y;
        out[IN_ACC_BASE + 4 * k + ACC_PREV_VALUE][row] = acc.z;
        out[IN_ACC_BASE + 4 * k + ACC_VALUE][row] = acc.w;
    }
}

__global__ void unpack_rows_kernel(const u32* __restrict__ rows, u32 n_real, u32 n_fields, u32 log_size, u32* const* __restrict__ out) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t n = (size_t)1 << log_size;
    if (t >= n * n_fields) return;
    u32 f = (u32)(t / n), row = (u32)(t % n);
    out[f][row] = row < n_real ? __ldg(rows + (size_t)row * n_fields + f) : 0u;
}

__global__ void bitwise_table_kernel(int k, u32* col) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (1u << BITWISE_STACKED_LOG_SIZE)) col[i] = bitwise_table_value(k, i);
}

__global__ void iota_kernel(u32* col, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) col[i] = (u32)i;
}

}  // namespace cm31

using namespace cm31;

extern "C" {

int cm31_unpack_bundles_slots(const uint32_t* bundles_dev, size_t n_real, uint32_t log_size, const uint32_t* accesses_dev,
                              size_t n_accesses, uint32_t* const* out_cols, uint32_t n_access_slots);
int cm31_unpack_bundles(const uint32_t* bundles_dev, size_t n_real, uint32_t log_size, const uint32_t* accesses_dev,
                        size_t n_accesses, uint32_t* const* out_cols) {
    return cm31_unpack_bundles_slots(bundles_dev, n_real, log_size, accesses_dev, n_accesses, out_cols, MAX_ACCESSES);
}
// n_access_slots: only the first n_access_slots x 4 access columns are written (a store_fp_imm step has 2 accesses, a jump
// none: the other columns of the 8-slot layout are never read by that component's trace program)
int cm31_unpack_bundles_slots(const uint32_t* bundles_dev, size_t n_real, uint32_t log_size, const uint32_t* accesses_dev,
                              size_t n_accesses, uint32_t* const* out_cols, uint32_t n_access_slots) {
    CM_REQUIRE(log_size <= 30 && n_real <= ((size_t)1 << log_size), "unpack_bundles: more rows than the padded size");
    CM_REQUIRE(n_access_slots <= (uint32_t)MAX_ACCESSES, "unpack_bundles: at most 8 access slots");
    static const bool full = getenv("CM31_FULL_UNPACK") != nullptr;  // A/B runs: write all 8 slots as the first version did
    if (full) n_access_slots = MAX_ACCESSES;
    UnpackOut outs;
    for (int k = 0; k < N_BUNDLE_INPUTS; k++) outs.p[k] = out_cols[k];
    size_t n = (size_t)1 << log_size;
    ProfScope prof("unpack_bundles", 48ull * n_real + 4ull * (IN_ACC_BASE + 4ull * n_access_slots) * n);
    unpack_bundles_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>((const uint4*)bundles_dev, (u32)n_real, log_size,
                                                                            (const uint4*)accesses_dev, (u32)n_accesses, outs, n_access_slots);
    CM_LAUNCH_CHECK();
    return 0;
}

int cm31_unpack_rows(const uint32_t* rows_dev, size_t n_real, uint32_t n_fields, uint32_t log_size, uint32_t* const* out_cols) {
    CM_REQUIRE(log_size <= 30 && n_real <= ((size_t)1 << log_size), "unpack_rows: more rows than the padded size");
    DeviceTable dout;
    if (int e = dout.upload(out_cols, n_fields * sizeof(void*))) return e;
    size_t total = ((size_t)1 << log_size) * n_fields;
    ProfScope prof("unpack_rows", 8ull * total);
    unpack_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream()>>>(rows_dev, (u32)n_real, n_fields, log_size, (u32* const*)dout.d);
    CM_LAUNCH_CHECK();
    return 0;
}

// preprocessed column k of the stacked bitwise table (crates/prover/src/preprocessed/bitwise.rs:253-290)
int cm31_bitwise_table_col(int k, uint32_t* col) {
    CM_REQUIRE(k >= 0 && k < 4 && col != nullptr, "bitwise_table_col: bad argument");
    ProfScope prof("bitwise_table", 4ull << BITWISE_STACKED_LOG_SIZE);
    bitwise_table_kernel<<<(1u << BITWISE_STACKED_LOG_SIZE) / 256, 256, 0, stream()>>>(k, col);
    CM_LAUNCH_CHECK();
    return 0;
}

int cm31_iota(uint32_t* col, size_t n) {
    if (n == 0) return 0;
    ProfScope prof("iota", 4ull * n);
    iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>(col, n);
    CM_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

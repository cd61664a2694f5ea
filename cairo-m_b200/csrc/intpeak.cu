// Integer issue-rate microbenchmark: the roof that actually binds this path.
//
// MEASURED_PEAKS.json has HBM and bf16-tensor peaks only; M31 arithmetic is 32-bit integer work
// that issues on two pipes (ALU: IADD3/LOP3/SHF/VIADDMNMX, FMA: IMAD).  Three kernels measure
// lane-operations per second for (0) an ALU-pipe-only chain, (1) an IMAD-only chain and (2) the
// 1:1 mix, with 8 independent chains per thread and every SM saturated.  bench.py divides each
// kernel's executed thread-instructions per second by the mixed figure.
#include "common.cuh"

namespace cm31 {

constexpr int INTPEAK_ITERS = 4096;
constexpr int INTPEAK_OPS_PER_ITER = 16;  // per thread per iteration

template <int MODE>
__global__ void __launch_bounds__(256) intpeak_kernel(u32* out, u32 a, u32 b) {
    u32 x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x * 8 + k + a;
#pragma unroll 1
    for (int it = 0; it < INTPEAK_ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (MODE == 0) {  // two ALU-pipe ops: LOP3, SHF (funnel rotate)
                x[k] = (x[k] ^ b) & a;
                x[k] = __funnelshift_l(x[k], x[k], 7);
            } else if (MODE == 1) {  // two FMA-pipe ops: IMAD x2
                x[k] = x[k] * a + b;
                x[k] = x[k] * b + a;
            } else {  // one of each
                x[k] = x[k] * a + b;
                x[k] = __funnelshift_l(x[k], x[k], 7);
            }
        }
    }
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace cm31

using namespace cm31;

extern "C" int cm31_int_peak(double tera_lane_ops_per_s[3]) {
    CM_REQUIRE(tera_lane_ops_per_s != nullptr, "int_peak: null output");
    int dev = 0, sms = 0;
    CM_CUDA(cudaGetDevice(&dev));
    CM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned blocks = (unsigned)sms * 8, threads = 256;
    u32* out = nullptr;
    CM_CUDA(cudaMallocAsync(&out, (size_t)blocks * threads * 4, stream()));
    cudaEvent_t e0, e1;
    CM_CUDA(cudaEventCreate(&e0));
    CM_CUDA(cudaEventCreate(&e1));
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {  // first repetition warms up
            CM_CUDA(cudaEventRecord(e0, stream()));
            if (mode == 0) intpeak_kernel<0><<<blocks, threads, 0, stream()>>>(out, 0x7ffffff1u + rep, 0x9e3779b9u);
            else if (mode == 1) intpeak_kernel<1><<<blocks, threads, 0, stream()>>>(out, 0x7ffffff1u + rep, 0x9e3779b9u);
            else intpeak_kernel<2><<<blocks, threads, 0, stream()>>>(out, 0x7ffffff1u + rep, 0x9e3779b9u);
            CM_CUDA(cudaEventRecord(e1, stream()));
            CM_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            CM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        double ops = (double)blocks * threads * INTPEAK_ITERS * INTPEAK_OPS_PER_ITER;
        tera_lane_ops_per_s[mode] = ops / (best * 1e-3) / 1e12;
    }
    CM_LAUNCH_CHECK();
    CM_CUDA(cudaEventDestroy(e0));
    CM_CUDA(cudaEventDestroy(e1));
    CM_CUDA(cudaFreeAsync(out, stream()));
    return 0;
}

// Proof-level entry points of libcm31 (host protocol driver over the CUDA backend ops).
#include <cstring>

#include "common.cuh"
#include "host/cuda_backend.hpp"
#include "host/framework.hpp"
#include "host/test_provers.hpp"

using namespace cm31;

static int write_out(const std::vector<uint8_t>& bytes, uint8_t* out, size_t cap, size_t* out_len) {
    if (out_len) *out_len = bytes.size();
    CM_REQUIRE(out == nullptr || bytes.size() <= cap, "proof buffer too small");
    if (out) memcpy(out, bytes.data(), bytes.size());
    return 0;
}

extern "C" {

int cm31_prove_wide_fibonacci(uint32_t log_n_rows, uint32_t n_cols, uint32_t pow_bits, uint32_t n_queries,
                              uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    try {
        PcsConfig cfg;
        cfg.pow_bits = pow_bits;
        cfg.fri_config.n_queries = n_queries;
        StarkProof proof = prove_wide_fibonacci<CudaBackend, FrameworkComponent<CudaBackend, WideFibonacciEval>>(log_n_rows, n_cols, cfg);
        ProofWriter w;
        w.proof(proof);
        return write_out(w.bytes, proof_out, proof_cap, proof_len);
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}

}  // extern "C"

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

// ------------------------------------------------------------------ cairo-m
#include <deque>

#include "cairo/prover.hpp"
#include "host/cuda_air_impl.hpp"

struct cm31_prover_input {
    cm31::ProverInput input;
    uint32_t return_value = 0;
    std::unique_ptr<cm31::StagedInput<cm31::CudaAirImpl>> staged;  // set by cm31_input_upload
    std::deque<std::unique_ptr<cm31::StagedInput<cm31::CudaAirImpl>>> prefetched;  // cm31_input_prefetch, consumed oldest first
    std::vector<void*> pinned;                                     // host ranges registered with CUDA
    ~cm31_prover_input() {
        for (void* p : pinned) cudaHostUnregister(p);
    }
};

static void pin_range(cm31_prover_input* h, const void* p, size_t bytes) {
    // page-lock the adapter's big vectors so the per-proof H2D copies run at PCIe/C2C speed;
    // best effort (no device here, or the range is tiny): a failure only means a slower copy.
    if (bytes < (1u << 16)) return;
    if (cudaHostRegister((void*)p, bytes, cudaHostRegisterDefault) == cudaSuccess) h->pinned.push_back((void*)p);
    else cudaGetLastError();
}

extern "C" {

// Runs the (host) VM + adapter for fibonacci_loop(n): the prover input of crates/prover/src/adapter.
int cm31_program_input_create(uint32_t program_id, uint32_t n, cm31_prover_input** out);
int cm31_fib_input_create(uint32_t n, cm31_prover_input** out) { return cm31_program_input_create(PROGRAM_FIBONACCI_LOOP, n, out); }
// program_id: 0 = fibonacci_loop(n), 1 = array_sum(n) (call / frame-pointer / double-deref / assert / le opcodes), 2 = u32_counter(n), 3 = u32_mix(n) (u32 mul / divrem / eq / lt and the two-word *_fp_imm u32 instructions)
int cm31_program_input_create(uint32_t program_id, uint32_t n, cm31_prover_input** out) {
    try {
        CM_REQUIRE(out != nullptr, "program_input_create: null out");
        VmTrace vm = run_program(program_by_id(program_id), n);
        cm31_prover_input* h = new cm31_prover_input();
        h->input = import_from_vm(vm);
        h->return_value = vm.return_value;
        pin_range(h, h->input.data_accesses.data(), h->input.data_accesses.size() * sizeof(DataAccess));
        for (auto& kv : h->input.states_by_opcodes) pin_range(h, kv.second.data(), kv.second.size() * sizeof(Bundle));
        *out = h;
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}
// Test hook: corrupt one value of the adapter output so that a POLYNOMIAL constraint of the store_fp_fp
// AIR (dst_val - res, store_fp_fp.rs evaluate) no longer holds; the prover must then refuse
// (ConstraintsNotSatisfied, stwo prover/mod.rs:76-82).  A corruption that only unbalances a lookup would
// still prove — as in the reference — and be caught by the verifier's logup-sum check instead.
//   kind 0: the value written by the middle StoreAddFpFp step;  kind 1: the second operand read by the middle StoreSubFpFp step
int cm31_input_tamper(cm31_prover_input* h, uint32_t kind) {
    CM_REQUIRE(h != nullptr, "input_tamper: null handle");
    h->staged.reset();
    h->prefetched.clear();
    auto it = h->input.states_by_opcodes.find(kind == 0 ? OP_STORE_ADD_FP_FP : OP_STORE_SUB_FP_FP);
    CM_REQUIRE(it != h->input.states_by_opcodes.end() && !it->second.empty(), "input_tamper: the program has no such step");
    const Bundle& b = it->second[it->second.size() / 2];
    CM_REQUIRE(b.span_len == 3, "input_tamper: unexpected access span");
    DataAccess& a = h->input.data_accesses[b.span_start + (kind == 0 ? 2 : 1)];
    a.value = m31_add(a.value, 1);
    return 0;
}
int cm31_input_destroy(cm31_prover_input* h) {
    delete h;
    return 0;
}
// Copies the prover input to HBM once; later cm31_prove_cairo_m calls on this handle skip the
// host->device copy (bench.py's device-resident `value`; without it every proof stages its input).
int cm31_input_upload(cm31_prover_input* h) {
    try {
        CM_REQUIRE(h != nullptr, "input_upload: null handle");
        h->staged.reset(new StagedInput<CudaAirImpl>(stage_input<CudaAirImpl>(h->input)));
        return cm31_sync();
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}
int cm31_input_release_device(cm31_prover_input* h) {
    CM_REQUIRE(h != nullptr, "input_release_device: null handle");
    h->staged.reset();
    h->prefetched.clear();
    return 0;
}
int cm31_input_prefetch(cm31_prover_input* h) {
    try {
        CM_REQUIRE(h != nullptr, "input_prefetch: null handle");
        CM_REQUIRE(h->prefetched.size() < 4, "input_prefetch: too many uploads in flight");
        h->prefetched.emplace_back(new StagedInput<CudaAirImpl>(stage_input<CudaAirImpl>(h->input)));
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}
// info[0] = VM steps, info[1] = data accesses, info[2] = boundary memory rows, info[3] = return value,
// info[4] = bytes of prover input copied host->device per proof
int cm31_input_info(const cm31_prover_input* h, uint64_t info[5]) {
    CM_REQUIRE(h != nullptr, "input_info: null handle");
    info[0] = h->input.n_steps;
    info[1] = h->input.data_accesses.size();
    info[2] = h->input.initial_memory.size() + h->input.final_memory.size();
    info[3] = h->return_value;
    info[4] = h->input.n_steps * sizeof(Bundle) + h->input.data_accesses.size() * sizeof(DataAccess) + info[2] * 32 +
              h->input.clock_update_data.size() * 24;
    return 0;
}

// prove_cairo_m::<Blake2sMerkleChannel> (crates/prover/src/prover.rs:23) on the CUDA backend.
// timings_ms (optional, 5 doubles): preprocessed, trace, interaction, stark, total.
int cm31_prove_cairo_m(const cm31_prover_input* h, uint32_t pow_bits, uint32_t n_queries, uint8_t* proof_out,
                       size_t proof_cap, size_t* proof_len, double* timings_ms) {
    try {
        CM_REQUIRE(h != nullptr, "prove_cairo_m: null input");
        PcsConfig cfg = PcsConfig::regular_96_bits();
        cfg.pow_bits = pow_bits;
        cfg.fri_config.n_queries = n_queries;
        ProveTimings t;
        std::unique_ptr<StagedInput<CudaAirImpl>> pre;
        if (!h->staged && !h->prefetched.empty()) {
            pre = std::move(const_cast<cm31_prover_input*>(h)->prefetched.front());
            const_cast<cm31_prover_input*>(h)->prefetched.pop_front();
        }
        CairoProof proof = h->staged ? prove_cairo_m<CudaAirImpl>(h->input, *h->staged, cfg, &t)
                           : pre     ? prove_cairo_m<CudaAirImpl>(h->input, *pre, cfg, &t)
                                     : prove_cairo_m<CudaAirImpl>(h->input, cfg, &t);
        if (timings_ms) {
            timings_ms[0] = t.preprocessed_ms;
            timings_ms[1] = t.trace_ms;
            timings_ms[2] = t.interaction_ms;
            timings_ms[3] = t.stark_ms;
            timings_ms[4] = t.total_ms;
        }
        int rc;
        {
            HostTimer ht("proof_to_bytes");
            rc = write_out(proof.to_bytes(), proof_out, proof_cap, proof_len);
        }
        HostTimer::report();
        return rc;
    } catch (const std::exception& e) {
        cm31_lanes_join();  // an error may have been raised while the side lane was current
        set_error(e.what());
        return -2;
    }
}

}  // extern "C"

// Backend-generic prove/verify drivers of the bring-up AIRs (shared by the CUDA product library
// and the CPU oracle, each instantiating it with its own backend + component type).
// Mirrors external/stwo/crates/examples/src/wide_fibonacci/mod.rs:172-229 (test_wide_fib_prove_with_blake).
#pragma once
#include "../air/wide_fibonacci.hpp"
#include "framework.hpp"
#include "stark.hpp"

namespace cm31 {

template <class B, class Component>
StarkProof prove_wide_fibonacci(u32 log_n_rows, u32 n_cols, PcsConfig config) {
    typename B::Twiddles twiddles;
    B::precompute_twiddles(log_n_rows + 1 + config.fri_config.log_blowup_factor, twiddles);
    Blake2sChannel channel;
    CommitmentSchemeProver<B> commitment_scheme(config, &twiddles);
    // preprocessed trace: empty tree
    commitment_scheme.commit_evals({}, channel);
    // trace
    std::vector<std::vector<u32>> host_trace = wide_fibonacci_trace(log_n_rows, n_cols);
    std::vector<CircleEvaluation<B>> trace;
    for (auto& col : host_trace) trace.push_back(CircleEvaluation<B>{B::from_host(col.data(), col.size()), log_n_rows});
    commitment_scheme.commit_evals(std::move(trace), channel);
    RelationSet relations;
    Component component(WideFibonacciEval{log_n_rows, n_cols}, &relations);
    TraceLocationAllocator alloc;
    component.allocate(alloc);
    std::vector<const ComponentProver<B>*> comps = {&component};
    return prove<B>(comps, channel, commitment_scheme);
}

}  // namespace cm31

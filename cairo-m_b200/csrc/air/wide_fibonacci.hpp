// WideFibonacci bring-up AIR: every row holds a Fibonacci-like sequence c = a^2 + b^2.
// Mirrors external/stwo/crates/examples/src/wide_fibonacci/mod.rs:22-66 (eval + generate_trace);
// used to exercise the whole commit -> composition -> OODS -> FRI pipeline before the cairo-m AIRs.
#pragma once
#include <vector>

#include "../field.cuh"

namespace cm31 {

struct WideFibonacciEval {
    u32 log_n_rows;
    u32 n_cols;
    u32 log_size() const { return log_n_rows; }
    u32 max_constraint_log_degree_bound() const { return log_n_rows + 1; }
    long cache_tag() const { return n_cols; }
    template <class E>
    void evaluate(E& eval) const {
        auto a = eval.next_trace_mask();
        auto b = eval.next_trace_mask();
        for (u32 i = 2; i < n_cols; i++) {
            auto c = eval.next_trace_mask();
            eval.add_constraint(c - (a * a + b * b));
            a = b;
            b = c;
        }
    }
};

// generate_test_trace (wide_fibonacci/mod.rs:105-132): a = 1, b = row index.
inline std::vector<std::vector<u32>> wide_fibonacci_trace(u32 log_n_rows, u32 n_cols) {
    size_t n = (size_t)1 << log_n_rows;
    std::vector<std::vector<u32>> trace(n_cols, std::vector<u32>(n));
    for (size_t r = 0; r < n; r++) {
        u32 a = 1, b = (u32)r;
        trace[0][r] = a;
        trace[1][r] = b;
        for (u32 c = 2; c < n_cols; c++) {
            u32 nb = m31_add(m31_sqr(a), m31_sqr(b));
            a = b;
            b = nb;
            trace[c][r] = b;
        }
    }
    return trace;
}

}  // namespace cm31

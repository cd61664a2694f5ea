// CudaAirImpl: the product's implementation choices for the cairo-m protocol driver
// (cairo/prover.hpp): CudaBackend ops + bytecode-driven components + device-side trace fill.
// Replaces the SimdBackend-concrete pieces listed in SURVEY.md §7 H2: `write_trace`
// (crates/prover/src/components/opcodes/*.rs), `LogupTraceGenerator`, the multiplicity histograms.
#pragma once
#include <chrono>
#include <map>

#include "../../../include/cm31.h"
#include "../cairo/vm.hpp"
#include "cuda_backend.hpp"
#include "framework.hpp"

extern "C" {
int cm31_bitwise_table_col(int k, uint32_t* col);
int cm31_unpack_bundles(const uint32_t* bundles_dev, size_t n_real, uint32_t log_size, const uint32_t* accesses_dev,
                        size_t n_accesses, uint32_t* const* out_cols);
int cm31_unpack_rows(const uint32_t* rows_dev, size_t n_real, uint32_t n_fields, uint32_t log_size, uint32_t* const* out_cols);
int cm31_iota(uint32_t* col, size_t n);
}

namespace cm31 {

struct CudaAirImpl {
    typedef CudaBackend B;
    typedef DeviceCol Col;
    template <class Eval>
    using Component = FrameworkComponent<CudaBackend, Eval>;
    typedef DeviceCol Words;  // staged prover-input records (u32 words) in HBM
    static double now_ms() {
        cm_check(cm31_sync());
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
    static Col iota(size_t n) {
        Col c(n);
        cm_check(cm31_iota(c.ptr(), n));
        return c;
    }
    static Col bitwise_table_col(int k) {
        Col c((size_t)1 << BITWISE_STACKED_LOG_SIZE);
        cm_check(cm31_bitwise_table_col(k, c.ptr()));
        return c;
    }
    static Col clone(const Col& c) {
        Col o(c.size());
        cm_check(cm31_d2d(o.ptr(), c.ptr(), c.size() * 4));
        return o;
    }
    static Words alloc_words(size_t n_words) { return DeviceCol(std::max<size_t>(4, n_words)); }
    // staging runs on the background copy stream; prove_cairo_m fences before the first use
    static void staging_begin() { cm_check(cm31_bg_begin()); }
    static void copy_words(Words& dst, size_t at, const u32* src, size_t n_words) {
        if (n_words) cm_check(cm31_h2d_bg_ordered(dst.ptr() + at, src, n_words * 4));
    }
    static Words upload_words(const u32* src, size_t n_words) {
        DeviceCol dev(std::max<size_t>(4, n_words));
        // small row tables come from temporaries: copy them synchronously-staged on the main stream
        if (n_words) cm_check(cm31_h2d(dev.ptr(), src, n_words * 4));
        return dev;
    }
    static void staging_fence() { cm_check(cm31_bg_fence()); }
    static u32 staging_mark() {
        u32 m = 0;
        cm_check(cm31_bg_mark(&m));
        return m;
    }
    static void staging_wait(u32 mark) { cm_check(cm31_bg_wait(mark)); }
    static std::vector<Col> unpack_bundles(const Words& rows, size_t n_real, const Words& accesses, size_t n_accesses, u32 log_size) {
        std::vector<Col> cols = Col::many(N_BUNDLE_INPUTS, (size_t)1 << log_size);
        std::vector<u32*> p;
        for (auto& c : cols) p.push_back(c.ptr());
        cm_check(cm31_unpack_bundles(rows.ptr(), n_real, log_size, accesses.ptr(), n_accesses, p.data()));
        return cols;
    }
    static std::vector<Col> unpack_rows(const Words& rows, size_t n_real, u32 n_fields, u32 log_size) {
        std::vector<Col> cols = Col::many(n_fields, (size_t)1 << log_size);
        std::vector<u32*> p;
        for (auto& c : cols) p.push_back(c.ptr());
        cm_check(cm31_unpack_rows(rows.ptr(), n_real, n_fields, log_size, p.data()));
        return cols;
    }
    template <class Eval>
    static std::vector<CircleEvaluation<B>> write_trace(const Eval& eval, const std::vector<Col>& inputs, u32 n_real) {
        // the trace-fill program of an AIR is captured once; only the enabler's row bound changes
        static std::map<long, AirProgram> cache;
        long tag = (long)air_cache_tag(eval, 0);
        auto it = cache.find(tag);
        if (it == cache.end()) {
            TraceProgramBuilder tb(n_real);
            eval.write_trace(tb);
            it = cache.emplace(tag, tb.compile()).first;
        }
        AirProgram prog = it->second;
        for (u32 slot : prog.rowlt_slots) prog.consts[slot] = n_real;
        std::vector<CircleEvaluation<B>> out(Eval::N_TRACE_COLUMNS);
        std::vector<Col> slab = Col::many(Eval::N_TRACE_COLUMNS, (size_t)1 << eval.log_size());
        std::vector<Col*> outp;
        for (size_t i = 0; i < out.size(); i++) {
            out[i].values = std::move(slab[i]);
            out[i].log_size = eval.log_size();
            outp.push_back(&out[i].values);
        }
        std::vector<const Col*> in;
        for (auto& c : inputs) in.push_back(&c);
        B::air_program(in, outp, eval.log_size(), prog);
        return out;
    }
    // one device->host copy for the claimed sums of every component (InteractionClaim, components/mod.rs:288-302)
    template <class Components>
    static void collect_claimed_sums(Components& components) {
        std::vector<QM31> sums = B::collect_sums();
        components.for_each([&](auto& c) {
            if (c.pending_sum_slot >= 0) {
                c.claimed_sum = sums.at((size_t)c.pending_sum_slot);
                c.pending_sum_slot = -1;
            }
        });
    }
    template <class Comp>
    static void emit_lookups(Comp& comp, int relation, const std::vector<const Col*>& trace_cols, Col& bins) {
        comp.emit_lookups(relation, trace_cols, bins);
    }
};

}  // namespace cm31

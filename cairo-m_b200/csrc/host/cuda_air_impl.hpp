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
    // Where a running proof lets the deferred bulk copy of the NEXT input start (cm31_bg_defer / cm31_bg_release), and how:
    // point 0 = at its own start, 1 = before the commitment of the execution traces, 2 = before the commitment of the
    // interaction traces, 3 = before the STARK phase (default), -1 = copies are never deferred (cm31_input_prefetch issues
    // them); CM31_PREFETCH_CTAS = CTAs of the throttled copy kernel (default 16, 0 = copy-engine DMA).
    // Same-box A/B, 2^22-step proofs, device-resident step 22.3 ms: DMA at prefetch time 25.9 ms per end-to-end step, DMA
    // deferred to the STARK phase 25.3, 16-CTA copy from the proof's start 24.5, 12-CTA 24.3, 16-CTA from the STARK phase 24.1.
    static int prefetch_point() {
        static const int at = getenv("CM31_PREFETCH_AT") ? atoi(getenv("CM31_PREFETCH_AT")) : 3;
        return at;
    }
    static int prefetch_ctas() {
        static const int n = getenv("CM31_PREFETCH_CTAS") ? atoi(getenv("CM31_PREFETCH_CTAS")) : 16;
        return n;
    }
    static void idle_gate_open() { B::idle_gate_open(); }
    // batched small programs (CudaBackend::AirBatch)
    typedef typename B::BatchScope BatchScope;
    static bool batch_small_components() { return B::batch_small_components(); }
    static void air_batch_flush() { B::air_batch_flush(); }
    // the input columns of a component WITHOUT real rows: every row is ExecutionBundle::default() (witness.cu), whatever the
    // component -- unpacked once per proof and shared by all such components (their trace programs run as one batch)
    static std::vector<Col> unpack_padding(const Words& accesses, size_t n_accesses, u32 log_size) {
        std::vector<Col> cols = Col::many(N_BUNDLE_INPUTS, (size_t)1 << log_size);
        std::vector<u32*> out;
        for (auto& c : cols) out.push_back(c.ptr());
        cm_check(cm31_unpack_bundles_slots(accesses.ptr(), 0, log_size, accesses.ptr(), n_accesses, out.data(), MAX_ACCESSES));
        return cols;
    }
    static void staging_release_point(int point) {
        if (point == prefetch_point()) cm_check(cm31_bg_release_throttled(0, prefetch_ctas()));
    }
    // The AoS -> SoA unpack of a component's bundles is issued by write_trace, which knows the component's trace program and
    // therefore how many of the 8 access slots it reads: the columns of the other slots are allocated (fixed column indices)
    // but never written — for store_fp_imm (2 accesses) 18 columns instead of 42.
    struct PendingUnpack {
        const u32 *rows = nullptr, *accesses = nullptr;
        size_t n_real = 0, n_accesses = 0;
        u32 log_size = 0;
        std::vector<u32*> out;
        bool active = false;
    };
    static PendingUnpack& pending_unpack() {
        static PendingUnpack p;
        return p;
    }
    static std::vector<Col> unpack_bundles(const Words& rows, size_t n_real, const Words& accesses, size_t n_accesses, u32 log_size) {
        std::vector<Col> cols = Col::many(N_BUNDLE_INPUTS, (size_t)1 << log_size);
        PendingUnpack& pu = pending_unpack();
        // (a still-active entry belongs to a proof that was aborted by an exception: its columns are gone, drop it)
        pu.out.clear();
        for (auto& c : cols) pu.out.push_back(c.ptr());
        pu.rows = rows.ptr();
        pu.accesses = accesses.ptr();
        pu.n_real = n_real;
        pu.n_accesses = n_accesses;
        pu.log_size = log_size;
        pu.active = true;
        return cols;
    }
    // access slots read by a trace program: the highest input column it loads
    static u32 access_slots_read(const AirProgram& prog) {
        u32 max_col = 0;
        for (uint64_t ins : prog.code)
            if ((ins & 0xff) == OP_LOAD) max_col = std::max<u32>(max_col, (u32)((ins >> 24) & 0xfffff));
        if (max_col < (u32)IN_ACC_BASE) return 0;
        return std::min<u32>(MAX_ACCESSES, (max_col - IN_ACC_BASE) / 4 + 1);
    }
    static std::vector<Col> unpack_rows(const Words& rows, size_t n_real, u32 n_fields, u32 log_size) {
        std::vector<Col> cols = Col::many(n_fields, (size_t)1 << log_size);
        std::vector<u32*> p;
        for (auto& c : cols) p.push_back(c.ptr());
        if (!Shard::get().skip()) cm_check(cm31_unpack_rows(rows.ptr(), n_real, n_fields, log_size, p.data()));
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
        PendingUnpack& pu = pending_unpack();
        if (pu.active) {  // an opcode component: its inputs are unpacked here, only as wide as the program reads
            if (inputs.empty() || pu.out.empty() || inputs[0].ptr() != pu.out[0]) throw std::logic_error("write_trace: pending unpack belongs to other inputs");
            pu.active = false;
            if (!Shard::get().skip())
                cm_check(cm31_unpack_bundles_slots(pu.rows, pu.n_real, pu.log_size, pu.accesses, pu.n_accesses, pu.out.data(), access_slots_read(prog)));
        }
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
    // a looked-up value outside its table raised the device error word instead of being counted (cm31_air_lookups)
    static void check_lookups() { CudaBackend::check_air_errors(); }
};

}  // namespace cm31

// FrameworkComponent: a component defined by a `FrameworkEval` (log_size + evaluate<E>) proven on
// a backend `B`.  Host-side mirror of
//   FrameworkEval / FrameworkComponent   external/stwo/crates/constraint_framework/src/component.rs:122-281
//   ComponentProver impl (domain eval)   component.rs:282-424
//   TraceLocationAllocator               component.rs:50-120
//   LogupTraceGenerator usage            crates/prover/src/components/opcodes/store_fp_fp.rs:321-479
// The AIR is captured once into a graph (air_expr.hpp); constraint evaluation, interaction-trace
// generation and lookup emission all run as bytecode programs on the backend.
#pragma once
#include <functional>
#include <memory>
#include <mutex>
#include <string>

#include "../air/cairo_components.hpp"
#include "air_expr.hpp"
#include "stark.hpp"

namespace cm31 {

struct RelationSet {
    std::vector<RelationElements> relations;  // indexed by relation id
    const RelationElements& get(int id) const { return relations.at(id); }
};

struct TreeSubspan {
    size_t tree_index, col_start, col_end;
};

struct TraceLocationAllocator {
    std::vector<size_t> next_tree_offsets;
    std::vector<std::string> preprocessed_columns;  // static allocation (ids of tree 0 columns)
    explicit TraceLocationAllocator(std::vector<std::string> preprocessed = {}) : preprocessed_columns(std::move(preprocessed)) {}
    std::vector<TreeSubspan> next_for_structure(const std::vector<size_t>& n_cols_per_tree) {
        if (n_cols_per_tree.size() > next_tree_offsets.size()) next_tree_offsets.resize(n_cols_per_tree.size(), 0);
        std::vector<TreeSubspan> out;
        for (size_t t = 0; t < n_cols_per_tree.size(); t++) {
            out.push_back(TreeSubspan{t, next_tree_offsets[t], next_tree_offsets[t] + n_cols_per_tree[t]});
            next_tree_offsets[t] += n_cols_per_tree[t];
        }
        return out;
    }
    size_t preprocessed_index(const std::string& id) const {
        for (size_t i = 0; i < preprocessed_columns.size(); i++)
            if (preprocessed_columns[i] == id) return i;
        throw std::logic_error("Preprocessed column " + id + " is missing from static allocation");
    }
};

// ------------------------------------------------------------------ the programs of one captured AIR
// (shared by the prover and by tools/gen_air_kernels.cpp, which emits each of them as an
// AOT-specialised CUDA kernel).  Column index convention of every program: preprocessed columns
// first (first-use order), then the trace columns, then the interaction columns.
inline AirProgram build_constraint_program(const ExprEvaluator& ev, bool emit_cuda = false) {
    std::vector<ProgramOutput> outs;
    size_t n_eval_params = ev.params.size();
    for (size_t k = 0; k < ev.constraints.size(); k++)
        outs.push_back(ProgramOutput{ProgramOutput::ConstraintSum, ev.constraints[k], (int)(n_eval_params + k)});
    size_t n_pre = ev.mask_offsets[0].size(), n_tr = ev.mask_offsets[1].size();
    return ProgramBuilder::compile(ev.g, outs, n_eval_params + ev.constraints.size(), [&](int interaction, int col) -> size_t {
        if (interaction == 0) return (size_t)col;
        if (interaction == 1) return n_pre + (size_t)col;
        return n_pre + n_tr + (size_t)col;
    }, emit_cuda);
}
// cumulative logup columns col_b = col_{b-1} + num_b/den_b  (logup.rs:123-320); extends the graph
inline AirProgram build_logup_program(ExprEvaluator& e, bool emit_cuda = false) {
    std::vector<ProgramOutput> outs;
    // 1/den_b for every batch of the row with ONE inversion (Montgomery's trick, the per-row
    // analogue of the reference's batch_inverse over a column, fields/mod.rs:69-99)
    size_t k = e.batch_fracs.size();
    std::vector<EFExpr> prefix(k), inv(k);
    for (size_t b = 0; b < k; b++) prefix[b] = b == 0 ? e.batch_fracs[0].den : e.ef_mul(prefix[b - 1], e.batch_fracs[b].den);
    EFExpr run = e.ef_inv(prefix[k - 1]);
    for (size_t b = k - 1; b >= 1; b--) {
        inv[b] = e.ef_mul(run, prefix[b - 1]);
        run = e.ef_mul(run, e.batch_fracs[b].den);
    }
    inv[0] = run;
    EFExpr cum = e.ef_zero();
    for (size_t b = 0; b < k; b++) {
        EFExpr term = e.ef_mul(e.batch_fracs[b].num, inv[b]);
        cum = b == 0 ? term : e.ef_add(cum, term);
        outs.push_back(ProgramOutput{ProgramOutput::StoreE, cum.id, (int)(4 * b)});
    }
    size_t n_pre = e.mask_offsets[0].size();
    return ProgramBuilder::compile(e.g, outs, e.params.size(), [&](int interaction, int col) -> size_t {
        if (interaction == 0) return (size_t)col;
        if (interaction == 1) return n_pre + (size_t)col;
        throw std::logic_error("logup program reads an interaction column");
    }, emit_cuda);
}
// multiplicity histogram of the tuples looked up in `relation` (empty program if none): the table
// row of a tuple is sum_i weights[i] * value_i (cairo_table_index_weights: the value itself for a
// range check, op*2^16 + in1*2^8 + in2 for the stacked bitwise table)
inline AirProgram build_lookup_program(const ExprEvaluator& ev, int relation, const std::vector<u32>& weights, bool emit_cuda = false) {
    Graph g = ev.g;  // the index expressions are appended to a private copy
    std::vector<ProgramOutput> outs;
    for (auto& u : ev.logup_uses) {
        if (u.relation != relation) continue;
        int idx = g.constf(0);
        for (size_t i = 0; i < u.values.size() && i < weights.size(); i++)
            if (weights[i] != 0) idx = g.addf(idx, g.mulf(g.constf(weights[i]), u.values[i]));
        outs.push_back(ProgramOutput{ProgramOutput::Hist, idx, 0});
    }
    if (outs.empty()) return AirProgram();
    return ProgramBuilder::compile(g, outs, ev.params.size(), [&](int interaction, int col) -> size_t {
        if (interaction == 1) return (size_t)col;
        throw std::logic_error("lookup emission reads a non-trace column");
    }, emit_cuda);
}

// Trace-fill program builder: `write_trace<T>` of a component captured into the AIR bytecode
// (replaces the per-opcode row loops, crates/prover/src/components/opcodes/*.rs write_trace).
struct TraceProgramBuilder {
    typedef FExpr F;
    ExprEvaluator ev;
    u32 n_real;
    std::vector<ProgramOutput> outs;
    explicit TraceProgramBuilder(u32 n) : n_real(n) {}
    F in(int i) { return ev.input(i); }
    F enabler() { return ev.row_lt(n_real); }
    F f_const(u32 v) { return ev.f_const(v); }
    F f_inv(F a) { return ev.f_inv(a); }
    F f_shr(F a, u32 k) { return ev.f_shr(a, k); }
    F f_and(F a, u32 m) { return ev.f_and(a, m); }
    F f_le(F a, F b) { return ev.f_le(a, b); }
    F f_divc(F a, u32 c) { return ev.f_divc(a, c); }
    F f_modc(F a, u32 c) { return ev.f_modc(a, c); }
    F f_u32_divrem(F n_lo, F n_hi, F d_lo, F d_hi, u32 part) { return ev.f_u32_divrem(n_lo, n_hi, d_lo, d_hi, part); }
    void out(int col, F v) { outs.push_back(ProgramOutput{ProgramOutput::StoreF, v.id, col}); }
    AirProgram compile(bool emit_cuda = false) {
        return ProgramBuilder::compile(ev.g, outs, 0, [](int interaction, int col) -> size_t {
            if (interaction != 3) throw std::logic_error("trace program reads a non-input column");
            return (size_t)col;
        }, emit_cuda);
    }
};

// One symbolic capture per AIR shape, shared by every component instance of every proof: the graph
// and its programs depend on the Eval type (and its `cache_tag()`, e.g. the relation of a range
// check or the width of wide_fibonacci), never on log_size or on drawn values.
struct CapturedAir {
    ExprEvaluator ev;
    AirProgram constraint_program, logup_program;
    std::map<int, AirProgram> lookup_programs;  // relation -> histogram program (built on demand)
};
template <class Eval>
auto air_cache_tag(const Eval& e, int) -> decltype(e.cache_tag()) { return e.cache_tag(); }
template <class Eval>
long air_cache_tag(const Eval&, long) { return 0; }

template <class Eval>
std::shared_ptr<CapturedAir> capture_air(const Eval& eval) {
    static std::map<long, std::shared_ptr<CapturedAir>> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    long tag = (long)air_cache_tag(eval, 0);
    auto it = cache.find(tag);
    if (it != cache.end()) return it->second;
    auto cap = std::make_shared<CapturedAir>();
    eval.evaluate(cap->ev);
    if (!cap->ev.logup_finalized) throw std::logic_error("LogupAtRow was not finalized");
    cap->constraint_program = build_constraint_program(cap->ev);
    if (!cap->ev.batch_fracs.empty()) cap->logup_program = build_logup_program(cap->ev);
    cache[tag] = cap;
    return cap;
}

template <class B, class Eval>
class FrameworkComponent : public ComponentProver<B> {
   public:
    typedef typename B::Col Col;
    Eval eval;
    std::shared_ptr<CapturedAir> captured;
    const ExprEvaluator& ev;  // captured AIR (shared, immutable after capture)
    const RelationSet* relations = nullptr;
    std::vector<TreeSubspan> trace_locations;
    std::vector<size_t> preprocessed_indices;
    QM31 claimed_sum = {0, 0, 0, 0};
    long pending_sum_slot = -1;  // index into the backend's pending claimed sums, -1 = none

    FrameworkComponent(Eval e, const RelationSet* rel) : eval(std::move(e)), captured(capture_air(eval)), ev(captured->ev), relations(rel) {}
    // FrameworkComponent::new (component.rs:139-180): trace locations in component creation order.
    void allocate(TraceLocationAllocator& alloc) {
        std::vector<size_t> n_cols(3, 0);
        n_cols[ORIGINAL_TRACE_IDX] = n_trace_columns();
        n_cols[INTERACTION_TRACE_IDX] = n_interaction_columns();
        trace_locations = alloc.next_for_structure(n_cols);
        preprocessed_indices.clear();
        for (auto& id : ev.preprocessed_ids) preprocessed_indices.push_back(alloc.preprocessed_index(id));
    }
    size_t n_trace_columns() const { return ev.mask_offsets[ORIGINAL_TRACE_IDX].size(); }
    size_t n_interaction_columns() const { return ev.mask_offsets[INTERACTION_TRACE_IDX].size(); }
    u32 log_size() const { return eval.log_size(); }

    // ---- Component
    size_t n_constraints() const override { return ev.n_constraints(); }
    u32 max_constraint_log_degree_bound() const override { return eval.max_constraint_log_degree_bound(); }
    std::vector<std::vector<u32>> trace_log_degree_bounds() const override {
        std::vector<std::vector<u32>> out(3);
        out[0].assign(preprocessed_indices.size(), log_size());
        out[1].assign(n_trace_columns(), log_size());
        out[2].assign(n_interaction_columns(), log_size());
        return out;
    }
    MaskPoints mask_points(SecurePoint point) const override {
        CirclePointM31 trace_step = CanonicCoset(log_size()).step();
        MaskPoints out(3);
        // offset 0 is the point itself; the few other offsets (-1 for the logup cumulative sums) are shifted once per
        // component, not once per column (this runs on the critical path between the composition root and the OODS launch)
        std::vector<std::pair<int, SecurePoint>> shifted;
        auto at = [&](int off) -> const SecurePoint& {
            if (off == 0) return point;
            for (auto& kv : shifted)
                if (kv.first == off) return kv.second;
            shifted.emplace_back(off, secure_point_add_m31(point, cp_mul_signed(trace_step, off)));
            return shifted.back().second;
        };
        shifted.reserve(8);
        for (int t = 1; t < 3; t++) {
            out[t].reserve(ev.mask_offsets[t].size());
            for (auto& offsets : ev.mask_offsets[t]) {
                out[t].emplace_back();
                std::vector<SecurePoint>& pts = out[t].back();
                pts.reserve(offsets.size());
                for (int off : offsets) pts.push_back(at(off));
            }
        }
        return out;
    }
    void fill_mask_points(SecurePoint point, MaskPoints& out, std::vector<size_t>& at) const override {
        CirclePointM31 trace_step = CanonicCoset(log_size()).step();
        std::vector<std::pair<int, SecurePoint>> shifted;
        shifted.reserve(8);
        auto at_off = [&](int off) -> const SecurePoint& {
            if (off == 0) return point;
            for (auto& kv : shifted)
                if (kv.first == off) return kv.second;
            shifted.emplace_back(off, secure_point_add_m31(point, cp_mul_signed(trace_step, off)));
            return shifted.back().second;
        };
        for (int t = 1; t < 3; t++)
            for (auto& offsets : ev.mask_offsets[t]) {
                std::vector<SecurePoint>& pts = out[(size_t)t][at[(size_t)t]++];
                for (size_t k = 0; k < offsets.size(); k++) pts[k] = at_off(offsets[k]);
            }
    }
    std::vector<size_t> preprocessed_column_indices() const override { return preprocessed_indices; }

    std::vector<QM31> eval_params() const {
        std::vector<QM31> p;
        for (auto& d : ev.params) {
            switch (d.kind) {
                case ExprEvaluator::ParamDesc::CumsumShift:
                    p.push_back(qm_mul_m31(claimed_sum, m31_inv((u32)(((u64)1 << log_size()) % P))));
                    break;
                case ExprEvaluator::ParamDesc::RelationZ: p.push_back(relations->get(d.relation).z); break;
                case ExprEvaluator::ParamDesc::RelationAlphaPow: {
                    const RelationElements& r = relations->get(d.relation);
                    if ((size_t)d.power >= r.alpha_powers.size()) throw std::logic_error("Not enough alpha powers to combine values");
                    p.push_back(r.alpha_powers[d.power]);
                    break;
                }
            }
        }
        return p;
    }

    // PointEvaluator path (component.rs:255-279, point.rs)
    void evaluate_constraint_quotients_at_point(SecurePoint point, const MaskValues& mask, PointEvaluationAccumulator& acc) const override {
        GraphPointEval pe(ev.g);
        pe.params = eval_params();
        for (size_t c = 0; c < ev.mask_offsets[0].size(); c++) pe.mask[ColumnRef{0, (int)c, 0}] = mask[0][preprocessed_indices[c]].at(0);
        for (int t = 1; t < 3; t++)
            for (size_t c = 0; c < ev.mask_offsets[t].size(); c++) {
                const std::vector<QM31>& vals = mask[t][trace_locations[t].col_start + c];
                const std::vector<int>& offs = ev.mask_offsets[t][c];
                if (vals.size() != offs.size()) throw std::logic_error("mask shape mismatch");
                for (size_t k = 0; k < offs.size(); k++) pe.mask[ColumnRef{t, (int)c, offs[k]}] = vals[k];
            }
        QM31 denom_inverse = qm_inv(coset_vanishing_qm31(CanonicCoset(log_size()).coset, point));
        for (int cid : ev.constraints) acc.accumulate(denom_inverse * pe.eval(cid));
    }

    // Domain path (component.rs:283-424) on the backend's AIR kernel.
    void evaluate_constraint_quotients_on_domain(const Trace<B>& trace, DomainEvaluationAccumulator<B>& accumulator) const override {
        if (n_constraints() == 0) return;
        u32 eval_log = max_constraint_log_degree_bound();
        u32 trace_log = log_size();
        std::vector<const Col*> cols;
        auto push = [&](size_t tree, size_t idx) {
            const CircleEvaluation<B>& e = (*trace.trees)[tree].evaluations.at(idx);
            if (e.log_size != eval_log) throw std::logic_error("constraint evaluation domain differs from the committed domain");
            cols.push_back(&e.values);
        };
        for (size_t c = 0; c < preprocessed_indices.size(); c++) push(0, preprocessed_indices[c]);
        for (int t = 1; t < 3; t++)
            for (size_t c = trace_locations[t].col_start; c < trace_locations[t].col_end; c++) push(t, c);
        // denominator inverses (component.rs:328-333)
        CircleDomain eval_domain = CanonicCoset(eval_log).circle_domain();
        Coset trace_coset = CanonicCoset(trace_log).coset;
        u32 log_expand = eval_log - trace_log;
        std::vector<u32> denom_inv((size_t)1 << log_expand);
        for (size_t i = 0; i < denom_inv.size(); i++) denom_inv[i] = m31_inv(coset_vanishing_m31(trace_coset, eval_domain.at(i)));
        {
            std::vector<u32> br(denom_inv.size());
            for (size_t i = 0; i < br.size(); i++) br[i] = denom_inv[bit_reverse((u32)i, log_expand)];
            denom_inv = br;
        }
        auto accum = accumulator.columns(eval_log, n_constraints());
        std::vector<QM31>& coeffs = accum.first;
        std::reverse(coeffs.begin(), coeffs.end());
        AirProgram prog = captured->constraint_program;
        std::vector<QM31> params = eval_params();
        params.insert(params.end(), coeffs.begin(), coeffs.end());
        fill_params(prog, params);
        B::constraint_eval(cols, trace_log, eval_log, prog, denom_inv, *accum.second);
    }

    // Interaction trace = cumulative logup columns (logup.rs:123-320), generated from the same AIR.
    // `trace_cols`: the component's trace columns on the trace domain; `preprocessed`: resolver of
    // preprocessed column ids to trace-domain columns.  Sets claimed_sum.
    std::vector<CircleEvaluation<B>> gen_interaction_trace(const std::vector<const Col*>& trace_cols,
                                                           const std::function<const Col*(const std::string&)>& preprocessed) {
        size_t n_batches = ev.batch_fracs.size();
        std::vector<CircleEvaluation<B>> out;
        if (n_batches == 0) return out;
        if (trace_cols.size() != n_trace_columns()) throw std::logic_error("gen_interaction_trace: wrong number of trace columns");
        std::vector<const Col*> in;
        for (auto& id : ev.preprocessed_ids) in.push_back(preprocessed(id));
        in.insert(in.end(), trace_cols.begin(), trace_cols.end());
        size_t n = (size_t)1 << log_size();
        out.resize(4 * n_batches);
        std::vector<Col> slab = B::uninit_many(4 * n_batches, n);
        std::vector<Col*> outp;
        for (size_t i = 0; i < out.size(); i++) {
            out[i].values = std::move(slab[i]);
            out[i].log_size = log_size();
            outp.push_back(&out[i].values);
        }
        AirProgram prog = captured->logup_program;
        fill_params(prog, eval_params());  // cumsum shift is unused by this program
        B::air_program(in, outp, log_size(), prog);
        std::array<Col*, 4> last = {outp[4 * n_batches - 4], outp[4 * n_batches - 3], outp[4 * n_batches - 2], outp[4 * n_batches - 1]};
        // stream-ordered: the sum is fetched later by the driver (Impl::collect_claimed_sums)
        pending_sum_slot = (long)B::logup_finalize_last_async(last, log_size());
        return out;
    }

    // Emits every value looked up in `relation` (first tuple element) into a multiplicity histogram
    // (crates/prover/src/components/opcodes/mod.rs:83-105 providers + range_check_macro.rs:72-84).
    void emit_lookups(int relation, const std::vector<const Col*>& trace_cols, Col& bins) const {
        auto it = captured->lookup_programs.find(relation);
        if (it == captured->lookup_programs.end())
            it = captured->lookup_programs.emplace(relation, build_lookup_program(ev, relation, cairo_table_index_weights(relation))).first;
        const AirProgram& prog = it->second;
        if (prog.code.empty()) return;
        B::air_lookups(trace_cols, bins, log_size(), prog);
    }

    const AirProgram& constraint_program() const { return captured->constraint_program; }

   private:
    static void fill_params(AirProgram& prog, const std::vector<QM31>& params) {
        if (params.size() != prog.param_slots.size()) throw std::logic_error("parameter count mismatch");
        for (size_t i = 0; i < params.size(); i++) {
            u32 s = prog.param_slots[i];
            prog.consts[s] = params[i].a;
            prog.consts[s + 1] = params[i].b;
            prog.consts[s + 2] = params[i].c;
            prog.consts[s + 3] = params[i].d;
        }
    }
};

}  // namespace cm31

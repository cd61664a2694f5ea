// ORACLE (test infrastructure, not product): direct (non-bytecode) evaluation of an AIR's
// `evaluate<E>` with scalar CPU evaluators, restating
//   CpuDomainEvaluator   external/stwo/crates/constraint_framework/src/cpu_domain.rs:17-113
//   PointEvaluator       constraint_framework/src/point.rs:12-70
//   ComponentProver      constraint_framework/src/component.rs:283-374 (CPU fallback loop)
//   LogupTraceGenerator  constraint_framework/src/logup.rs:123-320 (per-row fractions, finalize_last)
// The product captures the same `evaluate<E>` into bytecode; instantiating it here with M31/QM31
// value types gives an independent result to diff the CUDA path against.
#pragma once
#include <functional>
#include <map>
#include <string>

#include "host/air_expr.hpp"
#include "host/framework.hpp"
#include "oracle_backend.hpp"

namespace orc {

typedef std::vector<M31> OCol;

template <class Derived, class F_, class EF_>
struct EvalCommon : cm31::LogupMixin<Derived, F_, EF_> {
    const cm31::RelationSet* relations = nullptr;
    QM31 cumsum_shift_value;
    EF_ ef_zero() { return EF_(QM31::zero()); }
    EF_ ef_one() { return EF_(QM31::one()); }
    EF_ ef_const(cm31::QM31 v) { return EF_(to_orc(v)); }
    EF_ ef_add(EF_ a, EF_ b) { return a + b; }
    EF_ ef_sub(EF_ a, EF_ b) { return a - b; }
    EF_ ef_mul(EF_ a, EF_ b) { return a * b; }
    EF_ cumsum_shift() { return EF_(cumsum_shift_value); }
    F_ add_intermediate(F_ v) { return v; }
    EF_ add_extension_intermediate(EF_ v) { return v; }
    void add_to_relation(int relation, EF_ multiplicity, const std::vector<F_>& values) {
        Derived& self = static_cast<Derived&>(*this);
        const cm31::RelationElements& r = relations->get(relation);
        if (values.size() > r.alpha_powers.size()) throw std::logic_error("Not enough alpha powers to combine values");
        EF_ acc = ef_zero();
        for (size_t i = 0; i < values.size(); i++) acc = acc + self.mul_ef_f(EF_(to_orc(r.alpha_powers[i])), values[i]);
        EF_ den = acc - EF_(to_orc(r.z));
        self.on_relation_use(relation, values);
        this->write_logup_frac(multiplicity, den);
    }
};

// ---- domain evaluator: F = M31, EF = QM31, one LDE row
struct CpuDomainEvaluator : EvalCommon<CpuDomainEvaluator, M31, QM31> {
    typedef M31 F;
    typedef QM31 EF;
    const std::vector<std::vector<const OCol*>>* trace_eval;  // [interaction][col]
    std::vector<size_t> col_index;
    size_t row;
    const std::vector<QM31>* random_coeff_powers;
    QM31 row_res;
    size_t constraint_index = 0;
    u32 domain_log_size, eval_domain_log_size;

    F next_interaction_mask(int interaction, int offset) {
        size_t c = col_index[interaction]++;
        return at(interaction, c, offset);
    }
    F at(int interaction, size_t c, int offset) {
        const OCol& col = *(*trace_eval)[interaction][c];
        if (offset == 0) return col[row];
        size_t r = cm31::offset_bit_reversed_circle_domain_index(row, domain_log_size, eval_domain_log_size, offset);
        return col[r];
    }
    F next_trace_mask() { return next_interaction_mask(1, 0); }
    F get_preprocessed_column(const std::string&) { return next_interaction_mask(0, 0); }
    EF next_extension_interaction_mask_0(int interaction) {
        F v[4];
        for (int k = 0; k < 4; k++) v[k] = next_interaction_mask(interaction, 0);
        return QM31(v[0], v[1], v[2], v[3]);
    }
    void next_extension_interaction_mask_m1_0(int interaction, EF& prev, EF& cur) {
        F p[4], c[4];
        for (int k = 0; k < 4; k++) {
            size_t ci = col_index[interaction]++;
            p[k] = at(interaction, ci, -1);
            c[k] = at(interaction, ci, 0);
        }
        prev = QM31(p[0], p[1], p[2], p[3]);
        cur = QM31(c[0], c[1], c[2], c[3]);
    }
    void add_constraint(F c) { add_constraint_ef(QM31::from_m31(c)); }
    void add_constraint(EF c) { add_constraint_ef(c); }
    void add_constraint_ef(EF c) {
        row_res = row_res + (*random_coeff_powers)[constraint_index] * c;
        constraint_index++;
    }
    F f_const(u32 v) { return M31((u64)v); }
    F f_const_i(long long v) { return M31::from_i64(v); }
    EF ef(F v) { return QM31::from_m31(v); }
    EF mul_ef_f(EF a, F b) { return a * b; }
    void on_relation_use(int, const std::vector<F>&) {}
};

// ---- point evaluator: F = EF = QM31
struct PointEvaluator : EvalCommon<PointEvaluator, QM31, QM31> {
    typedef QM31 F;
    typedef QM31 EF;
    const std::vector<std::vector<const std::vector<cm31::QM31>*>>* mask;  // [interaction][col] -> samples
    std::vector<size_t> col_index;
    cm31::PointEvaluationAccumulator* acc;
    QM31 denom_inverse;

    const std::vector<cm31::QM31>& next_col(int interaction) { return *(*mask)[interaction][col_index[interaction]++]; }
    F next_interaction_mask(int interaction, int) {
        const auto& m = next_col(interaction);
        if (m.size() != 1) throw std::logic_error("mask size mismatch");
        return to_orc(m[0]);
    }
    F next_trace_mask() { return next_interaction_mask(1, 0); }
    F get_preprocessed_column(const std::string&) { return next_interaction_mask(0, 0); }
    EF next_extension_interaction_mask_0(int interaction) {
        QM31 v[4];
        for (int k = 0; k < 4; k++) v[k] = next_interaction_mask(interaction, 0);
        return QM31::from_partial_evals(v[0], v[1], v[2], v[3]);
    }
    void next_extension_interaction_mask_m1_0(int interaction, EF& prev, EF& cur) {
        QM31 p[4], c[4];
        for (int k = 0; k < 4; k++) {
            const auto& m = next_col(interaction);
            if (m.size() != 2) throw std::logic_error("mask size mismatch");
            p[k] = to_orc(m[0]);
            c[k] = to_orc(m[1]);
        }
        prev = QM31::from_partial_evals(p[0], p[1], p[2], p[3]);
        cur = QM31::from_partial_evals(c[0], c[1], c[2], c[3]);
    }
    void add_constraint(EF c) { add_constraint_ef(c); }
    void add_constraint_ef(EF c) { acc->accumulate(from_orc(denom_inverse * c)); }
    F f_const(u32 v) { return QM31::from_m31(M31((u64)v)); }
    F f_const_i(long long v) { return QM31::from_m31(M31::from_i64(v)); }
    EF ef(F v) { return v; }
    EF mul_ef_f(EF a, F b) { return a * b; }
    void on_relation_use(int, const std::vector<F>&) {}
};

// ---- per-trace-row logup evaluator: computes the cumulative logup values of one row
struct RowLogupEvaluator : EvalCommon<RowLogupEvaluator, M31, QM31> {
    typedef M31 F;
    typedef QM31 EF;
    const std::vector<const OCol*>* trace_cols;
    const std::vector<const OCol*>* preprocessed_cols;
    size_t n_trace = 0, n_pre = 0;
    size_t row;
    std::function<void(int, u32)> on_use;  // (relation, table row of the looked-up tuple) for histograms

    F next_trace_mask() { return (*(*trace_cols)[n_trace++])[row]; }
    F get_preprocessed_column(const std::string&) { return (*(*preprocessed_cols)[n_pre++])[row]; }
    F next_interaction_mask(int interaction, int) {
        if (interaction == 1) return next_trace_mask();
        return M31();  // interaction-trace masks are outputs here
    }
    EF next_extension_interaction_mask_0(int) { return QM31::zero(); }
    void next_extension_interaction_mask_m1_0(int, EF& prev, EF& cur) {
        prev = QM31::zero();
        cur = QM31::zero();
    }
    void add_constraint(F) {}
    void add_constraint(EF) {}
    void add_constraint_ef(EF) {}
    F f_const(u32 v) { return M31((u64)v); }
    F f_const_i(long long v) { return M31::from_i64(v); }
    EF ef(F v) { return QM31::from_m31(v); }
    EF mul_ef_f(EF a, F b) { return a * b; }
    void on_relation_use(int relation, const std::vector<F>& values) {
        if (!on_use) return;
        std::vector<u32> w = cm31::cairo_table_index_weights(relation);
        u64 idx = 0;
        for (size_t i = 0; i < values.size() && i < w.size(); i++) idx += (u64)w[i] * values[i].v;
        on_use(relation, (u32)idx);
    }
};

// ---- component on the oracle backend
template <class Eval>
class OracleComponent : public cm31::ComponentProver<OracleBackend> {
   public:
    Eval eval;
    const cm31::RelationSet* relations;
    std::vector<cm31::TreeSubspan> trace_locations;
    std::vector<size_t> preprocessed_indices;
    std::vector<std::vector<std::vector<int>>> mask_offsets;
    std::vector<std::string> preprocessed_ids;
    size_t n_constraints_ = 0;
    cm31::QM31 claimed_sum = {0, 0, 0, 0};

    OracleComponent(Eval e, const cm31::RelationSet* rel) : eval(std::move(e)), relations(rel) {
        // InfoEvaluator (constraint_framework/src/info.rs): structure only — taken from a symbolic run
        cm31::ExprEvaluator info;
        eval.evaluate(info);
        mask_offsets = info.mask_offsets;
        preprocessed_ids = info.preprocessed_ids;
        n_constraints_ = info.n_constraints();
    }
    void allocate(cm31::TraceLocationAllocator& alloc) {
        std::vector<size_t> n_cols = {0, mask_offsets[1].size(), mask_offsets[2].size()};
        trace_locations = alloc.next_for_structure(n_cols);
        preprocessed_indices.clear();
        for (auto& id : preprocessed_ids) preprocessed_indices.push_back(alloc.preprocessed_index(id));
    }
    u32 log_size() const { return eval.log_size(); }
    size_t n_trace_columns() const { return mask_offsets[1].size(); }
    size_t n_interaction_columns() const { return mask_offsets[2].size(); }
    size_t n_constraints() const override { return n_constraints_; }
    u32 max_constraint_log_degree_bound() const override { return eval.max_constraint_log_degree_bound(); }
    std::vector<std::vector<u32>> trace_log_degree_bounds() const override {
        std::vector<std::vector<u32>> out(3);
        out[0].assign(preprocessed_indices.size(), log_size());
        out[1].assign(n_trace_columns(), log_size());
        out[2].assign(n_interaction_columns(), log_size());
        return out;
    }
    cm31::MaskPoints mask_points(cm31::SecurePoint point) const override {
        cm31::CirclePointM31 trace_step = cm31::CanonicCoset(log_size()).step();
        cm31::MaskPoints out(3);
        for (int t = 1; t < 3; t++)
            for (auto& offsets : mask_offsets[t]) {
                std::vector<cm31::SecurePoint> pts;
                for (int off : offsets) pts.push_back(cm31::secure_point_add_m31(point, cm31::cp_mul_signed(trace_step, off)));
                out[t].push_back(pts);
            }
        return out;
    }
    std::vector<size_t> preprocessed_column_indices() const override { return preprocessed_indices; }
    QM31 cumsum_shift() const { return to_orc(claimed_sum) * M31((u64)1 << log_size()).inverse(); }

    void evaluate_constraint_quotients_at_point(cm31::SecurePoint point, const cm31::MaskValues& mask,
                                                cm31::PointEvaluationAccumulator& acc) const override {
        std::vector<std::vector<const std::vector<cm31::QM31>*>> m(3);
        for (size_t idx : preprocessed_indices) m[0].push_back(&mask[0][idx]);
        for (int t = 1; t < 3; t++)
            for (size_t c = trace_locations[t].col_start; c < trace_locations[t].col_end; c++) m[t].push_back(&mask[t][c]);
        PointEvaluator pe;
        pe.relations = relations;
        pe.cumsum_shift_value = cumsum_shift();
        pe.mask = &m;
        pe.col_index.assign(3, 0);
        pe.acc = &acc;
        pe.denom_inverse = to_orc(cm31::qm_inv(cm31::coset_vanishing_qm31(cm31::CanonicCoset(log_size()).coset, point)));
        eval.evaluate(pe);
    }

    void evaluate_constraint_quotients_on_domain(const cm31::Trace<OracleBackend>& trace,
                                                 cm31::DomainEvaluationAccumulator<OracleBackend>& accumulator) const override {
        if (n_constraints() == 0) return;
        u32 eval_log = max_constraint_log_degree_bound(), trace_log = log_size();
        std::vector<std::vector<const OCol*>> cols(3);
        for (size_t idx : preprocessed_indices) cols[0].push_back(&(*trace.trees)[0].evaluations.at(idx).values);
        for (int t = 1; t < 3; t++)
            for (size_t c = trace_locations[t].col_start; c < trace_locations[t].col_end; c++) cols[t].push_back(&(*trace.trees)[t].evaluations.at(c).values);
        ODomain eval_domain = ODomain::canonic(eval_log);
        u32 log_expand = eval_log - trace_log;
        std::vector<M31> denom_inv((size_t)1 << log_expand);
        for (size_t i = 0; i < denom_inv.size(); i++) {
            Point p = eval_domain.at(i);
            cm31::CirclePointM31 cp = {p.x.v, p.y.v};
            denom_inv[i] = M31((u64)cm31::coset_vanishing_m31(cm31::CanonicCoset(trace_log).coset, cp)).inverse();
        }
        {
            std::vector<M31> br(denom_inv.size());
            for (size_t i = 0; i < br.size(); i++) br[i] = denom_inv[bitrev((u32)i, log_expand)];
            denom_inv = br;
        }
        auto accum = accumulator.columns(eval_log, n_constraints());
        std::vector<QM31> powers;
        for (auto it = accum.first.rbegin(); it != accum.first.rend(); ++it) powers.push_back(to_orc(*it));
        std::array<OCol, 4>& col = *accum.second;
        size_t n_rows = (size_t)1 << eval_log;
        QM31 shift = cumsum_shift();
#pragma omp parallel for schedule(static)
        for (size_t row = 0; row < n_rows; row++) {
            CpuDomainEvaluator de;
            de.relations = relations;
            de.cumsum_shift_value = shift;
            de.trace_eval = &cols;
            de.col_index.assign(3, 0);
            de.row = row;
            de.random_coeff_powers = &powers;
            de.domain_log_size = trace_log;
            de.eval_domain_log_size = eval_log;
            eval.evaluate(de);
            QM31 v = QM31(col[0][row], col[1][row], col[2][row], col[3][row]) + de.row_res * denom_inv[row >> trace_log];
            col[0][row] = v.x.a;
            col[1][row] = v.x.b;
            col[2][row] = v.y.a;
            col[3][row] = v.y.b;
        }
    }

    // LogupTraceGenerator restated per row + finalize_last (logup.rs:211-251)
    std::vector<cm31::CircleEvaluation<OracleBackend>> gen_interaction_trace(const std::vector<const OCol*>& trace_cols,
                                                                            const std::function<const OCol*(const std::string&)>& preprocessed,
                                                                            const std::function<void(int, u32)>& on_use = nullptr) {
        size_t n = (size_t)1 << log_size();
        size_t n_batches = n_interaction_columns() / 4;
        std::vector<const OCol*> pre;
        for (auto& id : preprocessed_ids) pre.push_back(preprocessed(id));
        std::vector<cm31::CircleEvaluation<OracleBackend>> out(4 * n_batches);
        for (auto& c : out) {
            c.values.assign(n, M31());
            c.log_size = log_size();
        }
        bool mismatch = false;
#pragma omp parallel for schedule(static) if (!on_use)
        for (size_t row = 0; row < n; row++) {
            RowLogupEvaluator re;
            re.relations = relations;
            re.cumsum_shift_value = QM31::zero();
            re.trace_cols = &trace_cols;
            re.preprocessed_cols = &pre;
            re.row = row;
            re.on_use = on_use;
            eval.evaluate(re);
            if (re.batch_fracs.size() != n_batches) {
                mismatch = true;
                continue;
            }
            QM31 cum = QM31::zero();
            for (size_t b = 0; b < n_batches; b++) {
                cum = cum + re.batch_fracs[b].num * re.batch_fracs[b].den.inverse();
                out[4 * b + 0].values[row] = cum.x.a;
                out[4 * b + 1].values[row] = cum.x.b;
                out[4 * b + 2].values[row] = cum.y.a;
                out[4 * b + 3].values[row] = cum.y.b;
            }
        }
        if (mismatch) throw std::logic_error("logup batch count mismatch");
        if (n_batches == 0) return out;
        // finalize_last
        QM31 claimed = QM31::zero();
        size_t l0 = 4 * (n_batches - 1);
        for (size_t row = 0; row < n; row++)
            claimed = claimed + QM31(out[l0].values[row], out[l0 + 1].values[row], out[l0 + 2].values[row], out[l0 + 3].values[row]);
        claimed_sum = from_orc(claimed);
        QM31 shift = claimed * M31((u64)n).inverse();
        M31 sh[4] = {shift.x.a, shift.x.b, shift.y.a, shift.y.b};
        for (int k = 0; k < 4; k++) {
            for (size_t row = 0; row < n; row++) out[l0 + k].values[row] = out[l0 + k].values[row] - sh[k];
            out[l0 + k].values = inclusive_prefix_sum(out[l0 + k].values);
        }
        return out;
    }
};

}  // namespace orc

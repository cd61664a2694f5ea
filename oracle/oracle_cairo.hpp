// ORACLE (test infrastructure, not product): scalar CPU implementation choices for the cairo-m
// protocol driver (cairo-m_b200/csrc/cairo/prover.hpp) + the cairo-m verifier.
//   Pack::pack / get_access_field   crates/prover/src/utils/execution_bundle.rs:31-75, utils/data_accesses.rs:10-28
//   write_trace row loops           crates/prover/src/components/opcodes/*.rs (direct evaluation, F = M31)
//   multiplicity histograms         crates/prover/src/preprocessed/range_check/range_check_macro.rs:72-84
//   verify_cairo_m                  crates/prover/src/verifier.rs:17-95
//   initial_logup_sum               crates/prover/src/public_data.rs:287-399
#pragma once
#include <chrono>

#include "cairo/prover.hpp"
#include "oracle_air.hpp"
#include "verifier.hpp"

namespace orc {

// direct evaluation of a component's write_trace<T> for one row
struct TraceRowEvaluator {
    typedef M31 F;
    const std::vector<OCol>* inputs;
    std::vector<OCol>* outputs;
    size_t row;
    u32 n_real;
    F in(int i) { return (*inputs)[i][row]; }
    F enabler() { return M31((u64)(row < n_real ? 1 : 0)); }
    F f_const(u32 v) { return M31((u64)v); }
    F f_inv(F a) { return a.v == 0 ? M31() : a.inverse(); }
    F f_shr(F a, u32 k) { return M31((u64)(a.v >> k)); }
    F f_and(F a, u32 m) { return M31((u64)(a.v & m)); }
    F f_le(F a, F b) { return M31((u64)(a.v <= b.v ? 1 : 0)); }
    F f_divc(F a, u32 c) { return M31((u64)(a.v / c)); }
    F f_modc(F a, u32 c) { return M31((u64)(a.v % c)); }
    // u32_store_div_fp_fp.rs:360-400: euclidean division on the limb pairs, (0, 0) for a zero divisor
    F f_u32_divrem(F n_lo, F n_hi, F d_lo, F d_hi, u32 part) {
        u32 n = n_lo.v | (n_hi.v << 16), d = d_lo.v | (d_hi.v << 16);
        u32 q = d == 0 ? 0 : n / d, r = d == 0 ? 0 : n % d;
        u32 x = part < 2 ? q : r;
        return M31((u64)((part & 1) ? x >> 16 : x & 0xffff));
    }
    void out(int col, F v) { (*outputs)[col][row] = v; }
};

struct OracleAirImpl {
    typedef OracleBackend B;
    typedef OCol Col;
    template <class Eval>
    using Component = OracleComponent<Eval>;
    typedef std::vector<u32> Words;
    static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    static Col iota(size_t n) {
        Col c(n);
        for (size_t i = 0; i < n; i++) c[i] = M31((u64)i);
        return c;
    }
    static Col clone(const Col& c) { return c; }
    static Col bitwise_table_col(int k) {
        Col c((size_t)1 << cm31::BITWISE_STACKED_LOG_SIZE);
        for (size_t i = 0; i < c.size(); i++) c[i] = M31((u64)cm31::bitwise_table_value(k, (u32)i));
        return c;
    }
    static Words upload_words(const u32* src, size_t n_words) { return Words(src, src + n_words); }
    static Words alloc_words(size_t n_words) { return Words(n_words); }
    static void staging_fence() {}
    static void staging_begin() {}
    static u32 staging_mark() { return 0; }
    static void staging_wait(u32) {}
    static void staging_release_point(int) {}
    static void idle_gate_open() {}
    struct BatchScope {
        explicit BatchScope(bool) {}
    };
    static bool batch_small_components() { return false; }
    static void air_batch_flush() {}
    static std::vector<Col> unpack_padding(const Words&, size_t, u32) { throw std::logic_error("the oracle does not batch"); }
    static void copy_words(Words& dst, size_t at, const u32* src, size_t n_words) { std::copy(src, src + n_words, dst.begin() + at); }
    static std::vector<Col> unpack_bundles(const Words& row_words, size_t n_real, const Words& access_words, size_t n_accesses, u32 log_size) {
        const cm31::Bundle* rows = (const cm31::Bundle*)row_words.data();
        const cm31::DataAccess* log = (const cm31::DataAccess*)access_words.data();
        size_t n = (size_t)1 << log_size;
        std::vector<Col> cols(cm31::N_BUNDLE_INPUTS, Col(n));
        for (size_t r = 0; r < n; r++) {
            cm31::Bundle b;
            if (r < n_real) b = rows[r];
            else {
                memset(&b, 0, sizeof(b));
                b.inst[0] = cm31::OP_RET;  // ExecutionBundle::default()
            }
            u32 head[10] = {b.pc, b.fp, b.clock, b.inst_prev_clock, b.inst[0], b.inst[1], b.inst[2], b.inst[3], b.inst[4], b.inst[5]};
            for (int k = 0; k < 10; k++) cols[k][r] = M31((u64)head[k]);
            for (int k = 0; k < cm31::MAX_ACCESSES; k++) {
                cm31::DataAccess a{0, 0, 0, 0};
                if ((u32)k < b.span_len && (size_t)b.span_start + k < n_accesses) a = log[b.span_start + k];
                cols[cm31::in_acc(k, cm31::ACC_ADDRESS)][r] = M31((u64)a.address);
                cols[cm31::in_acc(k, cm31::ACC_PREV_CLOCK)][r] = M31((u64)a.prev_clock);
                cols[cm31::in_acc(k, cm31::ACC_PREV_VALUE)][r] = M31((u64)a.prev_value);
                cols[cm31::in_acc(k, cm31::ACC_VALUE)][r] = M31((u64)a.value);
            }
        }
        return cols;
    }
    static std::vector<Col> unpack_rows(const Words& rows, size_t n_real, u32 n_fields, u32 log_size) {
        size_t n = (size_t)1 << log_size;
        std::vector<Col> cols(n_fields, Col(n));
        for (size_t r = 0; r < n_real; r++)
            for (u32 f = 0; f < n_fields; f++) cols[f][r] = M31((u64)rows[r * n_fields + f]);
        return cols;
    }
    template <class Eval>
    static std::vector<cm31::CircleEvaluation<B>> write_trace(const Eval& eval, const std::vector<Col>& inputs, u32 n_real) {
        size_t n = (size_t)1 << eval.log_size();
        std::vector<OCol> outs(Eval::N_TRACE_COLUMNS, OCol(n));
#pragma omp parallel for schedule(static)
        for (size_t row = 0; row < n; row++) {
            TraceRowEvaluator t{&inputs, &outs, row, n_real};
            eval.write_trace(t);
        }
        std::vector<cm31::CircleEvaluation<B>> out;
        for (auto& c : outs) out.push_back(cm31::CircleEvaluation<B>{std::move(c), eval.log_size()});
        return out;
    }
    template <class Components>
    static void collect_claimed_sums(Components&) {}  // the oracle computes them eagerly
    template <class Comp>
    static void emit_lookups(Comp& comp, int relation, const std::vector<const Col*>& trace_cols, Col& bins) {
        size_t n = (size_t)1 << comp.log_size();
        std::vector<const OCol*> pre;
        std::vector<u32> counts(bins.size(), 0);
        bool bad = false;
#pragma omp parallel for schedule(static)
        for (size_t row = 0; row < n; row++) {
            RowLogupEvaluator re;
            re.relations = comp.relations;
            re.cumsum_shift_value = QM31::zero();
            re.trace_cols = &trace_cols;
            re.preprocessed_cols = &pre;
            re.row = row;
            re.on_use = [&](int rel, u32 idx) {
                if (rel != relation) return;
                if (idx >= counts.size()) {  // an out-of-table tuple is a witness bug
                    bad = true;
                    return;
                }
#pragma omp atomic
                counts[idx]++;
            };
            comp.eval.evaluate(re);
        }
        if (bad) throw std::runtime_error("lookup outside its table");
        for (size_t i = 0; i < bins.size(); i++) bins[i] = bins[i] + M31((u64)counts[i]);
    }
    static void check_lookups() {}  // emit_lookups throws on the spot
};

inline QM31 logup_residual(const cm31::CairoProof& proof, const cm31::ProverInput& input, const cm31::PcsConfig* verifier_config = nullptr);

// verify_cairo_m (verifier.rs:17-95), closing logup-sum check included (verifier.rs:84-92: InvalidLogupSum).
inline void verify_cairo_m(const cm31::CairoProof& proof, cm31::PcsConfig pcs_config) {
    OChannel channel;
    channel.mix_u64(pcs_config.pow_bits);
    channel.mix_u64(pcs_config.fri_config.log_blowup_factor);
    channel.mix_u64(pcs_config.fri_config.n_queries);
    channel.mix_u64(pcs_config.fri_config.log_last_layer_degree_bound);
    {  // public_data.mix_into
        const cm31::PublicData& pd = proof.public_data;
        u32 head[7] = {pd.initial_registers.pc, pd.initial_registers.fp, pd.final_registers.pc, pd.final_registers.fp, pd.clock, pd.initial_root, pd.final_root};
        channel.mix_u32s(head, 7);
        const cm31::PublicRanges& r = proof.public_ranges;
        u32 lens[3] = {r.program_end - r.program_start, r.input_end - r.input_start, r.output_end - r.output_start};
        channel.mix_u32s(lens, 3);
        for (const std::vector<cm31::PublicEntry>* v : {&pd.program, &pd.input, &pd.output}) {
            std::vector<u32> w;
            for (const cm31::PublicEntry& e : *v) {
                w.push_back(e.addr);
                for (int k = 0; k < 4; k++) w.push_back(e.value[k]);
                w.push_back(e.clock);
            }
            channel.mix_u32s(w.data(), w.size());
        }
    }
    if (proof.stark_proof.commitments.size() != 4) throw VerificationError("expected 4 commitments");
    CommitmentSchemeVerifier cs(pcs_config);
    cs.commit(proof.stark_proof.commitments[0], cm31::cairo_preprocessed_log_sizes(), channel);
    std::vector<u32> log_sizes;
    for (auto& kv : proof.claim.log_sizes) {
        log_sizes.push_back(kv.second);
        channel.mix_u64(kv.second);
    }
    // components with dummy relations first: only their shapes are needed to size tree 1
    cm31::RelationSet dummy;
    for (int r = 0; r < cm31::N_CAIRO_RELATIONS; r++) dummy.relations.push_back(cm31::RelationElements::dummy(cm31::cairo_relation_size(r)));
    {
        cm31::CairoComponents<OracleAirImpl> shape(log_sizes, &dummy);
        std::vector<u32> sizes;
        shape.for_each([&](auto& c) {
            for (size_t k = 0; k < c.n_trace_columns(); k++) sizes.push_back(c.log_size());
        });
        cs.commit(proof.stark_proof.commitments[1], sizes, channel);
    }
    channel.mix_u64(proof.interaction_pow);
    if (channel.trailing_zeros() < cm31::INTERACTION_POW_BITS) throw VerificationError("Proof of work verification failed.");
    cm31::RelationSet relations;  // Relations::draw with the oracle channel
    for (int r = 0; r < cm31::N_CAIRO_RELATIONS; r++) {
        cm31::RelationElements re;
        std::vector<QM31> za = channel.draw_secure_felts(2);
        re.z = from_orc(za[0]);
        re.alpha = from_orc(za[1]);
        QM31 cur = QM31::one();
        for (size_t i = 0; i < cm31::cairo_relation_size(r); i++) {
            re.alpha_powers.push_back(from_orc(cur));
            cur = cur * za[1];
        }
        relations.relations.push_back(re);
    }
    if (proof.interaction_claim.claimed_sums.size() != log_sizes.size()) throw VerificationError("claimed sum count mismatch");
    for (auto& s : proof.interaction_claim.claimed_sums) channel.mix_felts({to_orc(s)});
    cm31::CairoComponents<OracleAirImpl> components(log_sizes, &relations);
    components.set_claimed_sums(proof.interaction_claim.claimed_sums);
    {
        std::vector<u32> sizes;
        components.for_each([&](auto& c) {
            for (size_t k = 0; k < c.n_interaction_columns(); k++) sizes.push_back(c.log_size());
        });
        cs.commit(proof.stark_proof.commitments[2], sizes, channel);
    }
    cm31::TraceLocationAllocator alloc(cm31::cairo_preprocessed_ids());
    components.allocate(alloc);
    verify(components.provers(), channel, cs, proof.stark_proof);
    // the closing logup check replays the transcript: it must run under the VERIFIER's parameters, never under the ones the
    // (attacker-supplied) proof carries -- a proof whose embedded config differs from the expected one is rejected outright
    {
        const cm31::PcsConfig& pc = proof.stark_proof.config;
        if (pc.pow_bits != pcs_config.pow_bits || pc.fri_config.log_blowup_factor != pcs_config.fri_config.log_blowup_factor ||
            pc.fri_config.n_queries != pcs_config.fri_config.n_queries ||
            pc.fri_config.log_last_layer_degree_bound != pcs_config.fri_config.log_last_layer_degree_bound)
            throw VerificationError("proof was made under a different PcsConfig");
    }
    if (!(logup_residual(proof, cm31::ProverInput(), &pcs_config) == QM31::zero())) throw VerificationError("InvalidLogupSum");
}

// Logup balance (InteractionClaim::claimed_sum, components/mod.rs:288-302 + public_data.rs:287-399):
//   Σ claimed sums + public-data sum == 0.
// Every relation must balance exactly (the Merkle / Poseidon2 relations through the merkle and poseidon2 components,
// under the Poseidon2 tables of csrc/cairo/poseidon2_constants.hpp, pinned by the reference KAT).
inline QM31 logup_residual(const cm31::CairoProof& proof, const cm31::ProverInput& input, const cm31::PcsConfig* verifier_config) {
    // replay the transcript up to Relations::draw (under the verifier's config when one is given)
    OChannel channel;
    cm31::PcsConfig cfg = verifier_config ? *verifier_config : proof.stark_proof.config;
    channel.mix_u64(cfg.pow_bits);
    channel.mix_u64(cfg.fri_config.log_blowup_factor);
    channel.mix_u64(cfg.fri_config.n_queries);
    channel.mix_u64(cfg.fri_config.log_last_layer_degree_bound);
    const cm31::PublicData& pd = proof.public_data;
    {
        u32 head[7] = {pd.initial_registers.pc, pd.initial_registers.fp, pd.final_registers.pc, pd.final_registers.fp, pd.clock, pd.initial_root, pd.final_root};
        channel.mix_u32s(head, 7);
        const cm31::PublicRanges& r = proof.public_ranges;
        u32 lens[3] = {r.program_end - r.program_start, r.input_end - r.input_start, r.output_end - r.output_start};
        channel.mix_u32s(lens, 3);
        for (const std::vector<cm31::PublicEntry>* v : {&pd.program, &pd.input, &pd.output}) {
            std::vector<u32> w;
            for (const cm31::PublicEntry& e : *v) {
                w.push_back(e.addr);
                for (int k = 0; k < 4; k++) w.push_back(e.value[k]);
                w.push_back(e.clock);
            }
            channel.mix_u32s(w.data(), w.size());
        }
    }
    channel.mix_root(to_hash(proof.stark_proof.commitments[0]));
    for (auto& kv : proof.claim.log_sizes) channel.mix_u64(kv.second);
    channel.mix_root(to_hash(proof.stark_proof.commitments[1]));
    channel.mix_u64(proof.interaction_pow);
    struct Rel {
        QM31 z;
        std::vector<QM31> pw;
        QM31 combine(const std::vector<M31>& v) const {
            QM31 acc = QM31::zero();
            for (size_t i = 0; i < v.size(); i++) acc = acc + pw[i] * v[i];
            return acc - z;
        }
    };
    std::vector<Rel> rel;
    for (int r = 0; r < cm31::N_CAIRO_RELATIONS; r++) {
        std::vector<QM31> za = channel.draw_secure_felts(2);
        Rel x;
        x.z = za[0];
        QM31 cur = QM31::one();
        for (size_t i = 0; i < cm31::cairo_relation_size(r); i++) {
            x.pw.push_back(cur);
            cur = cur * za[1];
        }
        rel.push_back(x);
    }
    auto m = [](u32 v) { return M31((u64)v); };
    QM31 sum = QM31::zero();
    for (auto& s : proof.interaction_claim.claimed_sums) sum = sum + to_orc(s);
    // public data: registers
    sum = sum + rel[cm31::REL_REGISTERS].combine({m(pd.initial_registers.pc), m(pd.initial_registers.fp), M31(1)}).inverse();
    sum = sum - rel[cm31::REL_REGISTERS].combine({m(pd.final_registers.pc), m(pd.final_registers.fp), m(pd.clock) + M31(1)}).inverse();
    // public memory entries (program / input emitted, output consumed)
    auto add_public = [&](const std::vector<cm31::PublicEntry>& es, bool emit) {
        for (const cm31::PublicEntry& e : es) {
            QM31 inv = rel[cm31::REL_MEMORY].combine({m(e.addr), m(e.clock), m(e.value[0]), m(e.value[1]), m(e.value[2]), m(e.value[3])}).inverse();
            sum = emit ? sum + inv : sum - inv;
        }
    };
    add_public(pd.program, true);
    add_public(pd.input, true);
    add_public(pd.output, false);
    // public_data.rs:307-321: the two roots are consumed (index 0, depth 0, value = root, tree = root)
    sum = sum + rel[cm31::REL_MERKLE].combine({M31(0), M31(0), m(pd.initial_root), m(pd.initial_root)}).inverse();
    sum = sum + rel[cm31::REL_MERKLE].combine({M31(0), M31(0), m(pd.final_root), m(pd.final_root)}).inverse();
    // public_data.rs:323-382: every public cell also takes its four leaves out of the tree of its side
    auto public_leaves = [&](const std::vector<cm31::PublicEntry>& es, u32 root) {
        for (const cm31::PublicEntry& e : es)
            for (u32 k = 0; k < 4; k++)
                sum = sum - rel[cm31::REL_MERKLE].combine({m(e.addr) * M31(4) + M31((u64)k), M31((u64)cm31::TREE_HEIGHT), m(e.value[k]), m(root)}).inverse();
    };
    public_leaves(pd.program, pd.initial_root);
    public_leaves(pd.input, pd.initial_root);
    public_leaves(pd.output, pd.final_root);
    (void)input;
    return sum;
}

}  // namespace orc

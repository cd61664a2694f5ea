// ORACLE (test infrastructure, not product): STARK verifier restated from the reference, used to
// check that proofs produced by the CUDA path (and by the oracle prover) verify.
//   verify                    external/stwo/crates/prover/src/core/prover/mod.rs:87-143
//   CommitmentSchemeVerifier  core/pcs/verifier.rs:18-125
//   fri_answers               core/pcs/quotients.rs:104-166
//   FriVerifier               core/fri.rs:370-875 (commit, decommit, layer verifiers, SparseEvaluation)
//   MerkleVerifier            core/vcs/verifier.rs:12-160
#pragma once
#include <map>
#include <set>
#include <string>

#include "oracle_backend.hpp"

namespace orc {

struct VerificationError : std::runtime_error {
    explicit VerificationError(const std::string& m) : std::runtime_error(m) {}
};

inline Hash to_hash(const cm31::Hash32& h) {
    Hash o;
    memcpy(o.b, h.b, 32);
    return o;
}

struct MerkleVerifier {
    Hash root;
    std::vector<u32> column_log_sizes;
    std::map<u32, size_t> n_columns_per_log_size;
    MerkleVerifier(Hash r, std::vector<u32> sizes) : root(r), column_log_sizes(std::move(sizes)) {
        for (u32 s : column_log_sizes) n_columns_per_log_size[s]++;
    }
    void verify(const std::map<u32, std::vector<size_t>>& queries_per_log_size, const std::vector<u32>& queried_values_in,
                const cm31::MerkleDecommitment& decommitment) const {
        if (column_log_sizes.empty()) return;
        u32 max_log_size = *std::max_element(column_log_sizes.begin(), column_log_sizes.end());
        size_t qv = 0, hw = 0, cw = 0;
        std::vector<std::pair<size_t, Hash>> last_layer_hashes;
        bool have_last = false;
        for (int layer_log_size = (int)max_log_size; layer_log_size >= 0; layer_log_size--) {
            auto nit = n_columns_per_log_size.find((u32)layer_log_size);
            size_t n_columns_in_layer = nit == n_columns_per_log_size.end() ? 0 : nit->second;
            std::vector<std::pair<size_t, Hash>> layer_total_queries;
            static const std::vector<size_t> empty;
            auto qit = queries_per_log_size.find((u32)layer_log_size);
            const std::vector<size_t>& layer_column_queries = qit == queries_per_log_size.end() ? empty : qit->second;
            size_t pi = 0, hi = 0, ci = 0;
            while (pi < last_layer_hashes.size() || ci < layer_column_queries.size()) {
                size_t node_index;
                bool has_p = pi < last_layer_hashes.size(), has_c = ci < layer_column_queries.size();
                if (has_p && has_c) node_index = std::min(last_layer_hashes[pi].first / 2, layer_column_queries[ci]);
                else if (has_p) node_index = last_layer_hashes[pi].first / 2;
                else node_index = layer_column_queries[ci];
                while (pi < last_layer_hashes.size() && last_layer_hashes[pi].first / 2 == node_index) pi++;
                Hash left, right;
                if (have_last) {
                    if (hi < last_layer_hashes.size() && last_layer_hashes[hi].first == 2 * node_index) left = last_layer_hashes[hi++].second;
                    else {
                        if (hw >= decommitment.hash_witness.size()) throw VerificationError("Witness is too short");
                        left = to_hash(decommitment.hash_witness[hw++]);
                    }
                    if (hi < last_layer_hashes.size() && last_layer_hashes[hi].first == 2 * node_index + 1) right = last_layer_hashes[hi++].second;
                    else {
                        if (hw >= decommitment.hash_witness.size()) throw VerificationError("Witness is too short");
                        right = to_hash(decommitment.hash_witness[hw++]);
                    }
                }
                std::vector<M31> node_values;
                if (ci < layer_column_queries.size() && layer_column_queries[ci] == node_index) {
                    ci++;
                    if (qv + n_columns_in_layer > queried_values_in.size()) throw VerificationError("too few queried values");
                    for (size_t k = 0; k < n_columns_in_layer; k++) node_values.push_back(M31((u64)queried_values_in[qv++]));
                } else {
                    if (cw + n_columns_in_layer > decommitment.column_witness.size()) throw VerificationError("Witness is too short");
                    for (size_t k = 0; k < n_columns_in_layer; k++) node_values.push_back(M31((u64)decommitment.column_witness[cw++]));
                }
                layer_total_queries.push_back({node_index, hash_node(have_last ? &left : nullptr, have_last ? &right : nullptr, node_values.data(), node_values.size())});
            }
            last_layer_hashes = layer_total_queries;
            have_last = true;
        }
        if (hw != decommitment.hash_witness.size()) throw VerificationError("Witness is too long.");
        if (qv != queried_values_in.size()) throw VerificationError("too many Queried values");
        if (cw != decommitment.column_witness.size()) throw VerificationError("Witness is too long.");
        if (last_layer_hashes.size() != 1) throw VerificationError("Merkle verification did not end in a single root");
        if (last_layer_hashes[0].second != root) throw VerificationError("Root mismatch.");
    }
};

// queries.rs
struct OQueries {
    std::vector<size_t> positions;
    u32 log_domain_size;
    static OQueries generate(OChannel& ch, u32 log_domain_size, size_t n_queries) {
        std::set<size_t> q;
        size_t cnt = 0;
        u32 max_query = (u32)(((u64)1 << log_domain_size) - 1);
        for (;;) {
            Hash h = ch.draw_random_bytes();
            for (int i = 0; i < 8; i++) {
                u32 bits = (u32)h.b[4 * i] | ((u32)h.b[4 * i + 1] << 8) | ((u32)h.b[4 * i + 2] << 16) | ((u32)h.b[4 * i + 3] << 24);
                q.insert(bits & max_query);
                if (++cnt == n_queries) return OQueries{std::vector<size_t>(q.begin(), q.end()), log_domain_size};
            }
        }
    }
    OQueries fold(u32 n_folds) const {
        OQueries out{{}, log_domain_size - n_folds};
        for (size_t p : positions) {
            size_t f = p >> n_folds;
            if (out.positions.empty() || out.positions.back() != f) out.positions.push_back(f);
        }
        return out;
    }
};

struct SparseEvaluation {
    std::vector<std::vector<QM31>> subset_evals;
    std::vector<size_t> subset_domain_initial_indexes;
    // fold_line on a 2-point line domain with initial index `idx` of the source coset (fri.rs:1079-1090)
    std::vector<QM31> fold_line_(QM31 alpha, OCoset source_coset) const {
        std::vector<QM31> out;
        for (size_t k = 0; k < subset_evals.size(); k++) {
            u64 initial = source_coset.index_at(subset_domain_initial_indexes[k]);
            M31 x = point_from_index(initial).x;  // domain.at(bit_reverse(0)) of the 2-point fold domain
            QM31 f0 = subset_evals[k][0], f1 = subset_evals[k][1];
            ibutterfly_q(f0, f1, x.inverse());
            out.push_back(f0 + alpha * f1);
        }
        return out;
    }
    std::vector<QM31> fold_circle_(QM31 alpha, ODomain source_domain) const {
        std::vector<QM31> out;
        for (size_t k = 0; k < subset_evals.size(); k++) {
            u64 initial = source_domain.index_at(subset_domain_initial_indexes[k]);
            Point p = point_from_index(initial);
            QM31 f0 = subset_evals[k][0], f1 = subset_evals[k][1];
            ibutterfly_q(f0, f1, p.y.inverse());
            out.push_back(alpha * f1 + f0);  // dst (zero) * alpha^2 + f'
        }
        return out;
    }
};

inline void compute_decommitment_positions_and_rebuild_evals(const OQueries& queries, const std::vector<QM31>& query_evals,
                                                             const std::vector<cm31::QM31>& witness, size_t& wi, u32 fold_step,
                                                             std::vector<size_t>& decommitment_positions, SparseEvaluation& sparse) {
    size_t qe = 0, i = 0;
    const std::vector<size_t>& q = queries.positions;
    while (i < q.size()) {
        size_t j = i;
        while (j < q.size() && (q[j] >> fold_step) == (q[i] >> fold_step)) j++;
        size_t subset_start = (q[i] >> fold_step) << fold_step;
        std::vector<QM31> subset_eval;
        size_t qi = i;
        for (size_t position = subset_start; position < subset_start + ((size_t)1 << fold_step); position++) {
            decommitment_positions.push_back(position);
            if (qi < j && q[qi] == position) {
                qi++;
                subset_eval.push_back(query_evals.at(qe++));
            } else {
                if (wi >= witness.size()) throw VerificationError("insufficient FRI witness");
                subset_eval.push_back(to_orc(witness[wi++]));
            }
        }
        sparse.subset_evals.push_back(subset_eval);
        sparse.subset_domain_initial_indexes.push_back(bitrev((u32)subset_start, queries.log_domain_size));
        i = j;
    }
}

// LinePoly::eval_at_point (poly/line.rs:121-128)
inline QM31 line_poly_eval(const std::vector<cm31::QM31>& coeffs, QM31 x) {
    u32 log_size = 0;
    while (((size_t)1 << log_size) < coeffs.size()) log_size++;
    std::vector<QM31> doublings;
    for (u32 i = 0; i < log_size; i++) {
        doublings.push_back(x);
        x = qdouble_x(x);
    }
    std::function<QM31(size_t, size_t, size_t)> fold = [&](size_t off, size_t n, size_t lvl) -> QM31 {
        if (n == 1) return to_orc(coeffs[off]);
        return fold(off, n / 2, lvl + 1) + fold(off + n / 2, n / 2, lvl + 1) * doublings[lvl];
    };
    return fold(0, coeffs.size(), 0);
}

// A component as the verifier sees it (air/mod.rs:26-57 `Component`): reuse the ComponentProver
// interface of the oracle backend (only the Component half is called here).
typedef cm31::ComponentProver<OracleBackend> OComponent;

struct CommitmentSchemeVerifier {
    std::vector<MerkleVerifier> trees;
    cm31::PcsConfig config;
    explicit CommitmentSchemeVerifier(cm31::PcsConfig c) : config(c) {}
    void commit(const cm31::Hash32& commitment, const std::vector<u32>& log_sizes, OChannel& channel) {
        channel.mix_root(to_hash(commitment));
        std::vector<u32> ext;
        for (u32 s : log_sizes) ext.push_back(s + config.fri_config.log_blowup_factor);
        trees.push_back(MerkleVerifier(to_hash(commitment), ext));
    }

    void verify_values(const cm31::MaskPoints& sampled_points, const cm31::CommitmentSchemeProof& proof, OChannel& channel) const {
        std::vector<QM31> flat;
        for (auto& t : proof.sampled_values)
            for (auto& c : t)
                for (auto& v : c) flat.push_back(to_orc(v));
        channel.mix_felts(flat);
        QM31 random_coeff = channel.draw_secure_felt();
        // bounds: distinct column log sizes, descending, minus blowup
        std::set<u32> sizes;
        for (auto& t : trees)
            for (u32 s : t.column_log_sizes) sizes.insert(s);
        std::vector<u32> bounds;  // log degree bounds, descending
        for (auto it = sizes.rbegin(); it != sizes.rend(); ++it) bounds.push_back(*it - config.fri_config.log_blowup_factor);
        const cm31::FriConfig& fc = config.fri_config;
        const cm31::FriProof& fp = proof.fri_proof;

        // ---- FriVerifier::commit (fri.rs:370-436)
        channel.mix_root(to_hash(fp.first_layer.commitment));
        std::vector<u32> column_domain_log;  // commitment domain log sizes, descending
        for (u32 b : bounds) column_domain_log.push_back(b + fc.log_blowup_factor);
        QM31 first_layer_alpha = channel.draw_secure_felt();
        u32 layer_bound = bounds[0] - 1;  // fold_to_line
        u32 layer_domain_log = layer_bound + fc.log_blowup_factor;
        struct Inner {
            u32 degree_bound, domain_log;
            QM31 alpha;
        };
        std::vector<Inner> inner;
        for (auto& lp : fp.inner_layers) {
            channel.mix_root(to_hash(lp.commitment));
            inner.push_back(Inner{layer_bound, layer_domain_log, channel.draw_secure_felt()});
            if (layer_bound < 1) throw VerificationError("proof contains an invalid number of FRI layers");
            layer_bound -= 1;
            layer_domain_log -= 1;
        }
        if (layer_bound != fc.log_last_layer_degree_bound) throw VerificationError("proof contains an invalid number of FRI layers");
        u32 last_layer_domain_log = layer_domain_log;
        if (fp.last_layer_poly.size() > ((size_t)1 << fc.log_last_layer_degree_bound)) throw VerificationError("degree of last layer is invalid");
        {
            std::vector<QM31> llp;
            for (auto& v : fp.last_layer_poly) llp.push_back(to_orc(v));
            channel.mix_felts(llp);
        }
        // ---- proof of work
        channel.mix_u64(proof.proof_of_work);
        if (channel.trailing_zeros() < config.pow_bits) throw VerificationError("Proof of work verification failed.");
        // ---- query positions
        u32 max_column_log_size = column_domain_log[0];
        OQueries queries = OQueries::generate(channel, max_column_log_size, fc.n_queries);
        std::map<u32, std::vector<size_t>> query_positions_per_log_size;
        for (u32 ls : column_domain_log) query_positions_per_log_size[ls] = queries.fold(queries.log_domain_size - ls).positions;
        // ---- Merkle decommitments of the trace trees
        if (proof.decommitments.size() != trees.size() || proof.queried_values.size() != trees.size()) throw VerificationError("proof structure");
        for (size_t t = 0; t < trees.size(); t++) trees[t].verify(query_positions_per_log_size, proof.queried_values[t], proof.decommitments[t]);
        // ---- fri_answers (pcs/quotients.rs:104-166)
        struct ColInfo {
            u32 log_size;
            const std::vector<cm31::SecurePoint>* points;
            const std::vector<cm31::QM31>* values;
        };
        std::vector<ColInfo> cols;
        for (size_t t = 0; t < trees.size(); t++) {
            if (sampled_points[t].size() != trees[t].column_log_sizes.size() || proof.sampled_values[t].size() != sampled_points[t].size())
                throw VerificationError("Unexpected sampled_values structure");
            for (size_t c = 0; c < sampled_points[t].size(); c++) {
                if (sampled_points[t][c].size() != proof.sampled_values[t][c].size()) throw VerificationError("Unexpected sampled_values structure");
                cols.push_back(ColInfo{trees[t].column_log_sizes[c], &sampled_points[t][c], &proof.sampled_values[t][c]});
            }
        }
        std::vector<size_t> order(cols.size());
        for (size_t i = 0; i < order.size(); i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cols[a].log_size > cols[b].log_size; });
        std::vector<size_t> qv_pos(trees.size(), 0);
        std::vector<std::vector<QM31>> fri_answers;  // per log size descending
        size_t oi = 0;
        while (oi < order.size()) {
            u32 log_size = cols[order[oi]].log_size;
            std::vector<SampleBatch> batches;  // ColumnSampleBatch::new_vec
            size_t local = 0;
            while (oi < order.size() && cols[order[oi]].log_size == log_size) {
                const ColInfo& ci = cols[order[oi]];
                for (size_t s = 0; s < ci.points->size(); s++) {
                    QPoint p = to_orc((*ci.points)[s]);
                    size_t k = 0;
                    for (; k < batches.size(); k++)
                        if (batches[k].point.x == p.x && batches[k].point.y == p.y) break;
                    if (k == batches.size()) batches.push_back(SampleBatch{p, {}});
                    batches[k].columns_and_values.push_back({local, to_orc((*ci.values)[s])});
                }
                local++;
                oi++;
            }
            QuotientConstants qc = quotient_constants(batches, random_coeff);
            ODomain commitment_domain = ODomain::canonic(log_size);
            std::vector<QM31> answers;
            for (size_t qp : query_positions_per_log_size.at(log_size)) {
                Point dp = commitment_domain.at(bitrev((u32)qp, log_size));
                std::vector<M31> row;
                for (size_t t = 0; t < trees.size(); t++) {
                    auto nit = trees[t].n_columns_per_log_size.find(log_size);
                    size_t n_cols = nit == trees[t].n_columns_per_log_size.end() ? 0 : nit->second;
                    for (size_t k = 0; k < n_cols; k++) {
                        if (qv_pos[t] >= proof.queried_values[t].size()) throw VerificationError("too few queried values");
                        row.push_back(M31((u64)proof.queried_values[t][qv_pos[t]++]));
                    }
                }
                answers.push_back(accumulate_row_quotients(batches, row.data(), qc, dp));
            }
            fri_answers.push_back(answers);
        }
        // ---- FriVerifier::decommit
        // first layer (fri.rs:708-770)
        if (fri_answers.size() != column_domain_log.size()) throw VerificationError("fri answers shape");
        size_t wi = 0;
        std::map<u32, std::vector<size_t>> decommitment_positions_by_log_size;
        std::vector<SparseEvaluation> sparse_evals_by_column;
        std::vector<u32> decommitted_values;
        for (size_t c = 0; c < column_domain_log.size(); c++) {
            OQueries cq = queries.fold(queries.log_domain_size - column_domain_log[c]);
            std::vector<size_t> positions;
            SparseEvaluation se;
            compute_decommitment_positions_and_rebuild_evals(cq, fri_answers[c], fp.first_layer.fri_witness, wi, 1, positions, se);
            decommitment_positions_by_log_size[column_domain_log[c]] = positions;
            for (auto& sub : se.subset_evals)
                for (auto& v : sub) {
                    u32 o[4];
                    v.to_u32(o);
                    decommitted_values.insert(decommitted_values.end(), o, o + 4);
                }
            sparse_evals_by_column.push_back(se);
        }
        if (wi != fp.first_layer.fri_witness.size()) throw VerificationError("evaluations are invalid in the first layer");
        {
            std::vector<u32> sizes4;
            for (u32 ls : column_domain_log)
                for (int k = 0; k < 4; k++) sizes4.push_back(ls);
            MerkleVerifier mv(to_hash(fp.first_layer.commitment), sizes4);
            mv.verify(decommitment_positions_by_log_size, decommitted_values, fp.first_layer.decommitment);
        }
        // inner layers (fri.rs:476-520)
        OQueries layer_queries = queries.fold(1);
        std::vector<QM31> layer_query_evals(layer_queries.positions.size(), QM31::zero());
        size_t next_col = 0;
        QM31 previous_folding_alpha = first_layer_alpha;
        for (size_t li = 0; li < inner.size(); li++) {
            const Inner& layer = inner[li];
            while (next_col < bounds.size() && bounds[next_col] - 1 == layer.degree_bound) {
                std::vector<QM31> folded = sparse_evals_by_column[next_col].fold_circle_(previous_folding_alpha, ODomain::canonic(column_domain_log[next_col]));
                QM31 a2 = previous_folding_alpha * previous_folding_alpha;
                if (folded.size() != layer_query_evals.size()) throw VerificationError("fold size mismatch");
                for (size_t k = 0; k < folded.size(); k++) layer_query_evals[k] = layer_query_evals[k] * a2 + folded[k];
                next_col++;
            }
            // verify_and_fold (fri.rs:790-846)
            const cm31::FriLayerProof& lp = fp.inner_layers[li];
            size_t lwi = 0;
            std::vector<size_t> positions;
            SparseEvaluation se;
            compute_decommitment_positions_and_rebuild_evals(layer_queries, layer_query_evals, lp.fri_witness, lwi, 1, positions, se);
            if (lwi != lp.fri_witness.size()) throw VerificationError("evaluations are invalid in inner layer");
            std::vector<u32> vals;
            for (auto& sub : se.subset_evals)
                for (auto& v : sub) {
                    u32 o[4];
                    v.to_u32(o);
                    vals.insert(vals.end(), o, o + 4);
                }
            MerkleVerifier mv(to_hash(lp.commitment), std::vector<u32>(4, layer.domain_log));
            std::map<u32, std::vector<size_t>> m;
            m[layer.domain_log] = positions;
            mv.verify(m, vals, lp.decommitment);
            layer_query_evals = se.fold_line_(layer.alpha, OCoset::half_odds(layer.domain_log));
            layer_queries = layer_queries.fold(1);
            previous_folding_alpha = layer.alpha;
        }
        if (next_col != bounds.size()) throw VerificationError("not all first layer columns were folded");
        // last layer (fri.rs:522-541)
        OCoset last_dom = OCoset::half_odds(last_layer_domain_log);
        for (size_t k = 0; k < layer_queries.positions.size(); k++) {
            M31 x = last_dom.at(bitrev((u32)layer_queries.positions[k], last_layer_domain_log)).x;
            if (layer_query_evals[k] != line_poly_eval(fp.last_layer_poly, QM31::from_m31(x))) throw VerificationError("evaluations in the last layer are invalid");
        }
    }
};

inline QPoint o_get_random_point(OChannel& ch) {  // circle.rs:169-181
    QM31 t = ch.draw_secure_felt();
    QM31 t2 = t.square();
    QM31 inv = (t2 + QM31::one()).inverse();
    return QPoint{(QM31::one() - t2) * inv, (t + t) * inv};
}

// verify (prover/mod.rs:87-143). `trees[0..3]` must already be committed in `cs`.
inline void verify(const std::vector<const OComponent*>& components, OChannel& channel, CommitmentSchemeVerifier& cs,
                   const cm31::StarkProof& proof) {
    cm31::ComponentProvers<OracleBackend> comps{components, cs.trees[0].column_log_sizes.size()};
    QM31 random_coeff = channel.draw_secure_felt();
    cs.commit(proof.commitments.back(), std::vector<u32>(4, comps.composition_log_degree_bound()), channel);
    QPoint oods = o_get_random_point(channel);
    cm31::SecurePoint oods_pt{from_orc(oods.x), from_orc(oods.y)};
    cm31::MaskPoints sample_points = comps.mask_points(oods_pt);
    while (sample_points.size() > cs.trees.size() - 1 && sample_points.back().empty()) sample_points.pop_back();
    sample_points.push_back(std::vector<std::vector<cm31::SecurePoint>>(4, std::vector<cm31::SecurePoint>{oods_pt}));
    if (proof.sampled_values.empty() || proof.sampled_values.back().size() != 4) throw VerificationError("Unexpected sampled_values structure");
    const auto& cm = proof.sampled_values.back();
    for (auto& c : cm)
        if (c.size() != 1) throw VerificationError("Unexpected sampled_values structure");
    QM31 composition_oods_eval = QM31::from_partial_evals(to_orc(cm[0][0]), to_orc(cm[1][0]), to_orc(cm[2][0]), to_orc(cm[3][0]));
    QM31 expect = to_orc(comps.eval_composition_polynomial_at_point(oods_pt, proof.sampled_values, from_orc(random_coeff)));
    if (composition_oods_eval != expect) throw VerificationError("The composition polynomial OODS value does not match the trace OODS values");
    cs.verify_values(sample_points, proof, channel);
}

}  // namespace orc

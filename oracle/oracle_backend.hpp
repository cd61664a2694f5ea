// ORACLE (test infrastructure, not product): plugs the scalar CPU restatement
// (oracle/cpu_backend.hpp) into the backend-generic STARK driver (cairo-m_b200/csrc/host/stark.hpp)
// the same way Stwo instantiates its generic core with `CpuBackend`
// (external/stwo/crates/prover/src/core/backend/cpu/mod.rs).  Used as the bit-exact checker of
// whole proofs and as the CPU baseline of bench.py.
#pragma once
#include "cpu_backend.hpp"
#include "host/stark.hpp"

namespace orc {

inline QM31 to_orc(cm31::QM31 v) { return QM31::from_u32(v.a, v.b, v.c, v.d); }
inline cm31::QM31 from_orc(QM31 v) {
    u32 o[4];
    v.to_u32(o);
    return cm31::qm_make(o[0], o[1], o[2], o[3]);
}
inline QPoint to_orc(cm31::SecurePoint p) { return QPoint{to_orc(p.x), to_orc(p.y)}; }

struct OracleBackend {
    typedef std::vector<M31> Col;
    typedef std::vector<Hash> HashCol;
    typedef OTwiddles Twiddles;

    static size_t len(const Col& c) { return c.size(); }
    static Col zeros(size_t n) { return Col(n); }
    static Col uninit(size_t n) { return Col(n); }
    static std::vector<Col> uninit_many(size_t count, size_t n) { return std::vector<Col>(count, Col(n)); }
    static Col from_host(const u32* src, size_t n) {
        Col c(n);
        for (size_t i = 0; i < n; i++) c[i] = M31((u64)src[i]);
        return c;
    }
    static void to_host(const Col& c, u32* out) {
        for (size_t i = 0; i < c.size(); i++) out[i] = c[i].v;
    }
    static void precompute_twiddles(u32 log_size, Twiddles& out) { out = orc::precompute_twiddles(log_size); }
    // across columns when there are at least as many as threads, inside each FFT otherwise
    static bool many(size_t n_cols) { return n_cols >= (size_t)omp_get_max_threads(); }
    static void interpolate_columns(const std::vector<Col*>& cols, u32, const Twiddles& tw) {
#pragma omp parallel for schedule(dynamic) if (many(cols.size()))
        for (size_t i = 0; i < cols.size(); i++) interpolate(*cols[i], tw);
    }
    static void interpolate_columns_to(const std::vector<const Col*>& evals, const std::vector<Col*>& outs, u32, const Twiddles& tw) {
#pragma omp parallel for schedule(dynamic) if (many(evals.size()))
        for (size_t i = 0; i < evals.size(); i++) {
            *outs[i] = *evals[i];
            interpolate(*outs[i], tw);
        }
    }
    static void evaluate_polynomials(const std::vector<const Col*>& polys, const std::vector<Col*>& outs, u32, u32 log_eval, const Twiddles& tw) {
#pragma omp parallel for schedule(dynamic) if (many(polys.size()))
        for (size_t i = 0; i < polys.size(); i++) *outs[i] = evaluate(*polys[i], log_eval, tw);
    }
    static void eval_at_points(const std::vector<const Col*>& polys, const std::vector<u32>&, const std::vector<cm31::SecurePoint>& points,
                               const std::vector<u32>& point_idx, std::vector<cm31::QM31>& out) {
        out.resize(polys.size());
#pragma omp parallel for schedule(dynamic)
        for (size_t i = 0; i < polys.size(); i++) out[i] = from_orc(eval_at_point(*polys[i], to_orc(points[point_idx[i]])));
    }
    static HashCol commit_on_layer(u32 log_size, const HashCol* prev, const std::vector<const Col*>& cols) {
        return orc::commit_on_layer(log_size, prev, cols);
    }
    static void gather(const std::vector<const Col*>& cols, const std::vector<u32>& idx, std::vector<std::vector<u32>>& out) {
        out.assign(cols.size(), std::vector<u32>(idx.size()));
        for (size_t c = 0; c < cols.size(); c++)
            for (size_t q = 0; q < idx.size(); q++) out[c][q] = (*cols[c])[idx[q]].v;
    }
    static std::vector<HashCol> commit_layers_fused(u32 log_size, const HashCol* prev, const std::vector<const Col*>& cols, u32 n_levels) {
        std::vector<HashCol> out;
        for (u32 l = 0; l < n_levels; l++) {
            out.push_back(orc::commit_on_layer(log_size - l, l == 0 ? prev : &out[l - 1], l == 0 ? cols : std::vector<const Col*>()));
        }
        return out;
    }
    static std::vector<HashCol> commit_top_layers(u32 top_log, const HashCol* prev, const std::vector<std::vector<const Col*>>& cols_by_layer) {
        std::vector<HashCol> out(top_log + 1);
        for (int l = (int)top_log; l >= 0; l--) {
            out[l] = orc::commit_on_layer((u32)l, prev, cols_by_layer[l]);
            prev = &out[l];
        }
        return out;
    }
    // single-proof sharding is a property of the CUDA data plane: the oracle is always one "rank" holding everything
    static int shard_world() { return 1; }
    static int shard_rank() { return 0; }
    static u32 shard_stripe_log() { return 31; }
    static void shard_begin_proof() {}
    static void shard_end_proof() {}
    static void shard_barrier() {}
    static void component_scope(int) {}
    static void component_scope_index(size_t) {}
    static void set_component_owners(const std::vector<int>&) {}
    static void set_owner(const Col&, int) {}
    static void allreduce_m31(std::array<Col, 4>&) {}
    static void allreduce_bins(Col&) {}
    static HashCol commit_on_layer_striped(u32 log_size, const HashCol* prev, const std::vector<const Col*>& cols) { return commit_on_layer(log_size, prev, cols); }
    static std::vector<HashCol> commit_layers_fused_striped(u32 log_size, const HashCol* prev, const std::vector<const Col*>& cols, u32 n_levels) {
        return commit_layers_fused(log_size, prev, cols, n_levels);
    }
    static void join_striped_layer(HashCol&) {}
    static bool is_striped(const HashCol&) { return false; }
    static const u32* hash_node_source(const HashCol& layer, size_t) { return (const u32*)layer.data(); }
    static void lane(u32) {}  // the CUDA backend's stream lanes have no CPU counterpart
    static void lanes_join() {}
    static void prepare() {}
    static const u32* col_words(const Col& c) { return (const u32*)c.data(); }
    static const u32* hash_words(const HashCol& c) { return (const u32*)c.data(); }
    static void gather_runs(const std::vector<const u32*>& srcs, const std::vector<u32>& src_id, const std::vector<u32>& word,
                            const std::vector<u32>& out_off, const std::vector<u32>& cnt, const std::vector<u32>& grid_desc,
                            const std::vector<u32>& grid_cols, const std::vector<u32>& grid_rows, std::vector<u32>& out) {
        for (size_t k = 0; k < src_id.size(); k++)
            for (u32 j = 0; j < cnt[k]; j++) out[out_off[k] + j] = srcs[src_id[k]][word[k] + j];
        for (size_t g = 0; g + 5 <= grid_desc.size(); g += 5) {  // (col_off, n_cols, row_off, n_rows, out_base), row-major output
            const u32 col_off = grid_desc[g], n_cols = grid_desc[g + 1], row_off = grid_desc[g + 2], n_rows = grid_desc[g + 3], base = grid_desc[g + 4];
            for (u32 k = 0; k < n_rows; k++)
                for (u32 c = 0; c < n_cols; c++) out[base + (size_t)k * n_cols + c] = srcs[grid_cols[col_off + c]][grid_rows[row_off + k]];
        }
    }
    // the CPU backend gathers synchronously: the "asynchronous" form fills `store` at once
    static const u32* gather_runs_async(const std::vector<const u32*>& srcs, const std::vector<u32>& src_id, const std::vector<u32>& word,
                                        const std::vector<u32>& out_off, const std::vector<u32>& cnt, const std::vector<u32>& grid_desc,
                                        const std::vector<u32>& grid_cols, const std::vector<u32>& grid_rows, size_t n_words,
                                        std::vector<u32>& store) {
        store.assign(n_words, 0);
        gather_runs(srcs, src_id, word, out_off, cnt, grid_desc, grid_cols, grid_rows, store);
        return store.data();
    }
    static void gather_wait() {}
    static void range_push(const char*) {}
    static void range_pop() {}
    static u32 fri_tail_log() { return 0; }  // the CPU backend folds layer by layer
    template <class InnerLayer, class SecureEval, class Tw>
    static void fri_tail(cm31::Blake2sChannel&, std::array<Col, 4>&, u32&, u32, const std::vector<SecureEval>&, size_t&, const Tw&,
                         std::vector<InnerLayer>&) {
        throw std::logic_error("fri_tail: not available on the CPU backend");
    }
    static bool defer_proof_tail() { return false; }
    static void finish_deferred_tails() {}
    static cm31::Hash32 read_root(const HashCol& root_layer) {
        cm31::Hash32 h;
        memcpy(h.b, root_layer[0].b, 32);
        return h;
    }
    static void gather_hashes(const HashCol& layer, const std::vector<u32>& idx, std::vector<cm31::Hash32>& out) {
        out.resize(idx.size());
        for (size_t q = 0; q < idx.size(); q++) memcpy(out[q].b, layer[idx[q]].b, 32);
    }
    static std::array<Col, 4> accumulate_quotients(u32 log_size, const std::vector<const Col*>& cols, cm31::QM31 random_coeff,
                                                   const std::vector<cm31::ColumnSampleBatch>& batches, u32) {
        std::vector<SampleBatch> ob;
        for (auto& b : batches) {
            SampleBatch s;
            s.point = to_orc(b.point);
            for (auto& cv : b.columns_and_values) s.columns_and_values.push_back({cv.first, to_orc(cv.second)});
            ob.push_back(s);
        }
        SecureColumn sc = orc::accumulate_quotients(log_size, cols, to_orc(random_coeff), ob);
        return {sc.c[0], sc.c[1], sc.c[2], sc.c[3]};
    }
    static SecureColumn to_sc(const std::array<Col, 4>& a) {
        SecureColumn s;
        for (int k = 0; k < 4; k++) s.c[k] = a[k];
        return s;
    }
    static std::array<Col, 4> fold_line(const std::array<Col, 4>& src, u32, cm31::QM31 alpha, const Twiddles&) {
        SecureColumn r = orc::fold_line(to_sc(src), to_orc(alpha));
        return {r.c[0], r.c[1], r.c[2], r.c[3]};
    }
    static void fold_circle_into_line(std::array<Col, 4>& dst, const std::array<Col, 4>& src, u32, cm31::QM31 alpha, const Twiddles&) {
        SecureColumn d = to_sc(dst);
        orc::fold_circle_into_line(d, to_sc(src), to_orc(alpha));
        for (int k = 0; k < 4; k++) dst[k] = d.c[k];
    }
    static void accumulate(std::array<Col, 4>& dst, const std::array<Col, 4>& src) {  // cpu/accumulation.rs:8-15
        for (int k = 0; k < 4; k++)
            for (size_t i = 0; i < dst[k].size(); i++) dst[k][i] = dst[k][i] + src[k][i];
    }
    static std::vector<cm31::QM31> generate_secure_powers(cm31::QM31 felt, size_t n) {  // cpu/accumulation.rs:17-25
        std::vector<cm31::QM31> out;
        QM31 acc = QM31::one(), f = to_orc(felt);
        for (size_t i = 0; i < n; i++) {
            out.push_back(from_orc(acc));
            acc = acc * f;
        }
        return out;
    }
    static u64 grind(const cm31::Hash32& digest, u32 pow_bits) {
        OChannel ch;
        memcpy(ch.digest.b, digest.b, 32);
        return orc::grind(ch, pow_bits);
    }
};

}  // namespace orc

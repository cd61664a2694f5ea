// Backend-generic Circle-STARK prover core: Merkle commitment, PCS, FRI, prove().
//
// Host-side mirror (C++; no Rust toolchain in this image) of the generic-over-`B: Backend` layer
// of Stwo that drives the Backend ops:
//   MerkleProver            external/stwo/crates/prover/src/core/vcs/prover.rs:14-173
//   CommitmentSchemeProver  core/pcs/prover.rs:25-249, quotient batching core/pcs/quotients.rs:39-102
//   FriProver               core/fri.rs:142-368, layer provers fri.rs:877-1062
//   prove                   core/prover/mod.rs:28-85, ComponentProvers core/air/components.rs:16-138,
//   DomainEvaluationAccumulator core/air/accumulation.rs:49-154
// `B` supplies the ops (CudaBackend calls libcm31's C ABI; the test oracle supplies a scalar
// CpuBackend) — the order of channel mixes/draws below is part of the proof contract
// (SURVEY.md Appendix A).
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <array>
#include <functional>
#include <iterator>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <unordered_map>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "channel.hpp"
#include "qm_ops.hpp"

namespace cm31 {

// Host-side section timers (CM31_HOST_TIMING=1): where the host thread spends time between launches.
struct HostTimer {
    static bool enabled() {
        static const bool on = getenv("CM31_HOST_TIMING") != nullptr;
        return on;
    }
    static std::map<std::string, std::pair<double, size_t>>& table() {
        static std::map<std::string, std::pair<double, size_t>> t;
        return t;
    }
    const char* name;
    std::chrono::steady_clock::time_point t0;
    explicit HostTimer(const char* n) : name(n) {
        if (enabled()) t0 = std::chrono::steady_clock::now();
    }
    ~HostTimer() {
        if (!enabled()) return;
        auto& e = table()[name];
        e.first += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        e.second++;
    }
    static void report() {
        if (!enabled()) return;
        for (auto& kv : table()) fprintf(stderr, "[host] %-28s %9.3f ms  x%zu\n", kv.first.c_str(), kv.second.first, kv.second.second);
        table().clear();
    }
};


// A named range on the profiler timeline (NVTX on the CUDA backend, nothing on the CPU oracle), with the span names and
// `class` fields of the reference's `tracing` spans (S/prover/src/core/pcs/prover.rs:41,89,127,177,219,223,
// prover/mod.rs:45-51, air/accumulation.rs:111, pcs/quotients.rs:83; P/src/prover.rs:30), so an nsys / ncu timeline of this
// prover reads like the reference's SpanAccumulator CSV.
template <class B>
struct Span {
    explicit Span(const char* name) { B::range_push(name); }
    ~Span() { B::range_pop(); }
    Span(const Span&) = delete;
    Span& operator=(const Span&) = delete;
};

struct FriConfig {
    u32 log_blowup_factor = 1;
    u32 log_last_layer_degree_bound = 0;
    size_t n_queries = 80;
    size_t last_layer_domain_size() const { return (size_t)1 << (log_last_layer_degree_bound + log_blowup_factor); }
    void mix_into(Blake2sChannel& ch) const {  // fri.rs:80-89
        ch.mix_u64(log_blowup_factor);
        ch.mix_u64(n_queries);
        ch.mix_u64(log_last_layer_degree_bound);
    }
};
struct PcsConfig {
    u32 pow_bits = 16;
    FriConfig fri_config;
    void mix_into(Blake2sChannel& ch) const {  // pcs/mod.rs:42-49
        ch.mix_u64(pow_bits);
        fri_config.mix_into(ch);
    }
    // crates/prover/src/prover_config.rs:13-20 REGULAR_96_BITS
    static PcsConfig regular_96_bits() {
        PcsConfig c;
        c.pow_bits = 16;
        c.fri_config.log_blowup_factor = 1;
        c.fri_config.log_last_layer_degree_bound = 0;
        c.fri_config.n_queries = 80;
        return c;
    }
};

struct MerkleDecommitment {
    std::vector<Hash32> hash_witness;
    std::vector<u32> column_witness;
};
struct FriLayerProof {
    std::vector<QM31> fri_witness;
    MerkleDecommitment decommitment;
    Hash32 commitment;
};
struct FriProof {
    FriLayerProof first_layer;
    std::vector<FriLayerProof> inner_layers;
    std::vector<QM31> last_layer_poly;  // LinePoly coefficients (bit-reversed storage order)
};
struct CommitmentSchemeProof {
    PcsConfig config;
    std::vector<Hash32> commitments;
    std::vector<std::vector<std::vector<QM31>>> sampled_values;  // [tree][column][sample]
    std::vector<MerkleDecommitment> decommitments;
    std::vector<std::vector<u32>> queried_values;
    u64 proof_of_work = 0;
    FriProof fri_proof;
    // Deferred tail (B::defer_proof_tail()): everything after the proof-of-work nonce -- query generation, decommitment
    // planning, the gather of the queried values and the assembly of queried_values / decommitments / fri_proof -- is left to
    // a closure that owns the trees and FRI layers.  resolve() must run before the proof is read; the CUDA prover runs the
    // stages while the NEXT proof keeps the GPU busy (cm31_prove_cairo_m_async).
    // The tail runs in STAGES (each call does one and returns true after the last): planning + gather, then assembly, so that
    // each can be placed behind a different commitment of the next proof.
    std::shared_ptr<std::function<bool(CommitmentSchemeProof&)>> pending_tail;
    bool step() {  // one stage; true once nothing is pending
        if (!pending_tail) return true;
        if ((*pending_tail)(*this)) pending_tail.reset();
        return !pending_tail;
    }
    void resolve() {
        while (!step()) {
        }
    }
};
typedef CommitmentSchemeProof StarkProof;

struct ConstraintsNotSatisfied : std::runtime_error {
    ConstraintsNotSatisfied() : std::runtime_error("Constraints not satisfied.") {}
};

// Canonical byte serialisation (used for bit-exact proof comparison and as the proof blob
// handed back across the C ABI; the reference's serde/JSON wire format is out of scope).
class ProofWriter {
   public:
    std::vector<uint8_t> bytes;
    ProofWriter() { bytes.reserve((size_t)4 << 20); }
    // little-endian words; the host is little-endian (x86-64 / aarch64), so a word run is one memcpy
    void u32s(const u32* v, size_t n) {
        size_t at = bytes.size();
        bytes.resize(at + 4 * n);
        if (n) memcpy(bytes.data() + at, v, 4 * n);
    }
    void u32v(u32 v) { u32s(&v, 1); }
    void u64v(u64 v) {
        u32v((u32)v);
        u32v((u32)(v >> 32));
    }
    void qm(QM31 v) {
        u32v(v.a);
        u32v(v.b);
        u32v(v.c);
        u32v(v.d);
    }
    void hash(const Hash32& h) { bytes.insert(bytes.end(), h.b, h.b + 32); }
    void decommitment(const MerkleDecommitment& d) {
        u64v(d.hash_witness.size());
        for (auto& h : d.hash_witness) hash(h);
        u64v(d.column_witness.size());
        u32s(d.column_witness.data(), d.column_witness.size());
    }
    void fri_layer(const FriLayerProof& l) {
        u64v(l.fri_witness.size());
        for (auto& v : l.fri_witness) qm(v);
        decommitment(l.decommitment);
        hash(l.commitment);
    }
    void proof(const CommitmentSchemeProof& p) {
        u32v(p.config.pow_bits);
        u32v(p.config.fri_config.log_blowup_factor);
        u32v(p.config.fri_config.log_last_layer_degree_bound);
        u64v(p.config.fri_config.n_queries);
        u64v(p.commitments.size());
        for (auto& h : p.commitments) hash(h);
        u64v(p.sampled_values.size());
        for (auto& tree : p.sampled_values) {
            u64v(tree.size());
            for (auto& col : tree) {
                u64v(col.size());
                for (auto& v : col) qm(v);
            }
        }
        u64v(p.decommitments.size());
        for (auto& d : p.decommitments) decommitment(d);
        u64v(p.queried_values.size());
        for (auto& q : p.queried_values) {
            u64v(q.size());
            u32s(q.data(), q.size());
        }
        u64v(p.proof_of_work);
        fri_layer(p.fri_proof.first_layer);
        u64v(p.fri_proof.inner_layers.size());
        for (auto& l : p.fri_proof.inner_layers) fri_layer(l);
        u64v(p.fri_proof.last_layer_poly.size());
        for (auto& v : p.fri_proof.last_layer_poly) qm(v);
    }
};

class ProofReader {
   public:
    const uint8_t* p;
    size_t n, pos = 0;
    ProofReader(const uint8_t* data, size_t len) : p(data), n(len) {}
    void need(size_t k) {
        if (pos + k > n) throw std::runtime_error("proof blob truncated");
    }
    u32 u32v() {
        need(4);
        u32 v = (u32)p[pos] | ((u32)p[pos + 1] << 8) | ((u32)p[pos + 2] << 16) | ((u32)p[pos + 3] << 24);
        pos += 4;
        return v;
    }
    u64 u64v() {
        u64 lo = u32v();
        u64 hi = u32v();
        return lo | (hi << 32);
    }
    QM31 qm() {
        u32 a = u32v(), b = u32v(), c = u32v(), d = u32v();
        return qm_make(a, b, c, d);
    }
    Hash32 hash() {
        need(32);
        Hash32 h;
        memcpy(h.b, p + pos, 32);
        pos += 32;
        return h;
    }
    MerkleDecommitment decommitment() {
        MerkleDecommitment d;
        u64 nh = u64v();
        for (u64 i = 0; i < nh; i++) d.hash_witness.push_back(hash());
        u64 nc = u64v();
        for (u64 i = 0; i < nc; i++) d.column_witness.push_back(u32v());
        return d;
    }
    FriLayerProof fri_layer() {
        FriLayerProof l;
        u64 nw = u64v();
        for (u64 i = 0; i < nw; i++) l.fri_witness.push_back(qm());
        l.decommitment = decommitment();
        l.commitment = hash();
        return l;
    }
    CommitmentSchemeProof proof() {
        CommitmentSchemeProof pr;
        pr.config.pow_bits = u32v();
        pr.config.fri_config.log_blowup_factor = u32v();
        pr.config.fri_config.log_last_layer_degree_bound = u32v();
        pr.config.fri_config.n_queries = (size_t)u64v();
        u64 nc = u64v();
        for (u64 i = 0; i < nc; i++) pr.commitments.push_back(hash());
        u64 nt = u64v();
        for (u64 t = 0; t < nt; t++) {
            pr.sampled_values.emplace_back();
            u64 ncol = u64v();
            for (u64 c = 0; c < ncol; c++) {
                pr.sampled_values.back().emplace_back();
                u64 ns = u64v();
                for (u64 k = 0; k < ns; k++) pr.sampled_values.back().back().push_back(qm());
            }
        }
        u64 nd = u64v();
        for (u64 i = 0; i < nd; i++) pr.decommitments.push_back(decommitment());
        u64 nq = u64v();
        for (u64 i = 0; i < nq; i++) {
            pr.queried_values.emplace_back();
            u64 k = u64v();
            for (u64 j = 0; j < k; j++) pr.queried_values.back().push_back(u32v());
        }
        pr.proof_of_work = u64v();
        pr.fri_proof.first_layer = fri_layer();
        u64 ni = u64v();
        for (u64 i = 0; i < ni; i++) pr.fri_proof.inner_layers.push_back(fri_layer());
        u64 nl = u64v();
        for (u64 i = 0; i < nl; i++) pr.fri_proof.last_layer_poly.push_back(qm());
        return pr;
    }
};

inline u32 ilog2(size_t n) {
    u32 l = 0;
    while (((size_t)1 << (l + 1)) <= n) l++;
    return l;
}

// ------------------------------------------------------------------ deferred gathers
// The reference reads queried rows with Column::at(i) one element at a time (vcs/prover.rs:125-140,
// fri.rs:1002-1036).  Which elements are read depends only on the query positions, never on the
// values, so every decommitment of a proof (4 trees + all FRI layers) first records its reads here
// and receives slot numbers; ONE batched device gather then serves them all.
template <class B>
struct GatherQueue {
    // a request = `count` consecutive words (1 for a column element, 8 for a hash node) of one source
    std::vector<const u32*> srcs;
    std::unordered_map<const u32*, u32> src_index;
    std::vector<u32> src_id, word, out_off, cnt;  // per run request: cnt[k] words land at results[out_off[k]..]
    const u32* results = nullptr;  // n_words gathered words (a page-locked buffer of the backend, or `store`)
    std::vector<u32> store;
    size_t n_words = 0;
    GatherQueue() {  // a proof records ~50k run requests: no regrowth on the critical path after the proof-of-work nonce
        for (auto* v : {&src_id, &word, &out_off, &cnt}) v->reserve(1u << 16);
        srcs.reserve(8192);
        src_index.reserve(8192);
    }
    u32 source(const u32* base) {
        auto it = src_index.find(base);
        if (it != src_index.end()) return it->second;
        u32 id = (u32)srcs.size();
        srcs.push_back(base);
        src_index[base] = id;
        return id;
    }
    size_t request_run(u32 source_id, size_t first_word, u32 count) {  // returns the slot of the first word
        src_id.push_back(source_id);
        word.push_back((u32)first_word);
        out_off.push_back((u32)n_words);
        cnt.push_back(count);
        size_t slot = n_words;
        n_words += count;
        return slot;
    }
    size_t request(u32 source_id, size_t word_idx) { return request_run(source_id, word_idx, 1); }
    size_t request_hash(u32 source_id, size_t node) { return request_run(source_id, node * 8, 8); }
    // A Merkle layer's decommitment reads the same rows of ALL its columns: one descriptor
    // (col_off, n_cols, row_off, n_rows, out_base) instead of n_cols * n_rows single-word requests.
    // Results are row-major: the word of column c at rows[k] lands in slot base + k * n_cols + c.
    std::vector<u32> grid_desc, grid_cols, grid_rows;
    size_t request_rows(const std::vector<u32>& col_ids, const std::vector<size_t>& rows) {
        size_t base = n_words;
        if (col_ids.empty() || rows.empty()) return base;
        for (u32 v : {(u32)grid_cols.size(), (u32)col_ids.size(), (u32)grid_rows.size(), (u32)rows.size(), (u32)base}) grid_desc.push_back(v);
        grid_cols.insert(grid_cols.end(), col_ids.begin(), col_ids.end());
        for (size_t r : rows) grid_rows.push_back((u32)r);
        n_words += col_ids.size() * rows.size();
        return base;
    }
    // flush = flush_async (the reads are enqueued) + wait (their results are in `results`)
    void flush_async() {
        if (n_words) results = B::gather_runs_async(srcs, src_id, word, out_off, cnt, grid_desc, grid_cols, grid_rows, n_words, store);
    }
    void wait() {
        if (n_words) B::gather_wait();
    }
    void flush() {
        flush_async();
        wait();
    }
    Hash32 hash_at(size_t slot) const {
        Hash32 h;
        memcpy(h.b, results + slot, 32);
        return h;
    }
};

// ------------------------------------------------------------------ MerkleProver (vcs/prover.rs)
template <class B>
struct MerkleProver {
    typedef typename B::Col Col;
    typedef typename B::HashCol HashCol;
    std::vector<HashCol> layers;  // layers[0] = root layer

    static MerkleProver commit(const std::vector<const Col*>& columns) {
        MerkleProver mp;
        if (columns.empty()) {
            mp.layers.push_back(B::commit_on_layer(0, nullptr, {}));
            return mp;
        }
        std::vector<const Col*> sorted = columns;
        std::stable_sort(sorted.begin(), sorted.end(), [](const Col* a, const Col* b) { return B::len(*a) > B::len(*b); });
        u32 max_log = ilog2(B::len(*sorted[0]));
        size_t pos = 0;
        std::vector<HashCol> layers;
        // Layers of <= 2^TOP_LOG nodes are hashed by ONE single-CTA launch (which can inject columns at every layer): 2^10 when
        // columns of 2^8..2^10 rows exist; otherwise 2^7, so the 2^8..2^10 layers stay multi-CTA inside the fused launch
        // above them (FRI layer trees, the composition tree).
        int TOP_LOG = 7;
        for (const Col* c : sorted) {
            u32 l = ilog2(B::len(*c));
            if (l > 7 && l <= 10) TOP_LOG = 10;
        }
        int log_size = (int)max_log;
        if (B::shard_world() > 1) {
            // Single-proof sharding: while a layer has at least 2^stripe_log nodes every rank hashes ITS node
            // range -- children from its own range of the layer below, columns through the peer mapping where another rank
            // owns them (the exchange is fused into the leaf kernel) -- and the layers stay distributed: decommitment reads
            // a node from the rank that holds it.  The last striped layer is completed by one small all-gather (the
            // "single all-gather at the Merkle root"); everything above it is hashed redundantly by every rank.
            const int stripe = (int)B::shard_stripe_log();
            bool any = false;
            while (log_size >= stripe && log_size > TOP_LOG) {
                std::vector<const Col*> layer_cols;
                while (pos < sorted.size() && ilog2(B::len(*sorted[pos])) == (u32)log_size) layer_cols.push_back(sorted[pos++]);
                // the column-free layers below are hashed by the same launch, as in the unsharded planner
                int next_with_cols = pos < sorted.size() ? (int)ilog2(B::len(*sorted[pos])) : -1;
                int n_levels = 1;
                while (log_size - n_levels >= stripe && log_size - n_levels > TOP_LOG && log_size - n_levels > next_with_cols && n_levels < 9) n_levels++;
                const HashCol* prev = layers.empty() ? nullptr : &layers.back();
                if (n_levels == 1) {
                    layers.push_back(B::commit_on_layer_striped((u32)log_size, prev, layer_cols));
                } else {
                    std::vector<HashCol> fused = B::commit_layers_fused_striped((u32)log_size, prev, layer_cols, (u32)n_levels);
                    for (auto& l : fused) layers.push_back(std::move(l));
                }
                log_size -= n_levels;
                any = true;
            }
            if (any) B::join_striped_layer(layers.back());
        }
        while (log_size > TOP_LOG) {
            std::vector<const Col*> layer_cols;
            while (pos < sorted.size() && ilog2(B::len(*sorted[pos])) == (u32)log_size) layer_cols.push_back(sorted[pos++]);
            // the layers below that receive no columns are hashed by the same launch (<= 9 levels: 256 nodes -> 1 per CTA)
            int next_with_cols = pos < sorted.size() ? (int)ilog2(B::len(*sorted[pos])) : -1;
            int n_levels = 1;
            while (log_size - n_levels > TOP_LOG && log_size - n_levels > next_with_cols && n_levels < 9) n_levels++;
            const HashCol* prev = layers.empty() ? nullptr : &layers.back();
            if (n_levels == 1) {
                layers.push_back(B::commit_on_layer((u32)log_size, prev, layer_cols));
            } else {
                std::vector<HashCol> fused = B::commit_layers_fused((u32)log_size, prev, layer_cols, (u32)n_levels);
                for (auto& l : fused) layers.push_back(std::move(l));
            }
            log_size -= n_levels;
        }
        {
            std::vector<std::vector<const Col*>> by_layer(log_size + 1);
            while (pos < sorted.size()) {
                by_layer[ilog2(B::len(*sorted[pos]))].push_back(sorted[pos]);
                pos++;
            }
            const HashCol* prev = layers.empty() ? nullptr : &layers.back();
            std::vector<HashCol> top = B::commit_top_layers((u32)log_size, prev, by_layer);
            for (int l = log_size; l >= 0; l--) layers.push_back(std::move(top[l]));
        }
        std::reverse(layers.begin(), layers.end());
        mp.layers = std::move(layers);
        return mp;
    }

    Hash32 root() const {  // one 32-byte device->host copy, once
        if (!root_read_) {
            root_ = B::read_root(layers[0]);
            root_read_ = true;
        }
        return root_;
    }
    mutable Hash32 root_;
    mutable bool root_read_ = false;

    // vcs/prover.rs:82-156 in two phases: `plan` walks the layers exactly like the reference and
    // records every read in the gather queue; `finish` (after queue.flush()) assembles the values.
    struct PendingDecommit {
        std::vector<std::pair<size_t, size_t>> queried_runs, column_witness_runs;  // (first slot, count)
        std::vector<size_t> hash_witness_slots;
        std::pair<std::vector<u32>, MerkleDecommitment> finish(const GatherQueue<B>& q) const {
            std::vector<u32> queried_values;
            MerkleDecommitment d;
            for (auto& r : queried_runs) queried_values.insert(queried_values.end(), q.results + r.first, q.results + r.first + r.second);
            for (auto& r : column_witness_runs)
                d.column_witness.insert(d.column_witness.end(), q.results + r.first, q.results + r.first + r.second);
            for (size_t s : hash_witness_slots) d.hash_witness.push_back(q.hash_at(s));
            return {queried_values, d};
        }
    };
    PendingDecommit plan_decommit(GatherQueue<B>& queue, const std::map<u32, std::vector<size_t>>& queries_per_log_size,
                                  const std::vector<const Col*>& columns) const {
        PendingDecommit pd;
        std::vector<const Col*> sorted = columns;
        std::stable_sort(sorted.begin(), sorted.end(), [](const Col* a, const Col* b) { return B::len(*a) > B::len(*b); });
        size_t pos = 0;
        std::vector<size_t> last_layer_queries;
        for (int layer_log_size = (int)layers.size() - 1; layer_log_size >= 0; layer_log_size--) {
            std::vector<const Col*> layer_columns;
            while (pos < sorted.size() && ilog2(B::len(*sorted[pos])) == (u32)layer_log_size) layer_columns.push_back(sorted[pos++]);
            const HashCol* previous_layer_hashes = (size_t)layer_log_size + 1 < layers.size() ? &layers[layer_log_size + 1] : nullptr;
            static const std::vector<size_t> empty;
            auto qit = queries_per_log_size.find((u32)layer_log_size);
            const std::vector<size_t>& layer_column_queries = qit == queries_per_log_size.end() ? empty : qit->second;
            size_t pi = 0, ci = 0;
            std::vector<size_t> layer_total_queries;
            std::vector<char> node_is_queried;
            std::vector<u32> col_ids;
            for (const Col* c : layer_columns) col_ids.push_back(queue.source(B::col_words(*c)));
            // (a striped layer of a sharded proof is read from the rank that holds the node: hash_node_source)
            const bool prev_striped = previous_layer_hashes && B::is_striped(*previous_layer_hashes);
            const u32 prev_id_whole = previous_layer_hashes && !prev_striped ? queue.source(B::hash_words(*previous_layer_hashes)) : 0;
            auto prev_id_of = [&](size_t node) {
                return prev_striped ? queue.source(B::hash_node_source(*previous_layer_hashes, node)) : prev_id_whole;
            };
            while (pi < last_layer_queries.size() || ci < layer_column_queries.size()) {
                size_t node_index;
                bool has_p = pi < last_layer_queries.size(), has_c = ci < layer_column_queries.size();
                if (has_p && has_c) node_index = std::min(last_layer_queries[pi] / 2, layer_column_queries[ci]);
                else if (has_p) node_index = last_layer_queries[pi] / 2;
                else node_index = layer_column_queries[ci];
                if (previous_layer_hashes) {
                    if (pi < last_layer_queries.size() && last_layer_queries[pi] == 2 * node_index) pi++;
                    else pd.hash_witness_slots.push_back(queue.request_hash(prev_id_of(2 * node_index), 2 * node_index));
                    if (pi < last_layer_queries.size() && last_layer_queries[pi] == 2 * node_index + 1) pi++;
                    else pd.hash_witness_slots.push_back(queue.request_hash(prev_id_of(2 * node_index + 1), 2 * node_index + 1));
                }
                bool queried = ci < layer_column_queries.size() && layer_column_queries[ci] == node_index;
                if (queried) ci++;
                node_is_queried.push_back(queried);
                layer_total_queries.push_back(node_index);
            }
            if (!col_ids.empty()) {  // every column of the layer at every visited node: one row-grid request
                size_t base = queue.request_rows(col_ids, layer_total_queries);
                for (size_t k = 0; k < layer_total_queries.size(); k++)
                    (node_is_queried[k] ? pd.queried_runs : pd.column_witness_runs).push_back({base + k * col_ids.size(), col_ids.size()});
            }
            last_layer_queries = layer_total_queries;
        }
        return pd;
    }
    std::pair<std::vector<u32>, MerkleDecommitment> decommit(const std::map<u32, std::vector<size_t>>& queries_per_log_size,
                                                             const std::vector<const Col*>& columns) const {
        GatherQueue<B> q;
        PendingDecommit pd = plan_decommit(q, queries_per_log_size, columns);
        q.flush();
        return pd.finish(q);
    }
};

// ------------------------------------------------------------------ polynomials / evaluations
template <class B>
struct CirclePoly {
    typename B::Col coeffs;
    u32 log_size;
};
template <class B>
struct CircleEvaluation {  // values on CanonicCoset(log_size).circle_domain(), bit-reversed order
    typename B::Col values;
    u32 log_size;
};
template <class B>
struct SecureEvaluation {
    std::array<typename B::Col, 4> columns;
    u32 log_size;
};

struct PointSample {
    SecurePoint point;
    QM31 value;
};
struct ColumnSampleBatch {  // pcs/quotients.rs:39-70
    SecurePoint point;
    std::vector<std::pair<size_t, QM31>> columns_and_values;
    static std::vector<ColumnSampleBatch> new_vec(const std::vector<const std::vector<PointSample>*>& samples) {
        std::vector<ColumnSampleBatch> out;  // insertion order = first appearance of the point (IndexMap)
        for (size_t column_index = 0; column_index < samples.size(); column_index++) {
            for (const PointSample& s : *samples[column_index]) {
                size_t k = 0;
                for (; k < out.size(); k++)
                    if (cpq_eq(out[k].point, s.point)) break;
                if (k == out.size()) out.push_back(ColumnSampleBatch{s.point, {}});
                out[k].columns_and_values.push_back({column_index, s.value});
            }
        }
        return out;
    }
};

// ------------------------------------------------------------------ CommitmentTreeProver / CommitmentSchemeProver
template <class B>
struct CommitmentTreeProver {
    std::vector<CirclePoly<B>> polynomials;
    std::vector<CircleEvaluation<B>> evaluations;
    MerkleProver<B> commitment;

    static CommitmentTreeProver create(std::vector<CirclePoly<B>> polynomials, u32 log_blowup_factor, Blake2sChannel& channel,
                                       const typename B::Twiddles& twiddles) {
        CommitmentTreeProver t;
        t.polynomials = std::move(polynomials);
        t.evaluations.resize(t.polynomials.size());
        // B::evaluate_polynomials, batched per log size (ops.rs:51-65)
        std::map<u32, std::vector<size_t>> by_size;
        for (size_t i = 0; i < t.polynomials.size(); i++) by_size[t.polynomials[i].log_size].push_back(i);
        Span<B> commitment_span("Commitment");
        std::unique_ptr<Span<B>> phase(new Span<B>("Extension"));
        for (auto& kv : by_size) {
            u32 log_size = kv.first, log_eval = log_size + log_blowup_factor;
            B::lane(log_size);  // size groups are independent until the Merkle tree reads them all
            std::vector<const typename B::Col*> src;
            std::vector<typename B::Col*> dst;
            std::vector<typename B::Col> slab = B::uninit_many(kv.second.size(), (size_t)1 << log_eval);
            size_t si = 0;
            for (size_t i : kv.second) {
                t.evaluations[i].values = std::move(slab[si++]);
                t.evaluations[i].log_size = log_eval;
                src.push_back(&t.polynomials[i].coeffs);
                dst.push_back(&t.evaluations[i].values);
            }
            B::evaluate_polynomials(src, dst, log_size, log_eval, twiddles);
        }
        B::lanes_join();
        B::shard_barrier();  // sharded proof: every rank's LDE columns are complete before any rank hashes rows across them
        std::vector<const typename B::Col*> cols;
        for (auto& e : t.evaluations) cols.push_back(&e.values);
        phase.reset();
        phase.reset(new Span<B>("Merkle"));
        t.commitment = MerkleProver<B>::commit(cols);
        channel.mix_root(t.commitment.root());
        return t;
    }
    std::pair<std::vector<u32>, MerkleDecommitment> decommit(const std::map<u32, std::vector<size_t>>& queries) const {
        std::vector<const typename B::Col*> cols;
        for (auto& e : evaluations) cols.push_back(&e.values);
        return commitment.decommit(queries, cols);
    }
    typename MerkleProver<B>::PendingDecommit plan_decommit(GatherQueue<B>& queue, const std::map<u32, std::vector<size_t>>& queries) const {
        std::vector<const typename B::Col*> cols;
        for (auto& e : evaluations) cols.push_back(&e.values);
        return commitment.plan_decommit(queue, queries, cols);
    }
};

template <class B>
struct FriProver;

template <class B>
struct CommitmentSchemeProver {
    std::vector<CommitmentTreeProver<B>> trees;
    PcsConfig config;
    const typename B::Twiddles* twiddles;

    CommitmentSchemeProver(PcsConfig cfg, const typename B::Twiddles* tw) : config(cfg), twiddles(tw) {}

    // TreeBuilder::extend_evals + commit (pcs/prover.rs:173-203): columns are consumed.
    void commit_evals(std::vector<CircleEvaluation<B>> columns, Blake2sChannel& channel) {
        if (B::shard_world() > 1) {  // sharded proof: out-of-place, so the column transforms can be dealt out over the ranks
            std::vector<const CircleEvaluation<B>*> refs;
            for (auto& c : columns) refs.push_back(&c);
            commit_evals_keep(refs, channel);
            return;
        }
        std::map<u32, std::vector<typename B::Col*>> by_size;
        for (auto& c : columns) by_size[c.log_size].push_back(&c.values);
        {
            Span<B> span("Interpolation for commitment");
            for (auto& kv : by_size) {
                B::lane(kv.first);
                B::interpolate_columns(kv.second, kv.first, *twiddles);
            }
        }
        B::lane(0xffffffffu);  // back to lane 0 (no join: each size stays on its lane through the LDE)
        std::vector<CirclePoly<B>> polys;
        for (auto& c : columns) polys.push_back(CirclePoly<B>{std::move(c.values), c.log_size});
        commit_polys(std::move(polys), channel);
    }
    // Same commitment, but the evaluations are only borrowed: coefficients go to fresh columns.
    void commit_evals_keep(const std::vector<const CircleEvaluation<B>*>& columns, Blake2sChannel& channel) {
        B::shard_barrier();  // sharded proof: the values are complete on their owners before other ranks transform them
        std::vector<CirclePoly<B>> polys(columns.size());
        std::map<u32, std::vector<size_t>> idx_by_size;
        for (size_t i = 0; i < columns.size(); i++) idx_by_size[columns[i]->log_size].push_back(i);
        std::map<u32, std::pair<std::vector<const typename B::Col*>, std::vector<typename B::Col*>>> by_size;
        for (auto& kv : idx_by_size) {
            std::vector<typename B::Col> slab = B::uninit_many(kv.second.size(), (size_t)1 << kv.first);
            size_t si = 0;
            for (size_t i : kv.second) {
                polys[i].coeffs = std::move(slab[si++]);
                polys[i].log_size = kv.first;
                by_size[kv.first].first.push_back(&columns[i]->values);
                by_size[kv.first].second.push_back(&polys[i].coeffs);
            }
        }
        {
            Span<B> span("Interpolation for commitment");
            for (auto& kv : by_size) {
                B::lane(kv.first);
                B::interpolate_columns_to(kv.second.first, kv.second.second, kv.first, *twiddles);
            }
        }
        B::lane(0xffffffffu);
        commit_polys(std::move(polys), channel);
    }
    void commit_polys(std::vector<CirclePoly<B>> polys, Blake2sChannel& channel) {
        trees.push_back(CommitmentTreeProver<B>::create(std::move(polys), config.fri_config.log_blowup_factor, channel, *twiddles));
    }
    std::vector<Hash32> roots() const {
        std::vector<Hash32> r;
        for (auto& t : trees) r.push_back(t.commitment.root());
        return r;
    }

    // prove_values (pcs/prover.rs:83-153)
    // `after_sampling` (optional) runs on the host once the sampled values are known and the DEEP quotient kernels are
    // enqueued: prove() puts its OODS sanity check there so it overlaps device work instead of trailing the proof.
    typedef std::function<void(const std::vector<std::vector<std::vector<QM31>>>&)> SampledHook;
    CommitmentSchemeProof prove_values(const std::vector<std::vector<std::vector<SecurePoint>>>& sampled_points, Blake2sChannel& channel,
                                       const SampledHook& after_sampling = SampledHook());
};

// ------------------------------------------------------------------ FRI (core/fri.rs)
template <class B>
struct FriProver {
    typedef typename B::Col Col;
    struct InnerLayer {
        std::array<Col, 4> evaluation;
        u32 log_size;
        MerkleProver<B> merkle_tree;
        Hash32 root;
    };
    FriConfig config;
    const std::vector<SecureEvaluation<B>>* columns = nullptr;
    MerkleProver<B> first_layer_tree;
    Hash32 first_layer_root;
    std::vector<InnerLayer> inner_layers;
    std::vector<QM31> last_layer_poly;

    static std::vector<const Col*> coordinate_columns(const std::vector<SecureEvaluation<B>>& columns) {
        std::vector<const Col*> out;
        for (auto& sc : columns)
            for (auto& c : sc.columns) out.push_back(&c);
        return out;
    }

    static FriProver commit(Blake2sChannel& channel, FriConfig config, const std::vector<SecureEvaluation<B>>& columns,
                            const typename B::Twiddles& twiddles) {
        if (columns.empty()) throw std::logic_error("no columns");
        for (size_t i = 0; i + 1 < columns.size(); i++)
            if (!(columns[i].log_size > columns[i + 1].log_size)) throw std::logic_error("column sizes not decreasing");
        FriProver fp;
        fp.config = config;
        fp.columns = &columns;
        // commit_first_layer
        fp.first_layer_tree = MerkleProver<B>::commit(coordinate_columns(columns));
        fp.first_layer_root = fp.first_layer_tree.root();
        channel.mix_root(fp.first_layer_root);
        // commit_inner_layers
        u32 first_inner_log = columns[0].log_size - 1;
        std::array<Col, 4> layer_eval;
        for (auto& c : layer_eval) c = B::zeros((size_t)1 << first_inner_log);
        u32 layer_log = first_inner_log;
        size_t next_col = 0;
        QM31 folding_alpha = channel.draw_secure_felt();
        B::fold_circle_into_line(layer_eval, columns[next_col].columns, columns[next_col].log_size, folding_alpha, twiddles);
        next_col++;
        while (((size_t)1 << layer_log) > config.last_layer_domain_size()) {
            if (layer_log <= B::fri_tail_log()) {
                // the remaining (small) layers: one launch for the whole chain, Fiat-Shamir included; the roots come back and
                // the channel is replayed from them (same transcript)
                B::template fri_tail<InnerLayer>(channel, layer_eval, layer_log, ilog2(config.last_layer_domain_size()), columns, next_col,
                                                 twiddles, fp.inner_layers);
                break;
            }
            InnerLayer layer;
            std::vector<const Col*> cc;
            for (auto& c : layer_eval) cc.push_back(&c);
            layer.merkle_tree = MerkleProver<B>::commit(cc);
            layer.root = layer.merkle_tree.root();
            channel.mix_root(layer.root);
            QM31 alpha = channel.draw_secure_felt();
            std::array<Col, 4> folded = B::fold_line(layer_eval, layer_log, alpha, twiddles);
            layer.evaluation = std::move(layer_eval);
            layer.log_size = layer_log;
            layer_eval = std::move(folded);
            layer_log -= 1;
            if (next_col < columns.size() && columns[next_col].log_size - 1 == layer_log) {
                B::fold_circle_into_line(layer_eval, columns[next_col].columns, columns[next_col].log_size, alpha, twiddles);
                next_col++;
            }
            fp.inner_layers.push_back(std::move(layer));
        }
        if (next_col != columns.size()) throw std::logic_error("not all columns were folded");
        // commit_last_layer (fri.rs:268-290): interpolate the line evaluation on the host
        size_t n_last = (size_t)1 << layer_log;
        if (n_last != config.last_layer_domain_size()) throw std::logic_error("bad last layer size");
        std::vector<QM31> values(n_last);
        {
            std::vector<u32> h[4];
            for (int k = 0; k < 4; k++) {
                h[k].resize(n_last);
                B::to_host(layer_eval[k], h[k].data());
            }
            for (size_t i = 0; i < n_last; i++) values[i] = qm_make(h[0][i], h[1][i], h[2][i], h[3][i]);
        }
        fp.last_layer_poly = interpolate_last_layer(values, layer_log, config);
        channel.mix_felts(fp.last_layer_poly);
        return fp;
    }

    // LineEvaluation::interpolate + into_ordered_coefficients/from_ordered_coefficients (poly/line.rs)
    static std::vector<QM31> interpolate_last_layer(std::vector<QM31> values, u32 log_size, const FriConfig& config) {
        size_t n = values.size();
        // bit_reverse_column
        for (size_t i = 0; i < n; i++) {
            size_t j = bit_reverse((u32)i, log_size);
            if (i < j) std::swap(values[i], values[j]);
        }
        // line_ifft on LineDomain(half_odds(log_size))
        Coset dom = Coset::half_odds(log_size);
        size_t dsize = n;
        while (dsize > 1) {
            for (size_t chunk = 0; chunk < n; chunk += dsize) {
                for (size_t i = 0; i < dsize / 2; i++) {
                    u32 x = dom.at(i).x;
                    u32 xi = m31_inv(x);
                    QM31& l = values[chunk + i];
                    QM31& r = values[chunk + dsize / 2 + i];
                    QM31 tmp = l;
                    l = tmp + r;
                    r = qm_mul_m31(tmp - r, xi);
                }
            }
            dom = dom.doubled();
            dsize /= 2;
        }
        u32 len_inv = m31_inv((u32)n);
        for (auto& v : values) v = qm_mul_m31(v, len_inv);
        // into_ordered_coefficients: bit reverse
        for (size_t i = 0; i < n; i++) {
            size_t j = bit_reverse((u32)i, log_size);
            if (i < j) std::swap(values[i], values[j]);
        }
        size_t bound = (size_t)1 << config.log_last_layer_degree_bound;
        for (size_t i = bound; i < n; i++)
            if (!qm_is_zero(values[i]) && !getenv("CM31_DEBUG_SKIP_OODS_CHECK")) throw std::logic_error("invalid degree");
        values.resize(bound);
        // from_ordered_coefficients: bit reverse again (over `bound` elements)
        u32 lb = config.log_last_layer_degree_bound;
        for (size_t i = 0; i < bound; i++) {
            size_t j = bit_reverse((u32)i, lb);
            if (i < j) std::swap(values[i], values[j]);
        }
        return values;
    }

    // compute_decommitment_positions_and_witness_evals (fri.rs:1002-1036); reads are queued
    static void positions_and_witness(GatherQueue<B>& queue, const std::array<Col, 4>& column, const std::vector<size_t>& query_positions,
                                      u32 fold_step, std::vector<size_t>& decommitment_positions, std::vector<size_t>& witness_slots) {
        u32 ids[4];
        for (int k = 0; k < 4; k++) ids[k] = queue.source(B::col_words(column[k]));
        size_t i = 0;
        while (i < query_positions.size()) {
            size_t j = i;
            while (j < query_positions.size() && (query_positions[j] >> fold_step) == (query_positions[i] >> fold_step)) j++;
            size_t subset_start = (query_positions[i] >> fold_step) << fold_step;
            size_t qi = i;
            for (size_t position = subset_start; position < subset_start + ((size_t)1 << fold_step); position++) {
                decommitment_positions.push_back(position);
                if (qi < j && query_positions[qi] == position) {
                    qi++;
                    continue;
                }
                size_t first = queue.request(ids[0], position);
                for (int k = 1; k < 4; k++) queue.request(ids[k], position);
                witness_slots.push_back(first);  // 4 consecutive slots = one QM31
            }
            i = j;
        }
    }

    struct PendingLayer {
        std::vector<size_t> witness_slots;
        typename MerkleProver<B>::PendingDecommit decommit;
        Hash32 commitment;
        FriLayerProof finish(const GatherQueue<B>& q) const {
            FriLayerProof lp;
            for (size_t s : witness_slots) lp.fri_witness.push_back(qm_make(q.results[s], q.results[s + 1], q.results[s + 2], q.results[s + 3]));
            lp.decommitment = decommit.finish(q).second;
            lp.commitment = commitment;
            return lp;
        }
    };
    struct PendingFriProof {
        PendingLayer first_layer;
        std::vector<PendingLayer> inner_layers;
        std::vector<QM31> last_layer_poly;
        FriProof finish(const GatherQueue<B>& q) const {
            FriProof proof;
            proof.first_layer = first_layer.finish(q);
            for (auto& l : inner_layers) proof.inner_layers.push_back(l.finish(q));
            proof.last_layer_poly = last_layer_poly;
            return proof;
        }
    };

    // FriProver::decommit (fri.rs:295-330) with every read deferred to `queue`
    std::pair<PendingFriProof, std::map<u32, std::vector<size_t>>> plan_decommit(Blake2sChannel& channel, GatherQueue<B>& queue) {
        std::set<u32> column_log_sizes;
        for (auto& c : *columns) column_log_sizes.insert(c.log_size);
        u32 max_column_log_size = *column_log_sizes.rbegin();
        Queries queries = Queries::generate(channel, max_column_log_size, config.n_queries);
        std::map<u32, std::vector<size_t>> query_positions_by_log_size;
        for (u32 ls : column_log_sizes) query_positions_by_log_size[ls] = queries.fold(queries.log_domain_size - ls).positions;
        PendingFriProof proof;
        // first layer (fri.rs:906-945)
        {
            std::map<u32, std::vector<size_t>> decommitment_positions_by_log_size;
            for (auto& column : *columns) {
                Queries cq = queries.fold(queries.log_domain_size - column.log_size);
                std::vector<size_t> positions;
                positions_and_witness(queue, column.columns, cq.positions, 1, positions, proof.first_layer.witness_slots);
                decommitment_positions_by_log_size[column.log_size] = positions;
            }
            proof.first_layer.decommit = first_layer_tree.plan_decommit(queue, decommitment_positions_by_log_size, coordinate_columns(*columns));
            proof.first_layer.commitment = first_layer_root;
        }
        Queries layer_queries = queries.fold(1);
        for (auto& layer : inner_layers) {
            PendingLayer lp;
            std::vector<size_t> positions;
            positions_and_witness(queue, layer.evaluation, layer_queries.positions, 1, positions, lp.witness_slots);
            std::map<u32, std::vector<size_t>> m;
            m[layer.log_size] = positions;
            std::vector<const Col*> cc;
            for (auto& c : layer.evaluation) cc.push_back(&c);
            lp.decommit = layer.merkle_tree.plan_decommit(queue, m, cc);
            lp.commitment = layer.root;
            proof.inner_layers.push_back(std::move(lp));
            layer_queries = layer_queries.fold(1);
        }
        proof.last_layer_poly = last_layer_poly;
        return {proof, query_positions_by_log_size};
    }
};

// compute_fri_quotients (pcs/quotients.rs:77-102)
template <class B>
std::vector<SecureEvaluation<B>> compute_fri_quotients(const std::vector<const CircleEvaluation<B>*>& columns,
                                                       const std::vector<std::vector<PointSample>>& samples, QM31 random_coeff,
                                                       u32 log_blowup_factor) {
    std::vector<size_t> order(columns.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return columns[a]->log_size > columns[b]->log_size; });
    std::vector<SecureEvaluation<B>> out;
    size_t i = 0;
    while (i < order.size()) {
        u32 log_size = columns[order[i]]->log_size;
        B::lane(log_size > log_blowup_factor ? log_size - log_blowup_factor : 0);  // lane of the trace size this LDE belongs to
        std::vector<const typename B::Col*> cols;
        std::vector<const std::vector<PointSample>*> smp;
        while (i < order.size() && columns[order[i]]->log_size == log_size) {
            cols.push_back(&columns[order[i]]->values);
            smp.push_back(&samples[order[i]]);
            i++;
        }
        std::vector<ColumnSampleBatch> batches = ColumnSampleBatch::new_vec(smp);
        SecureEvaluation<B> ev;
        ev.log_size = log_size;
        ev.columns = B::accumulate_quotients(log_size, cols, random_coeff, batches, log_blowup_factor);
        out.push_back(std::move(ev));
    }
    B::lanes_join();
    return out;
}

template <class B>
CommitmentSchemeProof CommitmentSchemeProver<B>::prove_values(const std::vector<std::vector<std::vector<SecurePoint>>>& sampled_points,
                                                              Blake2sChannel& channel, const SampledHook& after_sampling) {
    // Evaluate polynomials on open points: one batched eval_at_point over every (column, point).
    std::unique_ptr<HostTimer> ht(new HostTimer("pv_prep"));
    std::vector<const typename B::Col*> polys;
    std::vector<u32> log_sizes, point_idx;
    std::vector<SecurePoint> points;
    for (size_t t = 0; t < trees.size(); t++) {
        if (sampled_points[t].size() != trees[t].polynomials.size()) throw std::logic_error("sampled_points shape mismatch");
        for (size_t c = 0; c < trees[t].polynomials.size(); c++) {
            for (const SecurePoint& p : sampled_points[t][c]) {
                size_t k = 0;
                for (; k < points.size(); k++)
                    if (cpq_eq(points[k], p)) break;
                if (k == points.size()) points.push_back(p);
                polys.push_back(&trees[t].polynomials[c].coeffs);
                log_sizes.push_back(trees[t].polynomials[c].log_size);
                point_idx.push_back((u32)k);
            }
        }
    }
    std::vector<QM31> values;
    ht.reset(new HostTimer("pv_eval_at_points"));
    {
        Span<B> span("Evaluate columns out of domain");
        B::eval_at_points(polys, log_sizes, points, point_idx, values);
    }
    ht.reset(new HostTimer("pv_assemble"));
    CommitmentSchemeProof proof;
    proof.config = config;
    std::vector<std::vector<PointSample>> samples_flat;
    std::vector<QM31> sampled_flat;
    size_t vi = 0;
    size_t n_cols_total = 0;
    for (auto& t : trees) n_cols_total += t.polynomials.size();
    samples_flat.reserve(n_cols_total);
    sampled_flat.reserve(values.size());
    for (size_t t = 0; t < trees.size(); t++) {
        for (size_t c = 0; c < trees[t].polynomials.size(); c++) {
            samples_flat.emplace_back();
            std::vector<PointSample>& col_samples = samples_flat.back();
            col_samples.reserve(sampled_points[t][c].size());
            for (const SecurePoint& p : sampled_points[t][c]) {
                col_samples.push_back(PointSample{p, values[vi]});
                sampled_flat.push_back(values[vi]);
                vi++;
            }
        }
    }
    channel.mix_felts(sampled_flat);

    std::vector<const CircleEvaluation<B>*> columns;
    for (auto& t : trees)
        for (auto& e : t.evaluations) columns.push_back(&e);
    QM31 random_coeff = channel.draw_secure_felt();
    ht.reset(new HostTimer("pv_quotients"));
    std::unique_ptr<Span<B>> qspan(new Span<B>("Compute FRI quotients"));
    std::vector<SecureEvaluation<B>> quotients = compute_fri_quotients<B>(columns, samples_flat, random_coeff, config.fri_config.log_blowup_factor);
    qspan.reset();
    {  // proof.sampled_values ([tree][column][sample]): assembled while the quotient kernels run
        size_t v2 = 0;
        for (size_t t = 0; t < trees.size(); t++) {
            proof.sampled_values.emplace_back();
            proof.sampled_values.back().reserve(trees[t].polynomials.size());
            for (size_t c = 0; c < trees[t].polynomials.size(); c++) {
                const size_t k = sampled_points[t][c].size();
                proof.sampled_values.back().emplace_back(values.begin() + v2, values.begin() + v2 + k);
                v2 += k;
            }
        }
    }

    if (after_sampling) {
        ht.reset(new HostTimer("pv_after_sampling_hook"));
        after_sampling(proof.sampled_values);
    }
    ht.reset(new HostTimer("pv_fri_commit"));
    std::unique_ptr<Span<B>> fspan(new Span<B>("FRI commitment"));
    FriProver<B> fri_prover = FriProver<B>::commit(channel, config.fri_config, quotients, *twiddles);
    fspan.reset();

    ht.reset(new HostTimer("pv_grind"));
    {
        Span<B> span("Grind");
        proof.proof_of_work = B::grind(channel.digest(), config.pow_bits);
    }
    channel.mix_u64(proof.proof_of_work);
    proof.commitments = roots();  // (cached at commit time: no device read)

    // FRI + the 4 trees decommit through one batched device gather
    struct Tail {
        // deferred form only: the prover state the decommitment reads, moved out of this call
        std::vector<CommitmentTreeProver<B>> trees;
        std::vector<SecureEvaluation<B>> quotients;
        std::unique_ptr<FriProver<B>> fri;
        Blake2sChannel channel;
        // both forms
        typename FriProver<B>::PendingFriProof fri_pending;
        std::vector<typename MerkleProver<B>::PendingDecommit> pending;
        GatherQueue<B> queue;
        int stage = 0;
    };
    auto plan_and_gather = [](std::vector<CommitmentTreeProver<B>>& tr, FriProver<B>& fp, Blake2sChannel& ch, Tail& st) {
        std::unique_ptr<HostTimer> h(new HostTimer("pv_decommit_plan"));
        auto fri_res = fp.plan_decommit(ch, st.queue);
        st.fri_pending = std::move(fri_res.first);
        h.reset(new HostTimer("pv_decommit_plan_trees"));
        for (auto& t : tr) st.pending.push_back(t.plan_decommit(st.queue, fri_res.second));
        h.reset(new HostTimer("pv_decommit_flush"));
        st.queue.flush_async();
    };
    auto finish = [](CommitmentSchemeProof& p, Tail& st) {
        HostTimer ht2("pv_decommit_finish");
        st.queue.wait();
        p.fri_proof = st.fri_pending.finish(st.queue);
        for (auto& pd : st.pending) {
            auto res = pd.finish(st.queue);
            p.queried_values.push_back(std::move(res.first));
            p.decommitments.push_back(std::move(res.second));
        }
    };
    ht.reset();
    B::finish_deferred_tails();  // a previous proof's tail must be complete before this one takes over the landing buffer / the hook
    auto st = std::make_shared<Tail>();
    if (B::defer_proof_tail()) {
        st->trees = std::move(trees);
        st->quotients = std::move(quotients);
        st->fri.reset(new FriProver<B>(std::move(fri_prover)));
        st->fri->columns = &st->quotients;
        st->channel = channel;
        proof.pending_tail = std::make_shared<std::function<bool(CommitmentSchemeProof&)>>([st, plan_and_gather, finish](CommitmentSchemeProof& p) {
            if (st->stage == 0) {
                plan_and_gather(st->trees, *st->fri, st->channel, *st);
                // the gather is enqueued: the device buffers of the proof go back to the pool (stream-ordered after it)
                st->fri.reset();
                st->quotients.clear();
                st->trees.clear();
                st->stage = 1;
                return false;
            }
            finish(p, *st);
            return true;
        });
    } else {
        plan_and_gather(trees, fri_prover, channel, *st);
        finish(proof, *st);
    }
    return proof;
}

// ------------------------------------------------------------------ components & prove()
// DomainEvaluationAccumulator (air/accumulation.rs:49-154)
template <class B>
struct DomainEvaluationAccumulator {
    std::vector<QM31> random_coeff_powers;
    std::vector<std::unique_ptr<std::array<typename B::Col, 4>>> sub_accumulations;  // index = log size

    DomainEvaluationAccumulator(QM31 random_coeff, u32 max_log_size, size_t total_columns) {
        random_coeff_powers = B::generate_secure_powers(random_coeff, total_columns);
        sub_accumulations.resize(max_log_size + 1);
    }
    u32 log_size() const { return (u32)sub_accumulations.size() - 1; }
    // columns([(log_size, n_cols)]): hands out the LAST n_cols powers (accumulation.rs:75-96)
    std::pair<std::vector<QM31>, std::array<typename B::Col, 4>*> columns(u32 log_size, size_t n_cols) {
        if (n_cols > random_coeff_powers.size()) throw std::logic_error("not enough random coefficient powers");
        std::vector<QM31> coeffs(random_coeff_powers.end() - n_cols, random_coeff_powers.end());
        random_coeff_powers.resize(random_coeff_powers.size() - n_cols);
        auto& slot = sub_accumulations.at(log_size);
        if (!slot) {
            slot.reset(new std::array<typename B::Col, 4>());
            for (auto& c : *slot) {
                c = B::zeros((size_t)1 << log_size);
                B::set_owner(c, -1);  // every rank holds an accumulator of its own components' terms
            }
        }
        return {coeffs, slot.get()};
    }
    // finalize (accumulation.rs:104-153): per size ascending, interpolate and lift the previous poly
    std::array<CirclePoly<B>, 4> finalize(const typename B::Twiddles& twiddles) {
        if (!random_coeff_powers.empty()) throw std::logic_error("not all random coefficients were used");
        u32 max_log = log_size();
        bool have = false;
        std::array<CirclePoly<B>, 4> cur;
        for (u32 ls = 1; ls <= max_log; ls++) {
            if (!sub_accumulations[ls]) continue;
            std::array<typename B::Col, 4>& values = *sub_accumulations[ls];
            if (have) {
                std::array<typename B::Col, 4> lifted;
                std::vector<const typename B::Col*> src;
                std::vector<typename B::Col*> dst;
                for (int k = 0; k < 4; k++) {
                    lifted[k] = B::uninit((size_t)1 << ls);
                    src.push_back(&cur[k].coeffs);
                    dst.push_back(&lifted[k]);
                }
                B::evaluate_polynomials(src, dst, cur[0].log_size, ls, twiddles);
                B::accumulate(values, lifted);
            }
            std::vector<typename B::Col*> cols;
            for (int k = 0; k < 4; k++) cols.push_back(&values[k]);
            B::interpolate_columns(cols, ls, twiddles);
            for (int k = 0; k < 4; k++) cur[k] = CirclePoly<B>{std::move(values[k]), ls};
            have = true;
        }
        if (!have)
            for (int k = 0; k < 4; k++) cur[k] = CirclePoly<B>{B::zeros((size_t)1 << max_log), max_log};
        return cur;
    }
};

struct PointEvaluationAccumulator {  // accumulation.rs:18-44
    QM31 random_coeff, accumulation;
    explicit PointEvaluationAccumulator(QM31 rc) : random_coeff(rc), accumulation(qm_zero()) {}
    void accumulate(QM31 evaluation) { accumulation = accumulation * random_coeff + evaluation; }
};

template <class B>
struct Trace {
    const std::vector<CommitmentTreeProver<B>>* trees;
};

typedef std::vector<std::vector<std::vector<SecurePoint>>> MaskPoints;   // [tree][column][point]
typedef std::vector<std::vector<std::vector<QM31>>> MaskValues;

// Component + ComponentProver (air/mod.rs:26-67)
template <class B>
struct ComponentProver {
    virtual ~ComponentProver() {}
    virtual size_t n_constraints() const = 0;
    virtual u32 max_constraint_log_degree_bound() const = 0;
    virtual std::vector<std::vector<u32>> trace_log_degree_bounds() const = 0;
    virtual MaskPoints mask_points(SecurePoint point) const = 0;
    // Overwrites the points of this component's columns inside an already SHAPED MaskPoints (trees 1.., `at[t]` = its first
    // column in tree t, advanced past its columns): same values as mask_points(point), no allocation.
    virtual void fill_mask_points(SecurePoint point, MaskPoints& out, std::vector<size_t>& at) const {
        MaskPoints mp = mask_points(point);
        for (size_t t = 1; t < mp.size(); t++)
            for (auto& col : mp[t]) out[t][at[t]++] = col;
    }
    virtual std::vector<size_t> preprocessed_column_indices() const = 0;
    virtual void evaluate_constraint_quotients_at_point(SecurePoint point, const MaskValues& mask, PointEvaluationAccumulator& acc) const = 0;
    virtual void evaluate_constraint_quotients_on_domain(const Trace<B>& trace, DomainEvaluationAccumulator<B>& acc) const = 0;
};

template <class B>
struct ComponentProvers {  // air/components.rs
    std::vector<const ComponentProver<B>*> components;
    size_t n_preprocessed_columns;

    u32 composition_log_degree_bound() const {
        u32 m = 0;
        for (auto* c : components) m = std::max(m, c->max_constraint_log_degree_bound());
        return m;
    }
    MaskPoints mask_points(SecurePoint point) const {
        MaskPoints out;
        for (auto* c : components) {
            MaskPoints mp = c->mask_points(point);
            if (out.size() < mp.size()) out.resize(mp.size());
            for (size_t t = 0; t < mp.size(); t++)
                out[t].insert(out[t].end(), std::make_move_iterator(mp[t].begin()), std::make_move_iterator(mp[t].end()));
        }
        out[PREPROCESSED_TRACE_IDX_()].assign(n_preprocessed_columns, {});
        for (auto* c : components)
            for (size_t idx : c->preprocessed_column_indices()) out[PREPROCESSED_TRACE_IDX_()][idx] = {point};
        return out;
    }
    // the same result written into `out`, which has the shape of an earlier mask_points(..) call (made with any point while
    // the GPU was busy): on the critical path between the composition root and the OODS launch only values are written
    void fill_mask_points(SecurePoint point, MaskPoints& out) const {
        std::vector<size_t> at(out.size(), 0);
        for (auto* c : components) c->fill_mask_points(point, out, at);
        for (auto* c : components)
            for (size_t idx : c->preprocessed_column_indices()) out[PREPROCESSED_TRACE_IDX_()][idx][0] = point;
    }
    QM31 eval_composition_polynomial_at_point(SecurePoint point, const MaskValues& mask_values, QM31 random_coeff) const {
        PointEvaluationAccumulator acc(random_coeff);
        for (auto* c : components) c->evaluate_constraint_quotients_at_point(point, mask_values, acc);
        return acc.accumulation;
    }
    std::array<CirclePoly<B>, 4> compute_composition_polynomial(QM31 random_coeff, const Trace<B>& trace, const typename B::Twiddles& tw) const {
        size_t total = 0;
        for (auto* c : components) total += c->n_constraints();
        DomainEvaluationAccumulator<B> acc(random_coeff, composition_log_degree_bound(), total);
        // accumulators are per evaluation size and a component's lane is a function of its size: no two lanes share one
        size_t ci = 0;
        for (auto* c : components) {
            B::lane(c->max_constraint_log_degree_bound() - 1);
            B::component_scope_index(ci++);  // sharded proof: only the owner evaluates this component's constraints
            c->evaluate_constraint_quotients_on_domain(trace, acc);
        }
        B::component_scope(-1);
        B::lanes_join();
        // sharded proof: the per-log-size accumulators are sums over components (air/accumulation.rs:49-58), so summing them
        // over the ranks (mod P) gives every rank the accumulators of the whole statement
        if (B::shard_world() > 1)
            for (auto& slot : acc.sub_accumulations)
                if (slot) B::allreduce_m31(*slot);
        Span<B> span("Constraints interpolation");
        return acc.finalize(tw);
    }
    static size_t PREPROCESSED_TRACE_IDX_() { return 0; }
};

// prove (core/prover/mod.rs:28-85)
template <class B>
StarkProof prove(const std::vector<const ComponentProver<B>*>& components, Blake2sChannel& channel, CommitmentSchemeProver<B>& commitment_scheme) {
    ComponentProvers<B> provers{components, commitment_scheme.trees[0].polynomials.size()};
    Trace<B> trace{&commitment_scheme.trees};
    QM31 random_coeff = channel.draw_secure_felt();
    std::unique_ptr<HostTimer> ht0(new HostTimer("composition"));
    std::unique_ptr<Span<B>> cspan(new Span<B>("Composition"));
    std::array<CirclePoly<B>, 4> composition = provers.compute_composition_polynomial(random_coeff, trace, *commitment_scheme.twiddles);
    cspan.reset();
    ht0.reset(new HostTimer("composition_commit"));
    std::vector<CirclePoly<B>> comp_polys;
    for (auto& p : composition) comp_polys.push_back(std::move(p));
    // the SHAPE of the sample points (which column is opened at how many points) does not depend on the OODS point: built now,
    // while the composition kernels are still running; the values are filled in once the point is known
    SecurePoint shape_point;
    shape_point.x = qm_one();
    shape_point.y = qm_zero();
    MaskPoints sample_points = provers.mask_points(shape_point);
    commitment_scheme.commit_polys(std::move(comp_polys), channel);
    ht0.reset();

    SecurePoint oods_point = get_random_point(channel);
    std::unique_ptr<HostTimer> ht(new HostTimer("mask_points"));
    provers.fill_mask_points(oods_point, sample_points);
    // a component set without interaction columns commits fewer trees (TreeVec is sized by use)
    size_t n_trace_trees = commitment_scheme.trees.size() - 1;
    while (sample_points.size() > n_trace_trees && sample_points.back().empty()) sample_points.pop_back();
    sample_points.push_back(std::vector<std::vector<SecurePoint>>(4, std::vector<SecurePoint>{oods_point}));

    ht.reset();
    // sanity check (prover/mod.rs:76-82): evaluated as soon as the sampled values exist (same inputs, same outcome as
    // after prove_values; the reference only returns Err(ConstraintsNotSatisfied) later)
    auto sanity_check = [&](const std::vector<std::vector<std::vector<QM31>>>& sampled_values) {
        const auto& comp_mask = sampled_values.back();
        QM31 composition_oods_eval = qm_from_partial_evals(comp_mask[0][0], comp_mask[1][0], comp_mask[2][0], comp_mask[3][0]);
        if (composition_oods_eval != provers.eval_composition_polynomial_at_point(oods_point, sampled_values, random_coeff)) {
            if (getenv("CM31_DEBUG_SKIP_OODS_CHECK")) return;  // debugging aid: let the (invalid) proof out so it can be diffed
            throw ConstraintsNotSatisfied();
        }
    };
    return commitment_scheme.prove_values(sample_points, channel, sanity_check);
}

}  // namespace cm31

// CudaBackend: the Stwo `Backend` op surface implemented over libcm31's C ABI (include/cm31.h).
// This C++ class plays the role of the Rust `CudaBackend` shim of INTEGRATION.md: it only calls
// `cm31_*` entry points, exactly what the shim's trait impls would bind
// (Backend: external/stwo/crates/prover/src/core/backend/mod.rs:19-65).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>

#include "../../../include/cm31.h"
#include "air_expr.hpp"
#include "stark.hpp"

namespace cm31 {

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};
inline void cm_check(int status) {
    if (status != 0) throw CudaError(std::string("libcm31: ") + cm31_last_error());
}

// Single-proof sharding (csrc/shard.cu): host-side view.  Every rank runs the same driver; a column is either
// replicated (owner -1: every rank computes and holds it) or owned by one rank (the others hold an untouched buffer at
// the same arena offset and read the owner's copy through the peer mapping).  `scope_owner` is the owner of the component
// whose per-component work (trace fill, lookups, logup columns, constraint evaluation) is being issued: columns allocated
// inside the scope are tagged with it, and on the other ranks the kernel launches of that scope are skipped.
struct Shard {
    bool on = false;
    int rank = 0, world = 1;
    int scope_owner = -1;
    std::unordered_map<const u32*, int> owner;   // column base pointer -> owning rank (absent = replicated)
    std::unordered_set<const u32*> striped;      // hash layers of which every rank holds only its node range
    std::vector<int> component_owner;            // by component index (claim order)
    std::vector<double> fft_load;                // per rank: column-transform work dealt out so far in this proof
    // Column-wise work (ICFFT + LDE + OODS evaluation + the column's DEEP quotient terms) is balanced per COLUMN, not per
    // component: one component can hold half of a proof's cells (store_fp_imm in fibonacci_loop).  The rank that gets a
    // column reads its trace-domain values from the component's owner through the peer mapping (a streaming read) and
    // owns the coefficients and the LDE from then on.
    int next_fft_rank(double weight) {
        if ((int)fft_load.size() != world) fft_load.assign((size_t)world, 0.0);
        int best = 0;
        for (int r = 1; r < world; r++)
            if (fft_load[(size_t)r] < fft_load[(size_t)best]) best = r;
        fft_load[(size_t)best] += weight;
        return best;
    }
    static Shard& get() {
        static Shard s;
        return s;
    }
    bool skip() const { return on && scope_owner >= 0 && scope_owner != rank; }
    int owner_of(const u32* p) const {
        if (!on) return -1;
        auto it = owner.find(p);
        return it == owner.end() ? -1 : it->second;
    }
    bool mine(const u32* p) const {
        int o = owner_of(p);
        return o < 0 || o == rank;
    }
    void tag(const u32* p, int o) {
        if (!on) return;
        if (o < 0) owner.erase(p);
        else owner[p] = o;
    }
    void on_alloc(const u32* p) {
        if (on && scope_owner >= 0) owner[p] = scope_owner;
    }
    const u32* peer(const u32* p, int r) const {
        if (!on || r < 0 || r == rank) return p;
        const void* q = nullptr;
        if (cm31_shard_peer(p, r, &q) != 0) throw std::runtime_error(std::string("libcm31: ") + cm31_last_error());
        return (const u32*)q;
    }
    const u32* resolve(const u32* p) const { return on ? peer(p, owner_of(p)) : p; }
    u32 stripe_log() const {  // layers / domains of at least 2^stripe_log rows are split into per-rank row ranges
        u32 l = 10;
        for (int w = world; w > 1; w >>= 1) l++;
        return l;
    }
};

// Owning device column (the reference `BaseColumn`, simd/column.rs:26-30).
class DeviceCol {
   public:
    DeviceCol() {}
    explicit DeviceCol(size_t n) : n_(n) {
        cm_check(cm31_malloc((void**)&p_, n * 4));
        Shard::get().on_alloc(p_);
    }
    DeviceCol(DeviceCol&& o) noexcept : p_(o.p_), n_(o.n_), slab_(std::move(o.slab_)) {
        o.p_ = nullptr;
        o.n_ = 0;
    }
    DeviceCol& operator=(DeviceCol&& o) noexcept {
        if (this != &o) {
            release();
            p_ = o.p_;
            n_ = o.n_;
            slab_ = std::move(o.slab_);
            o.p_ = nullptr;
            o.n_ = 0;
        }
        return *this;
    }
    DeviceCol(const DeviceCol&) = delete;
    DeviceCol& operator=(const DeviceCol&) = delete;
    ~DeviceCol() { release(); }
    u32* ptr() const { return p_; }
    size_t size() const { return n_; }

    // `count` columns of n words carved out of ONE allocation (freed when the last of them dies):
    // a component's columns are created and dropped together, so this replaces dozens of
    // stream-ordered malloc/free calls per component by one pair.
    static std::vector<DeviceCol> many(size_t count, size_t n) {
        std::vector<DeviceCol> out(count);
        if (count == 0) return out;
        size_t stride = (n + 3) & ~(size_t)3;  // keep every column 16-byte aligned
        u32* base = nullptr;
        cm_check(cm31_malloc((void**)&base, count * stride * 4));
        std::shared_ptr<void> slab(base, [](void* p) { cm31_free(p); });
        for (size_t i = 0; i < count; i++) {
            out[i].p_ = base + i * stride;
            out[i].n_ = n;
            out[i].slab_ = slab;
            Shard::get().on_alloc(out[i].p_);
        }
        return out;
    }

    // columns of DIFFERENT sizes carved out of one allocation (the levels of a small Merkle tree)
    static std::vector<DeviceCol> carve(const std::vector<size_t>& sizes) {
        std::vector<DeviceCol> out(sizes.size());
        size_t total = 0;
        for (size_t n : sizes) total += (n + 3) & ~(size_t)3;
        if (total == 0) return out;
        u32* base = nullptr;
        cm_check(cm31_malloc((void**)&base, total * 4));
        std::shared_ptr<void> slab(base, [](void* p) { cm31_free(p); });
        size_t at = 0;
        for (size_t i = 0; i < sizes.size(); i++) {
            out[i].p_ = base + at;
            out[i].n_ = sizes[i];
            out[i].slab_ = slab;
            Shard::get().on_alloc(out[i].p_);
            at += (sizes[i] + 3) & ~(size_t)3;
        }
        return out;
    }

   private:
    void release() {
        if (slab_) slab_.reset();
        else if (p_) cm31_free(p_);
        p_ = nullptr;
    }
    u32* p_ = nullptr;
    size_t n_ = 0;
    std::shared_ptr<void> slab_;
};

struct CudaTwiddles {
    cm31_twiddles* h = nullptr;
    u32 log_size = 0;
    CudaTwiddles() {}
    CudaTwiddles(const CudaTwiddles&) = delete;
    ~CudaTwiddles() {
        if (h) cm31_twiddles_destroy(h);
    }
};

struct CudaBackend {
    typedef DeviceCol Col;
    typedef DeviceCol HashCol;  // 8 words per node
    typedef CudaTwiddles Twiddles;

    // components / size groups of at most 2^LANE_SPLIT_LOG rows are issued on the side lane (cm31_lane)
    static constexpr u32 LANE_SPLIT_LOG = 12;
    static void range_push(const char* name) { cm31_range_push(name); }
    static void range_pop() { cm31_range_pop(); }
    static void lane(u32 log_size) {
        if (Shard::get().on) return;  // a sharded proof keeps one stream: its NCCL calls must be issued in one order on every rank
        cm_check(cm31_lane(log_size <= LANE_SPLIT_LOG ? 1 : 0));
    }
    static void lanes_join() {
        air_batch_flush();  // (defensive: the phases flush explicitly before they join)
        cm_check(cm31_lanes_join());
    }
    // ---- batched small programs (cm31_air_program_batch).  cairo-m proves all 34 components for every segment; the opcode
    // components a program does not use are 16 padding rows each (21 of 26 for fibonacci_loop), and their trace-fill, lookup
    // and logup programs cost ~7 us of launch path each when issued one by one.  Inside a BatchScope, programs over at most
    // 2^SMALL_BATCH_LOG rows are recorded instead of launched; air_batch_flush() issues them as ONE interpreter launch (plus
    // one launch for the recorded logup finalisations).  The caller guarantees that a recorded program's input and output
    // columns stay alive until the flush and that recorded programs are independent of each other.
    static constexpr u32 SMALL_BATCH_LOG = 6;
    struct AirBatch {
        struct Item {
            std::vector<const u32*> in;
            std::vector<u32*> out;
            std::vector<uint64_t> code;
            std::vector<u32> consts;
            u32 log_size, n_regs, hist_bins;
        };
        std::vector<Item> items;
        std::vector<cm31_logup_finalize_item> finals;
        int depth = 0;
    };
    static AirBatch& air_batch() {
        static AirBatch b;
        return b;
    }
    static bool batch_small_components() {
        static const bool off = getenv("CM31_NO_BATCH") != nullptr;
        return !off && !Shard::get().on;
    }
    struct BatchScope {
        bool on;
        explicit BatchScope(bool enable) : on(enable && batch_small_components()) {
            if (on) air_batch().depth++;
        }
        ~BatchScope() {
            if (on) air_batch().depth--;
        }
        BatchScope(const BatchScope&) = delete;
        BatchScope& operator=(const BatchScope&) = delete;
    };
    // Programs that need more than 512 interpreter registers stay on their own (generated) kernels: the interpreter keeps its
    // register file in local memory, and a launch that needs more local memory per thread than any kernel before makes the
    // driver re-size its local-memory arena -- a device-wide synchronisation, and a shrink afterwards (measured on the sha256
    // workload, whose unused components include 1024-register programs: trace phase 10 -> 95 ms).
    static constexpr u32 BATCH_MAX_REGS = 512;
    static bool batching(u32 log_size, u32 n_regs) { return air_batch().depth > 0 && log_size <= SMALL_BATCH_LOG && n_regs <= BATCH_MAX_REGS; }
    static void air_batch_flush() {
        AirBatch& b = air_batch();
        if (b.items.empty() && b.finals.empty()) return;
        // on the SIDE lane: the columns of the recorded programs were allocated there (lane(small size)), so allocation,
        // writes and reads follow one stream order; what they read from lane 0 was issued before the fork
        cm_check(cm31_lane(1));
        if (!b.items.empty()) {
            std::vector<cm31_air_batch_item> v(b.items.size());
            for (size_t i = 0; i < v.size(); i++) {
                const AirBatch::Item& it = b.items[i];
                v[i] = cm31_air_batch_item{it.in.data(), it.in.size(), it.out.data(), it.out.size(), it.log_size, it.code.data(), it.code.size(),
                                           it.n_regs, it.consts.data(), it.consts.size(), it.hist_bins};
            }
            cm_check(cm31_air_program_batch(v.data(), v.size()));
            b.items.clear();
        }
        if (!b.finals.empty()) {
            cm_check(cm31_logup_finalize_small_batch(b.finals.data(), b.finals.size()));
            b.finals.clear();
        }
    }
    static size_t len(const Col& c) { return c.size(); }
    // ---- single-proof sharding hooks (no-ops when sharding is off)
    static int shard_world() { return Shard::get().on ? Shard::get().world : 1; }
    static int shard_rank() { return Shard::get().rank; }
    static u32 shard_stripe_log() { return Shard::get().stripe_log(); }
    static void shard_begin_proof() {
        Shard& sh = Shard::get();
        int rank = 0, world = 0;
        cm_check(cm31_shard_info(&rank, &world, nullptr));
        sh.on = world > 0;
        sh.rank = rank;
        sh.world = world > 0 ? world : 1;
        sh.scope_owner = -1;
        sh.owner.clear();
        sh.striped.clear();
        sh.component_owner.clear();
        sh.fft_load.clear();
        if (!sh.on) return;
        cm_check(cm31_shard_arena_reset());  // also switches cm31_malloc to the arena until shard_end_proof
        sums().buf = DeviceCol(4 * 256);  // the persistent claimed-sum buffer lived in the arena that was just reset
        sums().used = 0;
        sums().owned.clear();
    }
    static void shard_end_proof() {
        Shard& sh = Shard::get();
        if (!sh.on) return;
        sh.scope_owner = -1;
        cm31_shard_arena_suspend();  // buffers allocated between proofs (staged inputs) come from the ordinary pool
    }
    static void component_scope(int owner) { Shard::get().scope_owner = Shard::get().on ? owner : -1; }
    static void component_scope_index(size_t component) {
        Shard& sh = Shard::get();
        sh.scope_owner = sh.on && component < sh.component_owner.size() ? sh.component_owner[component] : -1;
    }
    static void set_component_owners(const std::vector<int>& owners) { Shard::get().component_owner = owners; }
    static void set_owner(const Col& c, int owner) { Shard::get().tag(c.ptr(), owner); }
    static void shard_barrier() {
        if (Shard::get().on && Shard::get().world > 1) cm_check(cm31_shard_barrier());
    }
    static void debug_checksum(const char* what, const Col& c) {
        if (!getenv("CM31_SHARD_DEBUG")) return;
        std::vector<u32> h(c.size());
        cm_check(cm31_d2h(h.data(), c.ptr(), c.size() * 4));
        cm_check(cm31_sync());
        u64 sum = 0, mix = 0;
        for (size_t i = 0; i < h.size(); i++) sum += h[i], mix = mix * 1000003u + h[i];
        fprintf(stderr, "[shard debug rank %d] %s: n=%zu sum=%llu mix=%016llx ptr_off=%p\n", Shard::get().rank, what, h.size(), (unsigned long long)sum,
                (unsigned long long)mix, (void*)c.ptr());
    }
    static void allreduce_bins(Col& bins) {
        debug_checksum("bins before allreduce", bins);
        if (Shard::get().on) cm_check(cm31_shard_allreduce_u32(bins.ptr(), bins.size()));
        debug_checksum("bins after allreduce", bins);
    }
    // values: one entry per item, valid on the rank that computed it; keep[i] = this rank contributes entry i
    static void exchange_qm31(std::vector<QM31>& values, const std::vector<char>& keep) {
        Shard& sh = Shard::get();
        if (!sh.on || sh.world == 1) return;
        std::vector<u32> w(4 * values.size() + 4, 0);
        for (size_t i = 0; i < values.size(); i++)
            if (keep[i]) {
                w[4 * i] = values[i].a, w[4 * i + 1] = values[i].b, w[4 * i + 2] = values[i].c, w[4 * i + 3] = values[i].d;
            }
        cm_check(cm31_shard_allreduce_host_u32(w.data(), 4 * values.size()));
        for (size_t i = 0; i < values.size(); i++) values[i] = qm_make(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
    // mod-P sum over ranks of an accumulator (4 coordinate columns): returns the reduced copy, complete on every rank
    static void allreduce_m31(std::array<Col, 4>& acc) {
        Shard& sh = Shard::get();
        if (!sh.on || sh.world == 1) return;
        std::array<Col, 4> out;
        u32* d4[4];
        const u32* s4[4];
        for (int k = 0; k < 4; k++) {
            out[k] = Col(acc[k].size());
            sh.tag(out[k].ptr(), -1);
            d4[k] = out[k].ptr();
            s4[k] = acc[k].ptr();
        }
        cm_check(cm31_shard_reduce_m31(d4, s4, acc[0].size()));
        for (int k = 0; k < 4; k++) acc[k] = std::move(out[k]);
    }
    // hash layer of which this rank holds nodes [rank * n / world, (rank + 1) * n / world)
    static HashCol commit_on_layer_striped(u32 log_size, const HashCol* prev, const std::vector<const Col*>& cols) {
        Shard& sh = Shard::get();
        HashCol out(((size_t)1 << log_size) * 8);
        sh.tag(out.ptr(), -1);
        auto s = cptrs(cols);
        const size_t count = ((size_t)1 << log_size) / (size_t)sh.world, first = count * (size_t)sh.rank;
        cm_check(cm31_blake2s_commit_layer_range(log_size, prev ? prev->ptr() : nullptr, s.data(), s.size(), out.ptr(), first, count));
        sh.striped.insert(out.ptr());
        return out;
    }
    // several striped layers in one launch (only the first carries columns): this rank's node range and the subtree above it
    static std::vector<HashCol> commit_layers_fused_striped(u32 log_size, const HashCol* prev, const std::vector<const Col*>& cols, u32 n_levels) {
        Shard& sh = Shard::get();
        std::vector<HashCol> out;
        std::vector<u32*> outp;
        for (u32 l = 0; l < n_levels; l++) {
            out.emplace_back(((size_t)1 << (log_size - l)) * 8);
            sh.tag(out.back().ptr(), -1);
            sh.striped.insert(out.back().ptr());
            outp.push_back(out.back().ptr());
        }
        auto s = cptrs(cols);
        const size_t count = ((size_t)1 << log_size) / (size_t)sh.world, first = count * (size_t)sh.rank;
        cm_check(cm31_blake2s_commit_multi_range(log_size, prev ? prev->ptr() : nullptr, s.data(), s.size(), n_levels, outp.data(), first, count));
        return out;
    }
    // the last striped layer becomes complete on every rank (the all-gather at the Merkle root of the north star)
    static void join_striped_layer(HashCol& layer) {
        Shard& sh = Shard::get();
        cm_check(cm31_shard_allgather(layer.ptr(), layer.size() * 4 / (size_t)sh.world));
        sh.striped.erase(layer.ptr());
    }
    static bool is_striped(const HashCol& layer) { return Shard::get().on && Shard::get().striped.count(layer.ptr()) != 0; }
    // where node `node` of a hash layer can be read from: the owner's copy for a striped layer
    static const u32* hash_node_source(const HashCol& layer, size_t node) {
        Shard& sh = Shard::get();
        if (!sh.on || !sh.striped.count(layer.ptr())) return layer.ptr();
        const size_t per = (layer.size() / 8) / (size_t)sh.world;
        return sh.peer(layer.ptr(), (int)(node / per));
    }
    static Col zeros(size_t n) {
        Col c(n);
        cm_check(cm31_memset0(c.ptr(), n * 4));
        return c;
    }
    static Col uninit(size_t n) { return Col(n); }
    static std::vector<Col> uninit_many(size_t count, size_t n) { return Col::many(count, n); }
    static Col from_host(const u32* src, size_t n) {
        Col c(n);
        cm_check(cm31_h2d(c.ptr(), src, n * 4));
        return c;
    }
    static void to_host(const Col& c, u32* out) { cm_check(cm31_d2h(out, c.ptr(), c.size() * 4)); }
    static void precompute_twiddles(u32 log_size, Twiddles& out) {
        cm_check(cm31_twiddles_create(log_size, &out.h));
        out.log_size = log_size;
    }
    static std::vector<const u32*> cptrs(const std::vector<const Col*>& cols) {  // columns another rank owns: its copy
        std::vector<const u32*> p;
        const Shard& sh = Shard::get();
        for (auto* c : cols) p.push_back(sh.resolve(c->ptr()));
        return p;
    }
    static std::vector<u32*> ptrs(const std::vector<Col*>& cols) {
        std::vector<u32*> p;
        for (auto* c : cols) p.push_back(c->ptr());
        return p;
    }
    // Column-wise transforms: a rank only transforms the columns it owns (and the replicated ones); outputs inherit the owner.
    static void interpolate_columns(const std::vector<Col*>& cols, u32 log_size, const Twiddles& tw) {
        const Shard& sh = Shard::get();
        std::vector<u32*> p;
        for (auto* c : cols)
            if (sh.mine(c->ptr())) p.push_back(c->ptr());
        if (!p.empty()) cm_check(cm31_interpolate_batch(p.data(), p.size(), log_size, tw.h));
    }
    static void interpolate_columns_to(const std::vector<const Col*>& evals, const std::vector<Col*>& outs, u32 log_size, const Twiddles& tw) {
        Shard& sh = Shard::get();
        std::vector<const u32*> s;
        std::vector<u32*> d;
        for (size_t i = 0; i < evals.size(); i++) {
            int o = sh.owner_of(evals[i]->ptr());
            // sharded: an owned column's transform goes to the least loaded rank (it reads the values from the owner);
            // replicated columns stay replicated
            // (from 4 ranks on: with 2 ranks the remote reads cost more than the better balance gains, measured on 2^22 / 2^24)
            if (sh.on && sh.world >= 4 && o >= 0) o = sh.next_fft_rank((double)log_size * (double)((size_t)1 << log_size));
            sh.tag(outs[i]->ptr(), o);
            if (o >= 0 && o != sh.rank) continue;
            s.push_back(sh.resolve(evals[i]->ptr()));
            d.push_back(outs[i]->ptr());
        }
        if (!s.empty()) cm_check(cm31_interpolate_batch_to(s.data(), d.data(), s.size(), log_size, tw.h));
    }
    static void evaluate_polynomials(const std::vector<const Col*>& polys, const std::vector<Col*>& outs, u32 log_size, u32 log_eval,
                                     const Twiddles& tw) {
        Shard& sh = Shard::get();
        std::vector<const u32*> s;
        std::vector<u32*> d;
        for (size_t i = 0; i < polys.size(); i++) {
            sh.tag(outs[i]->ptr(), sh.owner_of(polys[i]->ptr()));
            if (!sh.mine(polys[i]->ptr())) continue;
            s.push_back(polys[i]->ptr());
            d.push_back(outs[i]->ptr());
        }
        if (!s.empty()) cm_check(cm31_evaluate_batch(s.data(), d.data(), s.size(), log_size, log_eval, tw.h));
    }
    static void eval_at_points(const std::vector<const Col*>& polys, const std::vector<u32>& log_sizes, const std::vector<SecurePoint>& points,
                               const std::vector<u32>& point_idx, std::vector<QM31>& out) {
        const Shard& sh = Shard::get();
        std::vector<u32> pts;
        for (auto& p : points)
            for (u32 w : {p.x.a, p.x.b, p.x.c, p.x.d, p.y.a, p.y.b, p.y.c, p.y.d}) pts.push_back(w);
        // a rank evaluates the polynomials it owns; replicated ones are spread round-robin; the values are then exchanged
        std::vector<size_t> sel;
        std::vector<char> keep(polys.size(), 0);
        for (size_t i = 0; i < polys.size(); i++) {
            int o = sh.owner_of(polys[i]->ptr());
            if (!sh.on || o == sh.rank || (o < 0 && (int)(i % (size_t)sh.world) == sh.rank)) {
                sel.push_back(i);
                keep[i] = 1;
            }
        }
        std::vector<const u32*> s;
        std::vector<u32> ls, pi;
        for (size_t i : sel) {
            s.push_back(polys[i]->ptr());
            ls.push_back(log_sizes[i]);
            pi.push_back(point_idx[i]);
        }
        std::vector<u32> res(4 * sel.size() + 4);
        if (!sel.empty()) cm_check(cm31_eval_at_point_batch(s.data(), ls.data(), s.size(), pts.data(), points.size(), pi.data(), res.data()));
        out.assign(polys.size(), qm_make(0, 0, 0, 0));
        for (size_t k = 0; k < sel.size(); k++) out[sel[k]] = qm_make(res[4 * k], res[4 * k + 1], res[4 * k + 2], res[4 * k + 3]);
        exchange_qm31(out, keep);
    }
    static HashCol commit_on_layer(u32 log_size, const HashCol* prev, const std::vector<const Col*>& cols) {
        HashCol out(((size_t)1 << log_size) * 8);
        auto s = cptrs(cols);
        cm_check(cm31_blake2s_commit_layer(log_size, prev ? prev->ptr() : nullptr, s.data(), s.size(), out.ptr()));
        return out;
    }
    // layers log_size .. log_size-n_levels+1 (only the first with columns) in one launch; returns them in that order
    static std::vector<HashCol> commit_layers_fused(u32 log_size, const HashCol* prev, const std::vector<const Col*>& cols, u32 n_levels) {
        std::vector<HashCol> out;
        std::vector<u32*> outp;
        for (u32 l = 0; l < n_levels; l++) {
            out.emplace_back(((size_t)1 << (log_size - l)) * 8);
            outp.push_back(out.back().ptr());
        }
        auto s = cptrs(cols);
        cm_check(cm31_blake2s_commit_multi(log_size, prev ? prev->ptr() : nullptr, s.data(), s.size(), n_levels, outp.data()));
        return out;
    }
    static void gather(const std::vector<const Col*>& cols, const std::vector<u32>& idx, std::vector<std::vector<u32>>& out) {
        auto s = cptrs(cols);
        std::vector<u32> flat(cols.size() * idx.size());
        cm_check(cm31_gather_u32(s.data(), s.size(), idx.data(), idx.size(), flat.data()));
        out.assign(cols.size(), std::vector<u32>());
        for (size_t c = 0; c < cols.size(); c++) out[c].assign(flat.begin() + c * idx.size(), flat.begin() + (c + 1) * idx.size());
    }
    // layers top_log..0 in one launch; cols_by_layer[l] = columns of 2^l rows. Returns layers[0..top_log].
    static std::vector<HashCol> commit_top_layers(u32 top_log, const HashCol* prev, const std::vector<std::vector<const Col*>>& cols_by_layer) {
        std::vector<HashCol> out;
        std::vector<u32*> outp;
        std::vector<const u32*> cols;
        std::vector<u32> start;
        for (u32 l = 0; l <= top_log; l++) {
            out.emplace_back(((size_t)1 << l) * 8);
            outp.push_back(out.back().ptr());
            start.push_back((u32)cols.size());
            for (auto* c : cols_by_layer[l]) cols.push_back(Shard::get().resolve(c->ptr()));  // the owner's copy
        }
        start.push_back((u32)cols.size());
        if (getenv("CM31_SHARD_DEBUG")) {  // what this rank sees through the (possibly peer) pointers
            u64 mix = 0;
            size_t n_peer = 0;
            for (u32 l = 0; l <= top_log; l++)
                for (auto* c : cols_by_layer[l]) {
                    const u32* q = Shard::get().resolve(c->ptr());
                    n_peer += q != c->ptr();
                    std::vector<u32> h(c->size());
                    cm_check(cm31_d2h(h.data(), q, h.size() * 4));
                    cm_check(cm31_sync());
                    for (u32 w : h) mix = mix * 1000003u + w;
                }
            fprintf(stderr, "[shard debug rank %d] commit_top top_log=%u cols=%zu peer=%zu mix=%016llx\n", Shard::get().rank, top_log, cols.size(), n_peer,
                    (unsigned long long)mix);
        }
        cm_check(cm31_blake2s_commit_top(top_log, prev ? prev->ptr() : nullptr, cols.data(), start.data(), outp.data()));
        return out;
    }
    static const u32* col_words(const Col& c) { return Shard::get().resolve(c.ptr()); }  // decommitment reads go to the owner's copy
    static const u32* hash_words(const HashCol& c) { return c.ptr(); }
    // the gathered words land in a page-locked buffer of the library; gather_wait() blocks until they have
    static const u32* gather_runs_async(const std::vector<const u32*>& srcs, const std::vector<u32>& src_id, const std::vector<u32>& word,
                                        const std::vector<u32>& out_off, const std::vector<u32>& cnt, const std::vector<u32>& grid_desc,
                                        const std::vector<u32>& grid_cols, const std::vector<u32>& grid_rows, size_t n_words,
                                        std::vector<u32>& /*store*/) {
        const u32* res = nullptr;
        cm_check(cm31_gather_batch_async(srcs.data(), srcs.size(), src_id.data(), word.data(), out_off.data(), cnt.data(), src_id.size(),
                                         grid_desc.data(), grid_desc.size() / 5, grid_cols.data(), grid_cols.size(), grid_rows.data(),
                                         grid_rows.size(), n_words, &res));
        return res;
    }
    static void gather_wait() { cm_check(cm31_gather_wait()); }
    // Deferred proof tails (cm31_prove_cairo_m_async): while `defer` is set, prove_values leaves the host-side assembly of the
    // decommitments to a closure; the owner of the pending proof registers `hook`, which runs it (and serialises the proof).
    // One stage of the hook runs at the first Merkle-root read after each idle_gate_open() -- the prover opens the gate before
    // the commitments of the execution traces (stage: queries, decommitment plan, gather) and of the interaction traces
    // (stage: assembly, serialisation), where milliseconds of FFT / Merkle kernels are queued and the host would only wait --
    // and whatever is left runs before the next proof defers its own tail (finish_deferred_tails).
    struct TailState {
        bool defer = false;
        bool gate = false;
        std::function<bool()> hook;  // runs ONE stage of the pending proof's tail; true after the last
    };
    static TailState& tail_state() {
        static TailState t;
        return t;
    }
    static bool defer_proof_tail() { return tail_state().defer && !Shard::get().on; }
    static void run_tail_stage() {
        TailState& t = tail_state();
        if (!t.hook) return;
        std::function<bool()> h = t.hook;  // (the stage may register nothing new; keep the hook until it reports completion)
        if (h()) t.hook = nullptr;
    }
    static void finish_deferred_tails() {
        TailState& t = tail_state();
        t.gate = false;
        while (t.hook) run_tail_stage();
    }
    static void idle_gate_open() { tail_state().gate = true; }
    static Hash32 read_root(const HashCol& root_layer) {
        if (tail_state().gate) {  // the kernels of this tree are queued: host time is free until its root arrives
            tail_state().gate = false;
            run_tail_stage();
        }
        Hash32 h;
        cm_check(cm31_d2h(h.b, root_layer.ptr(), 32));
        if (getenv("CM31_SHARD_DEBUG")) {
            cm_check(cm31_sync());
            fprintf(stderr, "[shard debug rank %d] root %02x%02x%02x%02x at %p\n", Shard::get().rank, h.b[0], h.b[1], h.b[2], h.b[3], (void*)root_layer.ptr());
        }
        return h;
    }
    static void gather_hashes(const HashCol& layer, const std::vector<u32>& idx, std::vector<Hash32>& out) {
        out.resize(idx.size());
        cm_check(cm31_gather_hash(layer.ptr(), idx.data(), idx.size(), (u32*)out.data()));
    }
    static std::array<Col, 4> accumulate_quotients(u32 log_size, const std::vector<const Col*>& cols, QM31 random_coeff,
                                                   const std::vector<ColumnSampleBatch>& batches, u32 /*log_blowup_factor*/) {
        std::array<Col, 4> out;
        u32* o4[4];
        for (int k = 0; k < 4; k++) {
            out[k] = Col((size_t)1 << log_size);
            o4[k] = out[k].ptr();
        }
        std::vector<u32> pts, starts = {0}, idx, vals;
        for (auto& b : batches) {
            for (u32 w : {b.point.x.a, b.point.x.b, b.point.x.c, b.point.x.d, b.point.y.a, b.point.y.b, b.point.y.c, b.point.y.d}) pts.push_back(w);
            for (auto& cv : b.columns_and_values) {
                idx.push_back((u32)cv.first);
                for (u32 w : {cv.second.a, cv.second.b, cv.second.c, cv.second.d}) vals.push_back(w);
            }
            starts.push_back((u32)idx.size());
        }
        auto s = cptrs(cols);
        u32 rc[4] = {random_coeff.a, random_coeff.b, random_coeff.c, random_coeff.d};
        Shard& sh = Shard::get();
        if (sh.on && sh.world > 1) {
            // The quotient is linear in the per-column terms: every rank accumulates the terms of the columns it owns
            // (replicated columns: round-robin) over ALL rows, from local memory only; the partial quotients are then summed
            // mod P over the ranks (reduce-scatter by peer loads + all-gather, cm31_shard_reduce_m31).
            std::vector<uint8_t> active(idx.size(), 0);
            for (size_t k = 0; k < idx.size(); k++) {
                int o = sh.owner_of(cols[idx[k]]->ptr());
                active[k] = (o == sh.rank || (o < 0 && (int)(idx[k] % (u32)sh.world) == sh.rank)) ? 1 : 0;
            }
            std::vector<const u32*> local;
            for (auto* c : cols) local.push_back(c->ptr());  // only active (= locally valid) columns are dereferenced
            std::array<Col, 4> part;
            u32* p4[4];
            const u32* s4[4];
            for (int k = 0; k < 4; k++) {
                part[k] = Col((size_t)1 << log_size);
                sh.tag(part[k].ptr(), -1);
                p4[k] = part[k].ptr();
                s4[k] = part[k].ptr();
            }
            cm_check(cm31_accumulate_quotients_partial(log_size, local.data(), local.size(), rc, batches.size(), pts.data(), starts.data(), idx.data(),
                                                       vals.data(), p4, 0, (size_t)1 << log_size, active.data()));
            cm_check(cm31_shard_reduce_m31(o4, s4, (size_t)1 << log_size));
        } else {
            cm_check(cm31_accumulate_quotients(log_size, s.data(), s.size(), rc, batches.size(), pts.data(), starts.data(), idx.data(), vals.data(), o4));
        }
        return out;
    }
    static std::array<Col, 4> fold_line(const std::array<Col, 4>& src, u32 log_size, QM31 alpha, const Twiddles& tw) {
        std::array<Col, 4> out;
        const u32* s4[4];
        u32* d4[4];
        for (int k = 0; k < 4; k++) {
            out[k] = Col((size_t)1 << (log_size - 1));
            s4[k] = src[k].ptr();
            d4[k] = out[k].ptr();
        }
        u32 a[4] = {alpha.a, alpha.b, alpha.c, alpha.d};
        cm_check(cm31_fold_line(s4, log_size, a, tw.h, d4));
        return out;
    }
    static void fold_circle_into_line(std::array<Col, 4>& dst, const std::array<Col, 4>& src, u32 log_size, QM31 alpha, const Twiddles& tw) {
        const u32* s4[4];
        u32* d4[4];
        for (int k = 0; k < 4; k++) {
            s4[k] = src[k].ptr();
            d4[k] = dst[k].ptr();
        }
        u32 a[4] = {alpha.a, alpha.b, alpha.c, alpha.d};
        cm_check(cm31_fold_circle_into_line(d4, s4, log_size, a, tw.h));
    }
    static void accumulate(std::array<Col, 4>& dst, const std::array<Col, 4>& src) {
        const u32* s4[4];
        u32* d4[4];
        for (int k = 0; k < 4; k++) {
            s4[k] = src[k].ptr();
            d4[k] = dst[k].ptr();
        }
        cm_check(cm31_accumulate(d4, s4, dst[0].size()));
    }
    // The inner FRI layers of at most 2^fri_tail_log() points: one launch for the whole chain (cm31_fri_tail), the channel
    // replayed on the host from the roots.  Appends the layers to `inner_layers`, leaves the last evaluation in layer_eval.
    static u32 fri_tail_log() {
        static const u32 l = getenv("CM31_FRI_TAIL_LOG") ? (u32)atoi(getenv("CM31_FRI_TAIL_LOG")) : 12;
        return Shard::get().on ? 0 : std::min<u32>(l, 14);
    }
    template <class InnerLayer, class SecureEval>
    static void fri_tail(Blake2sChannel& channel, std::array<Col, 4>& layer_eval, u32& layer_log, u32 last_log,
                         const std::vector<SecureEval>& columns, size_t& next_col, const Twiddles& tw, std::vector<InnerLayer>& inner_layers) {
        const size_t n_layers = layer_log - last_log;
        std::vector<cm31_fri_tail_layer> desc(n_layers);
        std::vector<std::vector<u32*>> level_ptrs(n_layers);
        std::vector<std::vector<HashCol>> trees(n_layers);
        std::vector<std::array<Col, 4>> evals(n_layers + 1);
        evals[0] = std::move(layer_eval);
        for (size_t i = 0; i < n_layers; i++) {
            const u32 k = layer_log - (u32)i;
            std::vector<Col> nxt = Col::many(4, (size_t)1 << (k - 1));
            for (int c = 0; c < 4; c++) evals[i + 1][c] = std::move(nxt[c]);
            std::vector<size_t> sizes;
            for (u32 j = 0; j <= k; j++) sizes.push_back(((size_t)1 << j) * 8);
            trees[i] = DeviceCol::carve(sizes);
            for (u32 j = 0; j <= k; j++) level_ptrs[i].push_back(trees[i][j].ptr());
            cm31_fri_tail_layer& d = desc[i];
            for (int c = 0; c < 4; c++) {
                d.ev_in[c] = evals[i][c].ptr();
                d.ev_out[c] = evals[i + 1][c].ptr();
                d.circle[c] = nullptr;
            }
            d.tree_levels = level_ptrs[i].data();
            d.log_size = k;
            if (next_col < columns.size() && columns[next_col].log_size == k) {  // joins the line of 2^(k-1) points (fri.rs:255-262)
                for (int c = 0; c < 4; c++) d.circle[c] = columns[next_col].columns[c].ptr();
                next_col++;
            }
        }
        const Hash32 digest = channel.digest();
        u32 din[8];
        memcpy(din, digest.b, 32);
        std::vector<u32> roots(8 * n_layers);
        cm_check(cm31_fri_tail(din, desc.data(), n_layers, tw.h, roots.data()));
        for (size_t i = 0; i < n_layers; i++) {
            InnerLayer layer;
            memcpy(layer.root.b, &roots[8 * i], 32);
            channel.mix_root(layer.root);
            (void)channel.draw_secure_felt();  // the folding alpha the kernel derived from the same digest
            layer.log_size = layer_log - (u32)i;
            layer.evaluation = std::move(evals[i]);
            layer.merkle_tree.layers = std::move(trees[i]);  // [0] = root layer .. [k] = leaves
            layer.merkle_tree.root_ = layer.root;
            layer.merkle_tree.root_read_ = true;
            inner_layers.push_back(std::move(layer));
        }
        layer_eval = std::move(evals[n_layers]);
        layer_log = last_log;
    }
    static std::vector<QM31> generate_secure_powers(QM31 felt, size_t n) {
        std::vector<u32> out(4 * n + 4);
        u32 f[4] = {felt.a, felt.b, felt.c, felt.d};
        cm_check(cm31_secure_powers(f, n, out.data()));
        std::vector<QM31> r(n);
        for (size_t i = 0; i < n; i++) r[i] = qm_make(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
        return r;
    }
    static u64 grind(const Hash32& digest, u32 pow_bits) {
        uint64_t nonce = 0;
        cm_check(cm31_grind_blake2s(digest.b, pow_bits, &nonce));
        return nonce;
    }
    static void constraint_eval(const std::vector<const Col*>& cols, u32 trace_log, u32 eval_log, const AirProgram& prog,
                                const std::vector<u32>& denom_inv, std::array<Col, 4>& acc) {
        if (Shard::get().skip()) return;  // another rank owns this component
        auto s = cptrs(cols);
        u32* a4[4] = {acc[0].ptr(), acc[1].ptr(), acc[2].ptr(), acc[3].ptr()};
        cm_check(cm31_constraint_eval(s.data(), s.size(), trace_log, eval_log, prog.code.data(), prog.code.size(), prog.n_regs,
                                      prog.consts.data(), prog.consts.size(), denom_inv.data(), a4));
    }
    static void air_program(const std::vector<const Col*>& in, const std::vector<Col*>& out, u32 log_size, const AirProgram& prog) {
        if (Shard::get().skip()) return;
        auto s = cptrs(in);
        auto d = ptrs(out);
        if (batching(log_size, prog.n_regs)) {
            air_batch().items.push_back(AirBatch::Item{std::move(s), std::move(d), prog.code, prog.consts, log_size, prog.n_regs, 0});
            return;
        }
        cm_check(cm31_air_program(s.data(), s.size(), d.data(), d.size(), log_size, prog.code.data(), prog.code.size(), prog.n_regs,
                                  prog.consts.data(), prog.consts.size()));
    }
    // multiplicity histogram: `bins` is the whole bin column, so its length bounds every looked-up value
    static void air_lookups(const std::vector<const Col*>& in, Col& bins, u32 log_size, const AirProgram& prog) {
        if (Shard::get().skip()) return;  // the owner counts this component's lookups; the bins are summed over ranks afterwards
        auto s = cptrs(in);
        u32 log_bins = 0;
        while (((size_t)1 << log_bins) < bins.size()) log_bins++;
        if (((size_t)1 << log_bins) != bins.size()) throw std::logic_error("air_lookups: the bin column is not a power of two");
        if (batching(log_size, prog.n_regs)) {
            air_batch().items.push_back(AirBatch::Item{std::move(s), std::vector<u32*>{bins.ptr()}, prog.code, prog.consts, log_size, prog.n_regs, 1u << log_bins});
            return;
        }
        cm_check(cm31_air_lookups(s.data(), s.size(), bins.ptr(), log_bins, log_size, prog.code.data(), prog.code.size(), prog.n_regs,
                                  prog.consts.data(), prog.consts.size()));
    }
    static void check_air_errors() { cm_check(cm31_air_error_check()); }
    // Claimed sums of one interaction phase: finalize_last is stream-ordered and writes into a device
    // arena; collect_sums() reads every pending sum with ONE copy.
    struct SumArena {
        DeviceCol buf;
        size_t used = 0;
        std::vector<char> owned;  // per pending slot: this rank computed it (single-proof sharding)
    };
    static SumArena& sums() {
        static SumArena a;
        return a;
    }
    static void prepare() {  // persistent buffers are created on lane 0, before any fork
        SumArena& a = sums();
        if (a.buf.size() == 0) a.buf = DeviceCol(4 * 256);
    }
    static size_t logup_finalize_last_async(const std::array<Col*, 4>& last, u32 log_size) {
        SumArena& a = sums();
        if (a.buf.size() == 0) a.buf = DeviceCol(4 * 256);
        if (a.used >= 256) throw CudaError("too many pending claimed sums");
        u32* l4[4] = {last[0]->ptr(), last[1]->ptr(), last[2]->ptr(), last[3]->ptr()};
        a.owned.push_back(Shard::get().skip() ? 0 : 1);
        if (a.owned.back() && batching(log_size, 0))  // after the logup program that writes these columns, recorded or already issued on the side lane
            air_batch().finals.push_back(cm31_logup_finalize_item{{l4[0], l4[1], l4[2], l4[3]}, log_size, a.buf.ptr() + 4 * a.used});
        else if (a.owned.back())
            cm_check(cm31_logup_finalize_last_async(l4, log_size, a.buf.ptr() + 4 * a.used));
        return a.used++;
    }
    static std::vector<QM31> collect_sums() {
        SumArena& a = sums();
        std::vector<u32> h(4 * a.used + 4);
        if (a.used) cm_check(cm31_d2h(h.data(), a.buf.ptr(), a.used * 16));
        std::vector<QM31> out;
        for (size_t i = 0; i < a.used; i++) out.push_back(qm_make(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]));
        exchange_qm31(out, a.owned);  // sharded: every claimed sum was computed by the rank that owns the component
        a.used = 0;
        a.owned.clear();
        return out;
    }
    static QM31 logup_finalize_last(const std::array<Col*, 4>& last, u32 log_size) {
        if (Shard::get().on) throw CudaError("logup_finalize_last: the synchronous form is not used by the sharded prover");
        u32* l4[4] = {last[0]->ptr(), last[1]->ptr(), last[2]->ptr(), last[3]->ptr()};
        u32 cs[4];
        cm_check(cm31_logup_finalize_last(l4, log_size, cs));
        return qm_make(cs[0], cs[1], cs[2], cs[3]);
    }
};

}  // namespace cm31

// CudaBackend: the Stwo `Backend` op surface implemented over libcm31's C ABI (include/cm31.h).
// This C++ class plays the role of the Rust `CudaBackend` shim of INTEGRATION.md: it only calls
// `cm31_*` entry points, exactly what the shim's trait impls would bind
// (Backend: external/stwo/crates/prover/src/core/backend/mod.rs:19-65).
#pragma once
#include <memory>
#include <stdexcept>
#include <string>

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

// Owning device column (the reference `BaseColumn`, simd/column.rs:26-30).
class DeviceCol {
   public:
    DeviceCol() {}
    explicit DeviceCol(size_t n) : n_(n) { cm_check(cm31_malloc((void**)&p_, n * 4)); }
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
    static void lane(u32 log_size) { cm_check(cm31_lane(log_size <= LANE_SPLIT_LOG ? 1 : 0)); }
    static void lanes_join() { cm_check(cm31_lanes_join()); }
    static size_t len(const Col& c) { return c.size(); }
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
    static std::vector<const u32*> cptrs(const std::vector<const Col*>& cols) {
        std::vector<const u32*> p;
        for (auto* c : cols) p.push_back(c->ptr());
        return p;
    }
    static std::vector<u32*> ptrs(const std::vector<Col*>& cols) {
        std::vector<u32*> p;
        for (auto* c : cols) p.push_back(c->ptr());
        return p;
    }
    static void interpolate_columns(const std::vector<Col*>& cols, u32 log_size, const Twiddles& tw) {
        auto p = ptrs(cols);
        cm_check(cm31_interpolate_batch(p.data(), p.size(), log_size, tw.h));
    }
    static void interpolate_columns_to(const std::vector<const Col*>& evals, const std::vector<Col*>& outs, u32 log_size, const Twiddles& tw) {
        auto s = cptrs(evals);
        auto d = ptrs(outs);
        cm_check(cm31_interpolate_batch_to(s.data(), d.data(), s.size(), log_size, tw.h));
    }
    static void evaluate_polynomials(const std::vector<const Col*>& polys, const std::vector<Col*>& outs, u32 log_size, u32 log_eval,
                                     const Twiddles& tw) {
        auto s = cptrs(polys);
        auto d = ptrs(outs);
        cm_check(cm31_evaluate_batch(s.data(), d.data(), s.size(), log_size, log_eval, tw.h));
    }
    static void eval_at_points(const std::vector<const Col*>& polys, const std::vector<u32>& log_sizes, const std::vector<SecurePoint>& points,
                               const std::vector<u32>& point_idx, std::vector<QM31>& out) {
        auto s = cptrs(polys);
        std::vector<u32> pts;
        for (auto& p : points)
            for (u32 w : {p.x.a, p.x.b, p.x.c, p.x.d, p.y.a, p.y.b, p.y.c, p.y.d}) pts.push_back(w);
        std::vector<u32> res(4 * polys.size());
        cm_check(cm31_eval_at_point_batch(s.data(), log_sizes.data(), s.size(), pts.data(), points.size(), point_idx.data(), res.data()));
        out.resize(polys.size());
        for (size_t i = 0; i < polys.size(); i++) out[i] = qm_make(res[4 * i], res[4 * i + 1], res[4 * i + 2], res[4 * i + 3]);
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
            for (auto* c : cols_by_layer[l]) cols.push_back(c->ptr());
        }
        start.push_back((u32)cols.size());
        cm_check(cm31_blake2s_commit_top(top_log, prev ? prev->ptr() : nullptr, cols.data(), start.data(), outp.data()));
        return out;
    }
    static const u32* col_words(const Col& c) { return c.ptr(); }
    static const u32* hash_words(const HashCol& c) { return c.ptr(); }
    static void gather_runs(const std::vector<const u32*>& srcs, const std::vector<u32>& src_id, const std::vector<u32>& word,
                            const std::vector<u32>& out_off, const std::vector<u32>& cnt, const std::vector<u32>& grid_desc,
                            const std::vector<u32>& grid_cols, const std::vector<u32>& grid_rows, std::vector<u32>& out) {
        cm_check(cm31_gather_batch(srcs.data(), srcs.size(), src_id.data(), word.data(), out_off.data(), cnt.data(), src_id.size(), grid_desc.data(),
                                   grid_desc.size() / 5, grid_cols.data(), grid_cols.size(), grid_rows.data(), grid_rows.size(), out.size(),
                                   out.data()));
    }
    static Hash32 read_root(const HashCol& root_layer) {
        Hash32 h;
        cm_check(cm31_d2h(h.b, root_layer.ptr(), 32));
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
        cm_check(cm31_accumulate_quotients(log_size, s.data(), s.size(), rc, batches.size(), pts.data(), starts.data(), idx.data(), vals.data(), o4));
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
        auto s = cptrs(cols);
        u32* a4[4] = {acc[0].ptr(), acc[1].ptr(), acc[2].ptr(), acc[3].ptr()};
        cm_check(cm31_constraint_eval(s.data(), s.size(), trace_log, eval_log, prog.code.data(), prog.code.size(), prog.n_regs,
                                      prog.consts.data(), prog.consts.size(), denom_inv.data(), a4));
    }
    static void air_program(const std::vector<const Col*>& in, const std::vector<Col*>& out, u32 log_size, const AirProgram& prog) {
        auto s = cptrs(in);
        auto d = ptrs(out);
        cm_check(cm31_air_program(s.data(), s.size(), d.data(), d.size(), log_size, prog.code.data(), prog.code.size(), prog.n_regs,
                                  prog.consts.data(), prog.consts.size()));
    }
    // multiplicity histogram: `bins` is the whole bin column, so its length bounds every looked-up value
    static void air_lookups(const std::vector<const Col*>& in, Col& bins, u32 log_size, const AirProgram& prog) {
        auto s = cptrs(in);
        u32 log_bins = 0;
        while (((size_t)1 << log_bins) < bins.size()) log_bins++;
        if (((size_t)1 << log_bins) != bins.size()) throw std::logic_error("air_lookups: the bin column is not a power of two");
        cm_check(cm31_air_lookups(s.data(), s.size(), bins.ptr(), log_bins, log_size, prog.code.data(), prog.code.size(), prog.n_regs,
                                  prog.consts.data(), prog.consts.size()));
    }
    static void check_air_errors() { cm_check(cm31_air_error_check()); }
    // Claimed sums of one interaction phase: finalize_last is stream-ordered and writes into a device
    // arena; collect_sums() reads every pending sum with ONE copy.
    struct SumArena {
        DeviceCol buf;
        size_t used = 0;
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
        cm_check(cm31_logup_finalize_last_async(l4, log_size, a.buf.ptr() + 4 * a.used));
        return a.used++;
    }
    static std::vector<QM31> collect_sums() {
        SumArena& a = sums();
        std::vector<u32> h(4 * a.used + 4);
        if (a.used) cm_check(cm31_d2h(h.data(), a.buf.ptr(), a.used * 16));
        std::vector<QM31> out;
        for (size_t i = 0; i < a.used; i++) out.push_back(qm_make(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]));
        a.used = 0;
        return out;
    }
    static QM31 logup_finalize_last(const std::array<Col*, 4>& last, u32 log_size) {
        u32* l4[4] = {last[0]->ptr(), last[1]->ptr(), last[2]->ptr(), last[3]->ptr()};
        u32 cs[4];
        cm_check(cm31_logup_finalize_last(l4, log_size, cs));
        return qm_make(cs[0], cs[1], cs[2], cs[3]);
    }
};

}  // namespace cm31

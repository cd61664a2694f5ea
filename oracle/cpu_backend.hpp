// ORACLE (test infrastructure, not product): scalar CPU restatement of Stwo's CpuBackend for the
// cairo-m proving hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline/reference legs may use it; the product library never links or calls it.
//
// PARITY PINNING: the Rust reference cannot be built in this image (no cargo/rustc),
// so this restatement is pinned against the reference's own known-answer and definitional tests
// (tests/test_oracle_kat.py): Blake2s("a") vcs/blake2_hash.rs:111-117, Blake2sChannel digests
// channel/blake2s.rs:190-224, M31 arithmetic vs %P fields/m31.rs:239-248, circle generator
// order circle.rs:186-205, evaluate == eval_at_point / interpolate∘evaluate = id
// cpu/circle.rs:295-367, eval_at_point closed forms cpu/circle.rs:257-293, quotients are low
// degree pcs/quotients.rs:177-198, fold degree tests fri.rs:1224-, Merkle verify/tamper
// vcs/blake2_merkle.rs:59-130, prefix-sum simd/prefix_sum.rs:153-187, and whole proofs verified
// by oracle/verifier.hpp.
#pragma once
#include <omp.h>

#include <algorithm>
#include <cassert>
#include <cstring>
#include <vector>

#include "fields.hpp"

namespace orc {

// ------------------------------------------------------------------ circle (core/circle.rs)
struct Point {
    M31 x, y;
};
inline Point padd(Point p, Point q) { return Point{p.x * q.x - p.y * q.y, p.x * q.y + p.y * q.x}; }
inline Point pconj(Point p) { return Point{p.x, -p.y}; }
inline Point point_from_index(u64 index) {  // CirclePointIndex::to_point, circle.rs:239
    Point res{M31(1), M31(0)}, cur{M31(2), M31(1268011823)};
    index &= 0x7fffffffull;
    while (index) {
        if (index & 1) res = padd(res, cur);
        cur = padd(cur, cur);
        index >>= 1;
    }
    return res;
}
inline M31 double_x(M31 x) { return x.square().dbl() - M31(1); }
struct QPoint {
    QM31 x, y;
};
inline QPoint qpadd(QPoint p, QPoint q) { return QPoint{p.x * q.x - p.y * q.y, p.x * q.y + p.y * q.x}; }
inline QM31 qdouble_x(QM31 x) { return x.square() + x.square() - M31(1); }

inline u32 bitrev(u32 i, u32 log_size) {  // core/utils.rs:53-58
    u32 r = 0;
    for (u32 k = 0; k < log_size; k++)
        if (i & (1u << k)) r |= 1u << (log_size - 1 - k);
    return r;
}

// Coset (circle.rs:287-330): initial index + i*step, indices mod 2^31.
struct OCoset {
    u64 initial, step;
    u32 log_size;
    static OCoset make(u64 initial, u32 log_size) { return OCoset{initial, log_size == 0 ? 0 : (1ull << (31 - log_size)), log_size}; }
    static OCoset odds(u32 log_size) { return make(1ull << (31 - (log_size + 1)), log_size); }
    static OCoset half_odds(u32 log_size) { return make(1ull << (31 - (log_size + 2)), log_size); }
    u64 index_at(u64 i) const { return (initial + step * i) & 0x7fffffffull; }
    Point at(u64 i) const { return point_from_index(index_at(i)); }
    OCoset doubled() const { return OCoset{(initial * 2) & 0x7fffffffull, (step * 2) & 0x7fffffffull, log_size - 1}; }
    size_t size() const { return (size_t)1 << log_size; }
};
// CanonicCoset(L).circle_domain() (canonic.rs:41-48, domain.rs:57-63)
struct ODomain {
    OCoset half;
    static ODomain canonic(u32 log_size) { return ODomain{OCoset::half_odds(log_size - 1)}; }
    u32 log_size() const { return half.log_size + 1; }
    size_t size() const { return (size_t)1 << log_size(); }
    u64 index_at(u64 i) const {
        u64 h = half.size();
        return i < h ? half.index_at(i) : ((1ull << 31) - half.index_at(i - h)) & 0x7fffffffull;
    }
    Point at(u64 i) const { return point_from_index(index_at(i)); }
};

// ------------------------------------------------------------------ twiddles (cpu/circle.rs:137-188)
struct OTwiddles {
    u32 log_size;  // canonic circle-domain log size of the root
    std::vector<M31> tw, itw;
};
inline OTwiddles precompute_twiddles(u32 log_size) {
    OTwiddles t;
    t.log_size = log_size;
    OCoset coset = OCoset::half_odds(log_size - 1);
    u32 k = coset.log_size;
    for (u32 lvl = 0; lvl < k; lvl++) {
        size_t i0 = t.tw.size();
        size_t half = coset.size() / 2;
        std::vector<M31> xs(half);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < half; i++) xs[i] = coset.at(i).x;
        u32 lg = coset.log_size - 1;
        t.tw.resize(i0 + half);
        for (size_t i = 0; i < half; i++) t.tw[i0 + i] = xs[bitrev((u32)i, lg)];  // bit_reverse(&mut twiddles[i0..])
        coset = coset.doubled();
    }
    t.tw.push_back(M31(1));
    t.itw.resize(t.tw.size());
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < t.tw.size(); i++) t.itw[i] = t.tw[i].inverse();
    return t;
}
// domain_line_twiddles_from_tree (poly/utils.rs:83-99): layer i of a domain with half-coset log k
inline const M31* line_twiddles(const std::vector<M31>& buf, u32 half_coset_log, u32 layer, size_t* len_out) {
    u32 i = half_coset_log - 1 - layer;  // after .rev()
    size_t len = (size_t)1 << i;
    *len_out = len;
    return buf.data() + (buf.size() - len * 2);
}

inline void butterfly(M31& v0, M31& v1, M31 t) {  // core/fft.rs:5-12
    M31 tmp = v1 * t;
    v1 = v0 - tmp;
    v0 = v0 + tmp;
}
inline void ibutterfly(M31& v0, M31& v1, M31 it) {  // core/fft.rs:14-21
    M31 tmp = v0;
    v0 = tmp + v1;
    v1 = (tmp - v1) * it;
}
template <class F>
inline void fft_layer_loop(std::vector<M31>& values, u32 i, size_t h, M31 t, F bf) {  // cpu/circle.rs:190-204
    for (size_t l = 0; l < ((size_t)1 << i); l++) {
        size_t idx0 = (h << (i + 1)) + l;
        size_t idx1 = idx0 + ((size_t)1 << i);
        bf(values[idx0], values[idx1], t);
    }
}
// One whole FFT layer: all (h, l) butterflies are independent, so a lone large column (e.g. the 2^20-row
// range-check table) still uses every host core when the caller is not already inside a parallel
// region over columns (bench.py's CPU arm is entitled to all cores).
template <class F>
inline void fft_layer(std::vector<M31>& values, u32 i, const M31* tw, size_t n_tw, F bf) {
    const size_t per = (size_t)1 << i, total = n_tw << i;
#pragma omp parallel for schedule(static) if (total >= (1u << 14) && !omp_in_parallel())
    for (size_t b = 0; b < total; b++) {
        size_t h = b >> i, l = b & (per - 1);
        size_t idx0 = (h << (i + 1)) + l;
        bf(values[idx0], values[idx0 + per], tw[h]);
    }
}
inline std::vector<M31> circle_twiddles_from_line(const M31* line0, size_t len) {  // cpu/circle.rs:209-229
    std::vector<M31> out;
    for (size_t i = 0; i + 1 < len; i += 2) {
        M31 x = line0[i], y = line0[i + 1];
        out.push_back(y);
        out.push_back(-y);
        out.push_back(-x);
        out.push_back(x);
    }
    return out;
}

// PolyOps::interpolate (cpu/circle.rs:18-71); values on CanonicCoset(L).circle_domain(), bit reversed.
inline void interpolate(std::vector<M31>& values, const OTwiddles& tw) {
    u32 L = 0;
    while (((size_t)1 << L) < values.size()) L++;
    ODomain dom = ODomain::canonic(L);
    if (L == 1) {
        M31 y = dom.half.at(0).y;
        M31 n(2);
        M31 yn_inv = (y * n).inverse();
        M31 y_inv = yn_inv * n, n_inv = yn_inv * y;
        ibutterfly(values[0], values[1], y_inv);
        values[0] = values[0] * n_inv;
        values[1] = values[1] * n_inv;
        return;
    }
    if (L == 2) {
        Point p = dom.half.at(0);
        M31 n(4);
        M31 xyn_inv = (p.x * p.y * n).inverse();
        M31 x_inv = xyn_inv * p.y * n, y_inv = xyn_inv * p.x * n, n_inv = xyn_inv * p.x * p.y;
        ibutterfly(values[0], values[1], y_inv);
        ibutterfly(values[2], values[3], -y_inv);
        ibutterfly(values[0], values[2], x_inv);
        ibutterfly(values[1], values[3], x_inv);
        for (int i = 0; i < 4; i++) values[i] = values[i] * n_inv;
        return;
    }
    u32 k = L - 1;
    size_t len0;
    const M31* l0 = line_twiddles(tw.itw, k, 0, &len0);
    std::vector<M31> ct = circle_twiddles_from_line(l0, len0);
    fft_layer(values, 0, ct.data(), ct.size(), ibutterfly);
    for (u32 layer = 0; layer < k; layer++) {
        size_t len;
        const M31* lt = line_twiddles(tw.itw, k, layer, &len);
        fft_layer(values, layer + 1, lt, len, ibutterfly);
    }
    M31 inv = M31((u64)values.size()).inverse();
    for (auto& v : values) v = v * inv;
}

// PolyOps::evaluate (cpu/circle.rs:97-135): zero-extend to the domain size, forward FFT.
inline std::vector<M31> evaluate(const std::vector<M31>& coeffs, u32 log_eval, const OTwiddles& tw) {
    std::vector<M31> values(coeffs);
    values.resize((size_t)1 << log_eval, M31());
    ODomain dom = ODomain::canonic(log_eval);
    if (log_eval == 1) {
        butterfly(values[0], values[1], dom.half.at(0).y);
        return values;
    }
    if (log_eval == 2) {
        Point p = dom.half.at(0);
        butterfly(values[0], values[2], p.x);
        butterfly(values[1], values[3], p.x);
        butterfly(values[0], values[1], p.y);
        butterfly(values[2], values[3], -p.y);
        return values;
    }
    u32 k = log_eval - 1;
    for (int layer = (int)k - 1; layer >= 0; layer--) {
        size_t len;
        const M31* lt = line_twiddles(tw.tw, k, (u32)layer, &len);
        fft_layer(values, (u32)layer + 1, lt, len, butterfly);
    }
    size_t len0;
    const M31* l0 = line_twiddles(tw.tw, k, 0, &len0);
    std::vector<M31> ct = circle_twiddles_from_line(l0, len0);
    fft_layer(values, 0, ct.data(), ct.size(), butterfly);
    return values;
}

// fold (poly/utils.rs:44-55) + eval_at_point (cpu/circle.rs:73-87)
inline QM31 fold_rec(const M31* values, size_t n, const QM31* factors) {
    if (n == 1) return QM31::from_m31(values[0]);
    QM31 l = fold_rec(values, n / 2, factors + 1);
    QM31 r = fold_rec(values + n / 2, n / 2, factors + 1);
    return l + r * factors[0];
}
inline QM31 eval_at_point(const std::vector<M31>& coeffs, QPoint p) {
    u32 L = 0;
    while (((size_t)1 << L) < coeffs.size()) L++;
    if (L == 0) return QM31::from_m31(coeffs[0]);
    std::vector<QM31> mappings;
    mappings.push_back(p.y);
    QM31 x = p.x;
    for (u32 i = 1; i < L; i++) {
        mappings.push_back(x);
        x = qdouble_x(x);
    }
    std::reverse(mappings.begin(), mappings.end());
    return fold_rec(coeffs.data(), coeffs.size(), mappings.data());
}

// ------------------------------------------------------------------ Blake2s (RFC 7693; in-tree
// restatement vcs/blake2s_ref.rs, simd/blake2s.rs:352-400)
struct OBlake2s {
    u32 h[8];
    uint8_t buf[64];
    size_t buflen;
    u64 t;
    OBlake2s() {
        static const u32 IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
        for (int i = 0; i < 8; i++) h[i] = IV[i];
        h[0] ^= 0x01010020;
        buflen = 0;
        t = 0;
    }
    static u32 rotr(u32 x, int n) { return (x >> n) | (x << (32 - n)); }
    void compress(const uint8_t block[64], bool last) {
        static const u32 IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
        static const uint8_t S[10][16] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
                                          {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
                                          {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
                                          {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
                                          {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
        u32 m[16], v[16];
        for (int i = 0; i < 16; i++) m[i] = (u32)block[4 * i] | ((u32)block[4 * i + 1] << 8) | ((u32)block[4 * i + 2] << 16) | ((u32)block[4 * i + 3] << 24);
        for (int i = 0; i < 8; i++) {
            v[i] = h[i];
            v[i + 8] = IV[i];
        }
        v[12] ^= (u32)t;
        v[13] ^= (u32)(t >> 32);
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, u32 x, u32 y) {
            v[a] = v[a] + v[b] + x;
            v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 12);
            v[a] = v[a] + v[b] + y;
            v[d] = rotr(v[d] ^ v[a], 8);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 7);
        };
        for (int r = 0; r < 10; r++) {
            const uint8_t* s = S[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]);
            G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            G(2, 6, 10, 14, m[s[4]], m[s[5]]);
            G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]);
            G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            G(2, 7, 8, 13, m[s[12]], m[s[13]]);
            G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    }
    void update(const void* data, size_t len) {
        const uint8_t* p = (const uint8_t*)data;
        while (len) {
            if (buflen == 64) {
                t += 64;
                compress(buf, false);
                buflen = 0;
            }
            size_t take = std::min(len, (size_t)64 - buflen);
            memcpy(buf + buflen, p, take);
            buflen += take;
            p += take;
            len -= take;
        }
    }
    void finalize(uint8_t out[32]) {
        t += buflen;
        memset(buf + buflen, 0, 64 - buflen);
        compress(buf, true);
        for (int i = 0; i < 8; i++) {
            out[4 * i] = (uint8_t)h[i];
            out[4 * i + 1] = (uint8_t)(h[i] >> 8);
            out[4 * i + 2] = (uint8_t)(h[i] >> 16);
            out[4 * i + 3] = (uint8_t)(h[i] >> 24);
        }
    }
};
struct Hash {
    uint8_t b[32];
    bool operator==(const Hash& o) const { return memcmp(b, o.b, 32) == 0; }
    bool operator!=(const Hash& o) const { return !(*this == o); }
};

// Blake2sMerkleHasher::hash_node (vcs/blake2_merkle.rs:14-30)
inline Hash hash_node(const Hash* left, const Hash* right, const M31* values, size_t n_values) {
    OBlake2s hs;
    if (left) {
        hs.update(left->b, 32);
        hs.update(right->b, 32);
    }
    for (size_t i = 0; i < n_values; i++) {
        uint8_t le[4] = {(uint8_t)values[i].v, (uint8_t)(values[i].v >> 8), (uint8_t)(values[i].v >> 16), (uint8_t)(values[i].v >> 24)};
        hs.update(le, 4);
    }
    Hash out;
    hs.finalize(out.b);
    return out;
}
// MerkleOps::commit_on_layer (cpu/blake2s.rs:9-23)
inline std::vector<Hash> commit_on_layer(u32 log_size, const std::vector<Hash>* prev, const std::vector<const std::vector<M31>*>& cols) {
    size_t n = (size_t)1 << log_size;
    std::vector<Hash> out(n);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        std::vector<M31> vals(cols.size());
        for (size_t c = 0; c < cols.size(); c++) vals[c] = (*cols[c])[i];
        out[i] = prev ? hash_node(&(*prev)[2 * i], &(*prev)[2 * i + 1], vals.data(), vals.size()) : hash_node(nullptr, nullptr, vals.data(), vals.size());
    }
    return out;
}

// ------------------------------------------------------------------ secure columns
struct SecureColumn {
    std::vector<M31> c[4];
    explicit SecureColumn(size_t n = 0) {
        for (auto& v : c) v.assign(n, M31());
    }
    size_t size() const { return c[0].size(); }
    QM31 at(size_t i) const { return QM31(c[0][i], c[1][i], c[2][i], c[3][i]); }
    void set(size_t i, QM31 v) {
        c[0][i] = v.x.a;
        c[1][i] = v.x.b;
        c[2][i] = v.y.a;
        c[3][i] = v.y.b;
    }
};

// ------------------------------------------------------------------ FRI folds (core/fri.rs:1132-1189)
inline void ibutterfly_q(QM31& v0, QM31& v1, M31 it) {
    QM31 tmp = v0;
    v0 = tmp + v1;
    v1 = (tmp - v1) * it;
}
// LineEvaluation on LineDomain(half_odds(log_size)), bit-reversed
inline SecureColumn fold_line(const SecureColumn& eval, QM31 alpha) {
    size_t n = eval.size();
    u32 log_size = 0;
    while (((size_t)1 << log_size) < n) log_size++;
    OCoset dom = OCoset::half_odds(log_size);
    SecureColumn out(n / 2);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n / 2; i++) {
        M31 x = dom.at(bitrev((u32)(i << 1), log_size)).x;
        QM31 f0 = eval.at(2 * i), f1 = eval.at(2 * i + 1);
        ibutterfly_q(f0, f1, x.inverse());
        out.set(i, f0 + alpha * f1);
    }
    return out;
}
inline void fold_circle_into_line(SecureColumn& dst, const SecureColumn& src, QM31 alpha) {
    size_t n = src.size();
    u32 log_size = 0;
    while (((size_t)1 << log_size) < n) log_size++;
    assert(dst.size() == n / 2);
    ODomain dom = ODomain::canonic(log_size);
    QM31 alpha_sq = alpha * alpha;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n / 2; i++) {
        Point p = dom.at(bitrev((u32)(i << 1), log_size));
        QM31 f0 = src.at(2 * i), f1 = src.at(2 * i + 1);
        ibutterfly_q(f0, f1, p.y.inverse());
        QM31 f_prime = alpha * f1 + f0;
        dst.set(i, dst.at(i) * alpha_sq + f_prime);
    }
}
// cpu/fri.rs:29-85
inline SecureColumn decompose(const SecureColumn& eval, QM31* lambda_out) {
    size_t n = eval.size(), half = n / 2;
    QM31 a = QM31::zero(), b = QM31::zero();
    for (size_t i = 0; i < half; i++) a = a + eval.at(i);
    for (size_t i = half; i < n; i++) b = b + eval.at(i);
    QM31 lambda = (a - b) * M31((u64)n).inverse();
    SecureColumn g(n);
    for (size_t i = 0; i < half; i++) g.set(i, eval.at(i) - lambda);
    for (size_t i = half; i < n; i++) g.set(i, eval.at(i) + lambda);
    *lambda_out = lambda;
    return g;
}

// ------------------------------------------------------------------ DEEP quotients (cpu/quotients.rs)
struct SampleBatch {
    QPoint point;
    std::vector<std::pair<size_t, QM31>> columns_and_values;
};
struct LineCoeffs {
    QM31 a, b, c;
};
inline LineCoeffs complex_conjugate_line_coeffs(QPoint point, QM31 value, QM31 alpha) {  // constraints.rs:98-113
    assert(point.y != point.y.complex_conjugate());
    QM31 a = value.complex_conjugate() - value;
    QM31 c = point.y.complex_conjugate() - point.y;
    QM31 b = value * c - a * point.y;
    return LineCoeffs{alpha * a, alpha * b, alpha * c};
}
struct QuotientConstants {
    std::vector<std::vector<LineCoeffs>> line_coeffs;
    std::vector<QM31> batch_random_coeffs;
};
inline QuotientConstants quotient_constants(const std::vector<SampleBatch>& batches, QM31 random_coeff) {
    QuotientConstants qc;
    for (auto& sb : batches) {
        QM31 alpha = QM31::one();
        std::vector<LineCoeffs> lc;
        for (auto& cv : sb.columns_and_values) {
            alpha = alpha * random_coeff;
            lc.push_back(complex_conjugate_line_coeffs(sb.point, cv.second, alpha));
        }
        qc.line_coeffs.push_back(lc);
        qc.batch_random_coeffs.push_back(random_coeff.pow(sb.columns_and_values.size()));
    }
    return qc;
}
inline QM31 accumulate_row_quotients(const std::vector<SampleBatch>& batches, const M31* row_values, const QuotientConstants& qc, Point dp) {
    QM31 acc = QM31::zero();
    for (size_t b = 0; b < batches.size(); b++) {
        const SampleBatch& sb = batches[b];
        CM31 prx = sb.point.x.x, pry = sb.point.y.x, pix = sb.point.x.y, piy = sb.point.y.y;
        CM31 den = (prx - CM31(dp.x, M31())) * piy - (pry - CM31(dp.y, M31())) * pix;
        CM31 den_inv = den.inverse();
        QM31 numerator = QM31::zero();
        for (size_t k = 0; k < sb.columns_and_values.size(); k++) {
            const LineCoeffs& lc = qc.line_coeffs[b][k];
            QM31 value = lc.c * row_values[sb.columns_and_values[k].first];
            QM31 linear_term = lc.a * dp.y + lc.b;
            numerator = numerator + (value - linear_term);
        }
        acc = acc * qc.batch_random_coeffs[b] + numerator.mul_cm31(den_inv);
    }
    return acc;
}
inline SecureColumn accumulate_quotients(u32 log_size, const std::vector<const std::vector<M31>*>& cols, QM31 random_coeff,
                                         const std::vector<SampleBatch>& batches) {
    ODomain dom = ODomain::canonic(log_size);
    size_t n = dom.size();
    SecureColumn out(n);
    QuotientConstants qc = quotient_constants(batches, random_coeff);
#pragma omp parallel for schedule(static)
    for (size_t row = 0; row < n; row++) {
        Point dp = dom.at(bitrev((u32)row, log_size));
        std::vector<M31> vals(cols.size());
        for (size_t c = 0; c < cols.size(); c++) vals[c] = (*cols[c])[row];
        out.set(row, accumulate_row_quotients(batches, vals.data(), qc, dp));
    }
    return out;
}

// ------------------------------------------------------------------ channel (channel/blake2s.rs:15-116)
struct OChannel {
    Hash digest;
    u32 n_sent;
    OChannel() {
        memset(digest.b, 0, 32);
        n_sent = 0;
    }
    void update_digest(const Hash& h) {
        digest = h;
        n_sent = 0;
    }
    void mix_u32s(const u32* data, size_t n) {
        OBlake2s hs;
        hs.update(digest.b, 32);
        for (size_t i = 0; i < n; i++) {
            uint8_t le[4] = {(uint8_t)data[i], (uint8_t)(data[i] >> 8), (uint8_t)(data[i] >> 16), (uint8_t)(data[i] >> 24)};
            hs.update(le, 4);
        }
        Hash h;
        hs.finalize(h.b);
        update_digest(h);
    }
    void mix_u64(u64 v) {
        u32 d[2] = {(u32)v, (u32)(v >> 32)};
        mix_u32s(d, 2);
    }
    void mix_felts(const std::vector<QM31>& felts) {
        std::vector<u32> w;
        for (auto& f : felts) {
            u32 o[4];
            f.to_u32(o);
            w.insert(w.end(), o, o + 4);
        }
        mix_u32s(w.data(), w.size());
    }
    void mix_root(const Hash& root) {  // Blake2sMerkleChannel::mix_root, vcs/blake2_merkle.rs:40-45
        OBlake2s hs;
        hs.update(digest.b, 32);
        hs.update(root.b, 32);
        Hash h;
        hs.finalize(h.b);
        update_digest(h);
    }
    Hash draw_random_bytes() {
        OBlake2s hs;
        hs.update(digest.b, 32);
        uint8_t c[4] = {(uint8_t)n_sent, (uint8_t)(n_sent >> 8), (uint8_t)(n_sent >> 16), (uint8_t)(n_sent >> 24)};
        hs.update(c, 4);
        n_sent++;
        Hash h;
        hs.finalize(h.b);
        return h;
    }
    void draw_base_felts(M31 out[8]) {
        for (;;) {
            Hash h = draw_random_bytes();
            u32 w[8];
            bool ok = true;
            for (int i = 0; i < 8; i++) {
                w[i] = (u32)h.b[4 * i] | ((u32)h.b[4 * i + 1] << 8) | ((u32)h.b[4 * i + 2] << 16) | ((u32)h.b[4 * i + 3] << 24);
                if (w[i] >= 2 * (u32)MODULUS) ok = false;
            }
            if (ok) {
                for (int i = 0; i < 8; i++) out[i] = M31((u64)w[i]);
                return;
            }
        }
    }
    QM31 draw_secure_felt() {
        M31 f[8];
        draw_base_felts(f);
        return QM31(f[0], f[1], f[2], f[3]);
    }
    std::vector<QM31> draw_secure_felts(size_t n) {
        std::vector<QM31> out;
        while (out.size() < n) {
            M31 f[8];
            draw_base_felts(f);
            out.push_back(QM31(f[0], f[1], f[2], f[3]));
            if (out.size() < n) out.push_back(QM31(f[4], f[5], f[6], f[7]));
        }
        return out;
    }
    u32 trailing_zeros() const {
        for (int i = 0; i < 16; i++) {
            if (digest.b[i]) return 8 * i + __builtin_ctz(digest.b[i]);
        }
        return 128;
    }
};
// GrindOps (cpu/grind.rs:5-16)
inline u64 grind(const OChannel& ch, u32 pow_bits) {
    for (u64 nonce = 0;; nonce++) {
        OChannel c = ch;
        c.mix_u64(nonce);
        if (c.trailing_zeros() >= pow_bits) return nonce;
    }
}

// inclusive prefix sum in coset order (simd/prefix_sum.rs:122-141 `inclusive_prefix_sum_slow`,
// index maps core/utils.rs:92-143)
inline std::vector<M31> inclusive_prefix_sum(const std::vector<M31>& bit_rev_circle_domain_evals) {
    size_t n = bit_rev_circle_domain_evals.size();
    u32 L = 0;
    while (((size_t)1 << L) < n) L++;
    std::vector<M31> nat(n), coset(n);
    for (size_t i = 0; i < n; i++) nat[bitrev((u32)i, L)] = bit_rev_circle_domain_evals[i];
    for (size_t i = 0; i < n / 2; i++) {
        coset[2 * i] = nat[i];
        coset[2 * i + 1] = nat[n - 1 - i];
    }
    M31 acc;
    for (size_t i = 0; i < n; i++) {
        acc = acc + coset[i];
        coset[i] = acc;
    }
    std::vector<M31> cd(n), out(n);
    for (size_t i = 0; i < n / 2; i++) {
        cd[i] = coset[2 * i];
        cd[n / 2 + i] = coset[n - 1 - 2 * i];
    }
    for (size_t i = 0; i < n; i++) out[i] = cd[bitrev((u32)i, L)];
    return out;
}

}  // namespace orc

// Circle group over M31 / QM31, cosets and canonic domains (host side).
//
// Mirrors external/stwo/crates/prover/src/core/circle.rs (CirclePoint :13, CirclePointIndex :216,
// Coset :287), poly/circle/canonic.rs:23 (CanonicCoset), poly/circle/domain.rs:17 (CircleDomain),
// poly/line.rs:22 (LineDomain) and core/utils.rs:53-143 (index maps).
#pragma once
#include <cassert>
#include <cstddef>
#include <vector>

#include "field.cuh"

namespace cm31 {

constexpr u32 M31_CIRCLE_LOG_ORDER = 31;
constexpr u32 M31_CIRCLE_GEN_X = 2;
constexpr u32 M31_CIRCLE_GEN_Y = 1268011823;

struct CirclePointM31 {
    u32 x, y;
};
CM_HD CirclePointM31 cp_add(CirclePointM31 p, CirclePointM31 q) {
    CirclePointM31 r;
    r.x = m31_sub(m31_mul(p.x, q.x), m31_mul(p.y, q.y));
    r.y = m31_add(m31_mul(p.x, q.y), m31_mul(p.y, q.x));
    return r;
}
CM_HD CirclePointM31 cp_double(CirclePointM31 p) { return cp_add(p, p); }
CM_HD CirclePointM31 cp_conj(CirclePointM31 p) {
    CirclePointM31 r;
    r.x = p.x;
    r.y = m31_neg(p.y);
    return r;
}
CM_HD CirclePointM31 cp_neg(CirclePointM31 p) { return cp_conj(p); }  // group inverse
CM_HD CirclePointM31 cp_sub(CirclePointM31 p, CirclePointM31 q) { return cp_add(p, cp_conj(q)); }
CM_HD CirclePointM31 cp_antipode(CirclePointM31 p) {
    CirclePointM31 r;
    r.x = m31_neg(p.x);
    r.y = m31_neg(p.y);
    return r;
}
CM_HD u32 double_x(u32 x) { return m31_sub(m31_double(m31_sqr(x)), 1); }

// index -> M31_CIRCLE_GEN * index   (circle.rs:239 `to_point`)
CM_HD CirclePointM31 cp_from_index(u32 index) {
    CirclePointM31 res = {1, 0};
    CirclePointM31 cur = {M31_CIRCLE_GEN_X, M31_CIRCLE_GEN_Y};
    index &= 0x7fffffffu;
    while (index) {
        if (index & 1) res = cp_add(res, cur);
        cur = cp_double(cur);
        index >>= 1;
    }
    return res;
}

// QM31 circle points
struct CirclePointQM31 {
    QM31 x, y;
};
CM_HD CirclePointQM31 cpq_add(CirclePointQM31 p, CirclePointQM31 q) {
    CirclePointQM31 r;
    r.x = qm_sub(qm_mul(p.x, q.x), qm_mul(p.y, q.y));
    r.y = qm_add(qm_mul(p.x, q.y), qm_mul(p.y, q.x));
    return r;
}
CM_HD CirclePointQM31 cpq_from_m31(CirclePointM31 p) {
    CirclePointQM31 r;
    r.x = qm_from_m31(p.x);
    r.y = qm_from_m31(p.y);
    return r;
}
CM_HD QM31 qm_double_x(QM31 x) {
    QM31 sx = qm_sqr(x);
    return qm_sub_m31(qm_add(sx, sx), 1);
}
CM_HD bool cpq_eq(CirclePointQM31 p, CirclePointQM31 q) { return qm_eq(p.x, q.x) && qm_eq(p.y, q.y); }

inline u32 idx_reduce(u64 v) { return (u32)(v & 0x7fffffffu); }
inline u32 idx_add(u32 a, u32 b) { return idx_reduce((u64)a + b); }
inline u32 idx_neg(u32 a) { return idx_reduce((1ull << 31) - a); }
inline u32 idx_sub(u32 a, u32 b) { return idx_add(a, idx_neg(b)); }
inline u32 idx_mul(u32 a, u64 k) { return idx_reduce((u64)a * k); }
inline u32 subgroup_gen(u32 log_size) {
    assert(log_size <= M31_CIRCLE_LOG_ORDER);
    return log_size == 0 ? 0u : (1u << (M31_CIRCLE_LOG_ORDER - log_size));
}

struct Coset {
    u32 initial_index = 0;
    u32 step_size = 0;
    u32 log_size = 0;

    static Coset make(u32 initial_index, u32 log_size) {
        Coset c;
        c.initial_index = initial_index;
        c.step_size = subgroup_gen(log_size);
        c.log_size = log_size;
        return c;
    }
    static Coset subgroup(u32 log_size) { return make(0, log_size); }
    static Coset odds(u32 log_size) { return make(subgroup_gen(log_size + 1), log_size); }
    static Coset half_odds(u32 log_size) { return make(subgroup_gen(log_size + 2), log_size); }
    size_t size() const { return (size_t)1 << log_size; }
    u32 index_at(size_t i) const { return idx_add(initial_index, idx_mul(step_size, i)); }
    CirclePointM31 at(size_t i) const { return cp_from_index(index_at(i)); }
    CirclePointM31 initial() const { return cp_from_index(initial_index); }
    CirclePointM31 step() const { return cp_from_index(step_size); }
    Coset doubled() const {
        assert(log_size > 0);
        Coset c;
        c.initial_index = idx_mul(initial_index, 2);
        c.step_size = idx_mul(step_size, 2);
        c.log_size = log_size - 1;
        return c;
    }
    Coset repeated_double(u32 n) const {
        Coset c = *this;
        for (u32 i = 0; i < n; i++) c = c.doubled();
        return c;
    }
    bool operator==(const Coset& o) const {
        return initial_index == o.initial_index && step_size == o.step_size && log_size == o.log_size;
    }
    bool is_doubling_of(const Coset& other) const {
        return log_size <= other.log_size && *this == other.repeated_double(other.log_size - log_size);
    }
    Coset conjugate() const {
        Coset c;
        c.initial_index = idx_neg(initial_index);
        c.step_size = idx_neg(step_size);
        c.log_size = log_size;
        return c;
    }
    Coset shifted(u32 shift) const {
        Coset c = *this;
        c.initial_index = idx_add(initial_index, shift);
        return c;
    }
};

struct CircleDomain {
    Coset half_coset;
    u32 log_size() const { return half_coset.log_size + 1; }
    size_t size() const { return (size_t)1 << log_size(); }
    u32 index_at(size_t i) const {
        size_t h = half_coset.size();
        return i < h ? half_coset.index_at(i) : idx_neg(half_coset.index_at(i - h));
    }
    CirclePointM31 at(size_t i) const { return cp_from_index(index_at(i)); }
    bool is_canonic() const { return idx_mul(half_coset.initial_index, 4) == half_coset.step_size; }
};

struct CanonicCoset {
    Coset coset;
    explicit CanonicCoset(u32 log_size) {
        assert(log_size > 0);
        coset = Coset::odds(log_size);
    }
    u32 log_size() const { return coset.log_size; }
    Coset half_coset() const { return Coset::half_odds(log_size() - 1); }
    CircleDomain circle_domain() const {
        CircleDomain d;
        d.half_coset = half_coset();
        return d;
    }
    CirclePointM31 step() const { return coset.step(); }
};

struct LineDomain {
    Coset coset;
    u32 log_size() const { return coset.log_size; }
    size_t size() const { return coset.size(); }
    u32 at(size_t i) const { return coset.at(i).x; }
    LineDomain doubled() const {
        LineDomain d;
        d.coset = coset.doubled();
        return d;
    }
};

// core/utils.rs:74-90
inline size_t offset_bit_reversed_circle_domain_index(size_t i, u32 domain_log_size, u32 eval_log_size,
                                                      long offset) {
    long prev_index = (long)bit_reverse((u32)i, eval_log_size);
    long half_size = 1l << (eval_log_size - 1);
    long step_size = offset * (1l << (eval_log_size - domain_log_size - 1));
    auto rem_euclid = [](long a, long m) {
        long r = a % m;
        return r < 0 ? r + m : r;
    };
    if (prev_index < half_size) {
        prev_index = rem_euclid(prev_index + step_size, half_size);
    } else {
        prev_index = rem_euclid(prev_index - step_size, half_size) + half_size;
    }
    return bit_reverse((u32)prev_index, eval_log_size);
}
inline size_t circle_domain_index_to_coset_index(size_t circle_index, u32 log_domain_size) {
    size_t n = (size_t)1 << log_domain_size;
    return circle_index < n / 2 ? circle_index * 2 : (n - 1 - circle_index) * 2 + 1;
}
inline size_t coset_index_to_circle_domain_index(size_t coset_index, u32 log_domain_size) {
    return coset_index % 2 == 0 ? coset_index / 2 : (((size_t)2 << log_domain_size) - coset_index) / 2;
}

// constraints.rs:11-34 — vanishing polynomial of a coset, at an M31 point.
inline u32 coset_vanishing_m31(const Coset& coset, CirclePointM31 p) {
    p = cp_add(cp_sub(p, coset.initial()), cp_from_index(coset.step_size >> 1));
    u32 x = p.x;
    for (u32 i = 1; i < coset.log_size; i++) x = double_x(x);
    return x;
}
inline QM31 coset_vanishing_qm31(const Coset& coset, CirclePointQM31 p) {
    p = cpq_add(p, cpq_from_m31(cp_conj(coset.initial())));
    p = cpq_add(p, cpq_from_m31(cp_from_index(coset.step_size >> 1)));
    QM31 x = p.x;
    for (u32 i = 1; i < coset.log_size; i++) x = qm_double_x(x);
    return x;
}

}  // namespace cm31

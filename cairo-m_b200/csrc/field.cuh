// M31 / CM31 / QM31 arithmetic, host + device.
//
// Mirrors the value semantics of the reference field tower
//   M31   : external/stwo/crates/prover/src/core/fields/m31.rs:32   (P = 2^31-1)
//   CM31  : external/stwo/crates/prover/src/core/fields/cm31.rs:44  (M31[i]/(i^2+1))
//   QM31  : external/stwo/crates/prover/src/core/fields/qm31.rs:78  (CM31[u]/(u^2-2-i))
// Every value is kept CANONICAL in [0, P), so results equal the
// reference CpuBackend bit for bit.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define CM_HD __host__ __device__ __forceinline__
#else
#define CM_HD inline
#endif

namespace cm31 {

typedef uint32_t u32;
typedef uint64_t u64;

constexpr u32 P = 0x7fffffffu;

// ---------------------------------------------------------------- M31 (raw u32 functions)
CM_HD u32 m31_add(u32 a, u32 b) {
    u32 s = a + b;  // < 2P < 2^32
    u32 t = s - P;
    return t < s ? t : s;  // umin(s, s-P): s-P wraps above s when s < P
}
CM_HD u32 m31_sub(u32 a, u32 b) {
    u32 d = a - b;
    u32 t = d + P;
    return t < d ? t : d;  // if a<b, d wrapped (huge) and d+P is the answer
}
CM_HD u32 m31_neg(u32 a) { return a == 0 ? 0 : P - a; }
CM_HD u32 m31_reduce64(u64 x) {
    // any u64: 2^31 == 1 (mod P), fold twice then one conditional subtract.
    u64 y = (x & P) + (x >> 31);             // < 2^34
    u32 z = (u32)(y & P) + (u32)(y >> 31);   // <= P + 7
    return z >= P ? z - P : z;
}
CM_HD u32 m31_mul(u32 a, u32 b) {
    u64 p = (u64)a * b;  // < 2^62
    u32 lo = (u32)p & P;
    u32 hi = (u32)(p >> 31);  // < 2^31 - 1  (since a,b <= P-1)
    u32 s = lo + hi;
    u32 t = s - P;
    return t < s ? t : s;
}
CM_HD u32 m31_sqr(u32 a) { return m31_mul(a, a); }
CM_HD u32 m31_double(u32 a) { return m31_add(a, a); }
CM_HD u32 m31_sqn(u32 v, int n) {
    for (int i = 0; i < n; i++) v = m31_sqr(v);
    return v;
}
// v^(P-2); same addition chain as fields/m31.rs:197-205 (37 multiplications).
CM_HD u32 m31_inv(u32 v) {
    u32 t0 = m31_mul(m31_sqn(v, 2), v);
    u32 t1 = m31_mul(m31_sqn(t0, 1), t0);
    u32 t2 = m31_mul(m31_sqn(t1, 3), t0);
    u32 t3 = m31_mul(m31_sqn(t2, 1), t0);
    u32 t4 = m31_mul(m31_sqn(t3, 8), t3);
    u32 t5 = m31_mul(m31_sqn(t4, 8), t3);
    return m31_mul(m31_sqn(t5, 7), t2);
}
CM_HD u32 m31_from_i64(long long v) {
    long long r = v % (long long)P;
    if (r < 0) r += P;
    return (u32)r;
}

// ---------------------------------------------------------------- CM31
struct CM31 {
    u32 a, b;
};
CM_HD CM31 cm_make(u32 a, u32 b) {
    CM31 r;
    r.a = a;
    r.b = b;
    return r;
}
CM_HD CM31 cm_add(CM31 x, CM31 y) { return cm_make(m31_add(x.a, y.a), m31_add(x.b, y.b)); }
CM_HD CM31 cm_sub(CM31 x, CM31 y) { return cm_make(m31_sub(x.a, y.a), m31_sub(x.b, y.b)); }
CM_HD CM31 cm_neg(CM31 x) { return cm_make(m31_neg(x.a), m31_neg(x.b)); }
CM_HD CM31 cm_mul(CM31 x, CM31 y) {
    // (a+bi)(c+di) = (ac-bd) + (ad+bc)i, accumulated in u64 before one reduction.
    u64 ac = (u64)x.a * y.a, bd = (u64)x.b * y.b;
    u64 ad = (u64)x.a * y.b, bc = (u64)x.b * y.a;
    u32 re = m31_sub(m31_reduce64(ac), m31_reduce64(bd));
    u32 im = m31_reduce64(ad + bc);  // < 2^63
    return cm_make(re, im);
}
CM_HD CM31 cm_mul_m31(CM31 x, u32 s) { return cm_make(m31_mul(x.a, s), m31_mul(x.b, s)); }
CM_HD CM31 cm_sqr(CM31 x) { return cm_mul(x, x); }
CM_HD CM31 cm_inv(CM31 x) {
    u32 n = m31_add(m31_sqr(x.a), m31_sqr(x.b));
    u32 ni = m31_inv(n);
    return cm_make(m31_mul(x.a, ni), m31_mul(m31_neg(x.b), ni));
}

// ---------------------------------------------------------------- QM31
struct QM31 {
    u32 a, b, c, d;  // (a + b i) + (c + d i) u
};
CM_HD QM31 qm_make(u32 a, u32 b, u32 c, u32 d) {
    QM31 r;
    r.a = a;
    r.b = b;
    r.c = c;
    r.d = d;
    return r;
}
// Witness helper of the u32 DivRem components (crates/prover/src/components/opcodes/u32_store_div_fp_fp.rs:360-400):
// v = (n_lo, n_hi, d_lo, d_hi) as 16-bit limbs; part 0..3 = q_lo, q_hi, r_lo, r_hi of n / d, (0, 0) when d == 0.
CM_HD u32 u32_divrem_part(QM31 v, u32 part) {
    u32 n = v.a | (v.b << 16), d = v.c | (v.d << 16);
    u32 q = d == 0 ? 0u : n / d, r = d == 0 ? 0u : n % d;
    u32 x = part < 2 ? q : r;
    return (part & 1u) ? x >> 16 : x & 0xffffu;
}
CM_HD QM31 qm_zero() { return qm_make(0, 0, 0, 0); }
CM_HD QM31 qm_one() { return qm_make(1, 0, 0, 0); }
CM_HD QM31 qm_from_m31(u32 v) { return qm_make(v, 0, 0, 0); }
CM_HD CM31 qm_lo(QM31 x) { return cm_make(x.a, x.b); }
CM_HD CM31 qm_hi(QM31 x) { return cm_make(x.c, x.d); }
CM_HD QM31 qm_from_cm(CM31 lo, CM31 hi) { return qm_make(lo.a, lo.b, hi.a, hi.b); }
CM_HD bool qm_eq(QM31 x, QM31 y) { return x.a == y.a && x.b == y.b && x.c == y.c && x.d == y.d; }
CM_HD bool qm_is_zero(QM31 x) { return (x.a | x.b | x.c | x.d) == 0; }
CM_HD QM31 qm_add(QM31 x, QM31 y) {
    return qm_make(m31_add(x.a, y.a), m31_add(x.b, y.b), m31_add(x.c, y.c), m31_add(x.d, y.d));
}
CM_HD QM31 qm_sub(QM31 x, QM31 y) {
    return qm_make(m31_sub(x.a, y.a), m31_sub(x.b, y.b), m31_sub(x.c, y.c), m31_sub(x.d, y.d));
}
CM_HD QM31 qm_neg(QM31 x) { return qm_make(m31_neg(x.a), m31_neg(x.b), m31_neg(x.c), m31_neg(x.d)); }
CM_HD QM31 qm_add_m31(QM31 x, u32 s) { return qm_make(m31_add(x.a, s), x.b, x.c, x.d); }
CM_HD QM31 qm_sub_m31(QM31 x, u32 s) { return qm_make(m31_sub(x.a, s), x.b, x.c, x.d); }
CM_HD QM31 qm_mul_m31(QM31 x, u32 s) {
    return qm_make(m31_mul(x.a, s), m31_mul(x.b, s), m31_mul(x.c, s), m31_mul(x.d, s));
}
CM_HD QM31 qm_mul_cm31(QM31 x, CM31 s) { return qm_from_cm(cm_mul(qm_lo(x), s), cm_mul(qm_hi(x), s)); }
CM_HD QM31 qm_mul(QM31 x, QM31 y) {
    // (A + Bu)(C + Du) = (AC + R*BD) + (AD + BC)u,   R = 2 + i   (qm31.rs:14,78-87)
    CM31 A = qm_lo(x), B = qm_hi(x), C = qm_lo(y), D = qm_hi(y);
    CM31 ac = cm_mul(A, C), bd = cm_mul(B, D);
    // R*bd = (2+i)(p+qi) = (2p - q) + (p + 2q) i
    CM31 rbd = cm_make(m31_sub(m31_double(bd.a), bd.b), m31_add(bd.a, m31_double(bd.b)));
    CM31 lo = cm_add(ac, rbd);
    CM31 hi = cm_add(cm_mul(A, D), cm_mul(B, C));
    return qm_from_cm(lo, hi);
}
CM_HD QM31 qm_sqr(QM31 x) { return qm_mul(x, x); }
CM_HD QM31 qm_inv(QM31 x) {
    // (A + Bu)^-1 = (A - Bu) / (A^2 - (2+i) B^2)     (qm31.rs:119-129)
    CM31 A = qm_lo(x), B = qm_hi(x);
    CM31 b2 = cm_sqr(B);
    CM31 ib2 = cm_make(m31_neg(b2.b), b2.a);
    CM31 den = cm_sub(cm_sqr(A), cm_add(cm_add(b2, b2), ib2));
    CM31 di = cm_inv(den);
    return qm_from_cm(cm_mul(A, di), cm_mul(cm_neg(B), di));
}
// Conjugation over CM31 (u -> -u); `ComplexConjugate` in fields/mod.rs:411.
CM_HD QM31 qm_conj(QM31 x) { return qm_make(x.a, x.b, m31_neg(x.c), m31_neg(x.d)); }
CM_HD QM31 qm_pow(QM31 x, u64 e) {
    QM31 r = qm_one();
    while (e) {
        if (e & 1) r = qm_mul(r, x);
        x = qm_sqr(x);
        e >>= 1;
    }
    return r;
}
// from_partial_evals (qm31.rs:51-57): e0 + e1*i + e2*u + e3*iu
CM_HD QM31 qm_from_partial_evals(QM31 e0, QM31 e1, QM31 e2, QM31 e3) {
    QM31 r = e0;
    r = qm_add(r, qm_mul(e1, qm_make(0, 1, 0, 0)));
    r = qm_add(r, qm_mul(e2, qm_make(0, 0, 1, 0)));
    r = qm_add(r, qm_mul(e3, qm_make(0, 0, 0, 1)));
    return r;
}

CM_HD u32 bit_reverse(u32 i, u32 log_size) {
    if (log_size == 0) return i;
#if defined(__CUDA_ARCH__)
    return __brev(i) >> (32 - log_size);
#else
    u32 r = 0;
    for (u32 k = 0; k < log_size; k++) r |= ((i >> k) & 1u) << (log_size - 1 - k);
    return r;
#endif
}

}  // namespace cm31

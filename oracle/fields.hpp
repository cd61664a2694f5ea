// ORACLE (test infrastructure, not product): plain scalar restatement of the reference field
// tower, written independently of cairo-m_b200/csrc/field.cuh (everything goes through `% P`).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may use it.
//
// Follows external/stwo/crates/prover/src/core/fields/m31.rs:32-120 (M31, reduce :58),
// cm31.rs:44-73 (CM31 mul/inverse), qm31.rs:14-129 (QM31, R = 2+i, mul, inverse),
// fields/mod.rs:405-414 (complex_conjugate).
#pragma once
#include <cstdint>
#include <cstdlib>

namespace orc {

typedef uint32_t u32;
typedef uint64_t u64;
static const u64 MODULUS = 2147483647ull;  // 2^31 - 1

struct M31 {
    u32 v;
    M31() : v(0) {}
    explicit M31(u64 x) : v((u32)(x % MODULUS)) {}
    static M31 from_i64(long long x) {
        long long r = x % (long long)MODULUS;
        if (r < 0) r += (long long)MODULUS;
        return M31((u64)r);
    }
    M31 operator+(M31 o) const { return M31((u64)v + o.v); }
    M31 operator-(M31 o) const { return M31((u64)v + MODULUS - o.v); }
    M31 operator*(M31 o) const { return M31((u64)v * o.v); }
    M31 operator-() const { return M31(MODULUS - v); }
    bool operator==(M31 o) const { return v == o.v; }
    bool operator!=(M31 o) const { return v != o.v; }
    M31 pow(u64 e) const {
        M31 r(1), b = *this;
        while (e) {
            if (e & 1) r = r * b;
            b = b * b;
            e >>= 1;
        }
        return r;
    }
    M31 inverse() const {
        if (v == 0) abort();  // "0 has no inverse"
        return pow(MODULUS - 2);
    }
    M31 square() const { return *this * *this; }
    M31 dbl() const { return *this + *this; }
};

struct CM31 {
    M31 a, b;
    CM31() {}
    CM31(M31 a_, M31 b_) : a(a_), b(b_) {}
    CM31 operator+(CM31 o) const { return CM31(a + o.a, b + o.b); }
    CM31 operator-(CM31 o) const { return CM31(a - o.a, b - o.b); }
    CM31 operator-() const { return CM31(-a, -b); }
    CM31 operator*(CM31 o) const { return CM31(a * o.a - b * o.b, a * o.b + b * o.a); }
    CM31 operator*(M31 o) const { return CM31(a * o, b * o); }
    bool operator==(CM31 o) const { return a == o.a && b == o.b; }
    CM31 square() const { return *this * *this; }
    CM31 inverse() const { return CM31(a, -b) * (a.square() + b.square()).inverse(); }
};

struct QM31 {
    CM31 x, y;  // x + y u
    QM31() {}
    QM31(CM31 x_, CM31 y_) : x(x_), y(y_) {}
    QM31(M31 a, M31 b, M31 c, M31 d) : x(a, b), y(c, d) {}
    static QM31 from_u32(u32 a, u32 b, u32 c, u32 d) { return QM31(M31(a), M31(b), M31(c), M31(d)); }
    static QM31 from_m31(M31 a) { return QM31(a, M31(), M31(), M31()); }
    static QM31 zero() { return QM31(); }
    static QM31 one() { return from_u32(1, 0, 0, 0); }
    QM31 operator+(QM31 o) const { return QM31(x + o.x, y + o.y); }
    QM31 operator-(QM31 o) const { return QM31(x - o.x, y - o.y); }
    QM31 operator-() const { return QM31(-x, -y); }
    QM31 operator*(QM31 o) const {
        CM31 R(M31(2), M31(1));
        return QM31(x * o.x + R * y * o.y, x * o.y + y * o.x);
    }
    QM31 operator*(M31 o) const { return QM31(x * o, y * o); }
    QM31 operator+(M31 o) const { return QM31(CM31(x.a + o, x.b), y); }
    QM31 operator-(M31 o) const { return QM31(CM31(x.a - o, x.b), y); }
    QM31 mul_cm31(CM31 o) const { return QM31(x * o, y * o); }
    bool operator==(QM31 o) const { return x == o.x && y == o.y; }
    bool operator!=(QM31 o) const { return !(*this == o); }
    bool is_zero() const { return x.a.v == 0 && x.b.v == 0 && y.a.v == 0 && y.b.v == 0; }
    QM31 square() const { return *this * *this; }
    QM31 inverse() const {
        if (is_zero()) abort();
        CM31 b2 = y.square();
        CM31 ib2(-b2.b, b2.a);
        CM31 denom = x.square() - (b2 + b2 + ib2);
        CM31 di = denom.inverse();
        return QM31(x * di, -y * di);
    }
    QM31 complex_conjugate() const { return QM31(x, -y); }
    QM31 pow(u64 e) const {
        QM31 r = one(), b = *this;
        while (e) {
            if (e & 1) r = r * b;
            b = b * b;
            e >>= 1;
        }
        return r;
    }
    void to_u32(u32 out[4]) const {
        out[0] = x.a.v;
        out[1] = x.b.v;
        out[2] = y.a.v;
        out[3] = y.b.v;
    }
    static QM31 from_partial_evals(QM31 e0, QM31 e1, QM31 e2, QM31 e3) {
        return e0 + e1 * from_u32(0, 1, 0, 0) + e2 * from_u32(0, 0, 1, 0) + e3 * from_u32(0, 0, 0, 1);
    }
};

}  // namespace orc

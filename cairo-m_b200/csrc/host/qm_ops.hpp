// Operator sugar for cm31::QM31 in host-side protocol code.
#pragma once
#include <array>
#include <cstring>
#include <vector>

#include "../circle.hpp"
#include "../field.cuh"

namespace cm31 {

CM_HD QM31 operator+(QM31 a, QM31 b) { return qm_add(a, b); }
CM_HD QM31 operator-(QM31 a, QM31 b) { return qm_sub(a, b); }
CM_HD QM31 operator*(QM31 a, QM31 b) { return qm_mul(a, b); }
CM_HD QM31 operator-(QM31 a) { return qm_neg(a); }
CM_HD bool operator==(QM31 a, QM31 b) { return qm_eq(a, b); }
CM_HD bool operator!=(QM31 a, QM31 b) { return !qm_eq(a, b); }

struct Hash32 {
    uint8_t b[32];
    bool operator==(const Hash32& o) const { return memcmp(b, o.b, 32) == 0; }
    bool operator!=(const Hash32& o) const { return !(*this == o); }
};

typedef CirclePointQM31 SecurePoint;

// CirclePoint<SecureField> + CirclePoint<M31>.into_ef()
inline SecurePoint secure_point_add_m31(SecurePoint p, CirclePointM31 q) { return cpq_add(p, cpq_from_m31(q)); }
// CirclePoint::mul_signed on the M31 circle (circle.rs:104-110)
inline CirclePointM31 cp_mul_signed(CirclePointM31 p, long off) {
    CirclePointM31 base = off >= 0 ? p : cp_conj(p);
    unsigned long k = off >= 0 ? (unsigned long)off : (unsigned long)(-off);
    CirclePointM31 res = {1, 0};
    while (k) {
        if (k & 1) res = cp_add(res, base);
        base = cp_double(base);
        k >>= 1;
    }
    return res;
}

}  // namespace cm31

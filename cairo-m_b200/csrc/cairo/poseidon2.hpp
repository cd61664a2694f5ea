// Poseidon2 over M31, state size 16, 8 full + 14 partial rounds, S-box x^5: the hash of the memory commitment
// (crates/prover/src/poseidon2.rs:10-57, crates/prover/src/components/poseidon2.rs:70-140).
//
// Constants: the reference takes its round constants and internal diagonal from `zkhash 0.2.0`
// (git+https://github.com/AntoineFONDEUR/poseidon2?branch=poseidon2-M31#5f715d0c:
// poseidon2_instance_m31::{RC16, MAT_DIAG16_M_1}, extracted by crates/prover/build.rs:26-106), a dependency that is not
// vendored in the reference checkout.  `poseidon2_constants.hpp` holds those tables re-derived with the Poseidon2
// parameter script's Grain LFSR (tools/gen_poseidon2_constants.py); they reproduce the reference known-answer test
// (crates/prover/tests/poseidon2.rs:15-35), which is a hard test in tests/test_oracle_cairo.py.
#pragma once
#include <array>
#include <cstdint>

#include "../field.cuh"
#include "poseidon2_constants.hpp"

namespace cm31 {

constexpr int POSEIDON2_T = 16, POSEIDON2_FULL_ROUNDS = 8, POSEIDON2_PARTIAL_ROUNDS = 14;

struct Poseidon2Constants {
    u32 external[POSEIDON2_FULL_ROUNDS][POSEIDON2_T];  // EXTERNAL_ROUND_CONSTS (first half, then second half)
    u32 internal[POSEIDON2_PARTIAL_ROUNDS];            // INTERNAL_ROUND_CONSTS
    u32 diag[POSEIDON2_T];                             // INTERNAL_MATRIX (the diagonal of M_I - 1)
    Poseidon2Constants() {
        for (int r = 0; r < POSEIDON2_FULL_ROUNDS; r++)
            for (int i = 0; i < POSEIDON2_T; i++) external[r][i] = POSEIDON2_EXTERNAL_RC[r][i];
        for (int r = 0; r < POSEIDON2_PARTIAL_ROUNDS; r++) internal[r] = POSEIDON2_INTERNAL_RC[r];
        for (int i = 0; i < POSEIDON2_T; i++) diag[i] = POSEIDON2_INTERNAL_DIAG[i];
    }
};
inline const Poseidon2Constants& poseidon2_constants() {
    static const Poseidon2Constants c;
    return c;
}

// The permutation written once over an abstract field element type F (M31 values on the host, AIR expressions in
// evaluate / write_trace): `mulc(x, c)` = x * constant, `addc(x, c)` = x + constant.
template <class F>
inline void poseidon2_apply_m4(F& a, F& b, F& c, F& d) {  // components/poseidon2.rs:70-87
    F t0 = a + b;
    F t02 = t0 + t0;
    F t1 = c + d;
    F t12 = t1 + t1;
    F t2 = b + b + t1;
    F t3 = d + d + t0;
    F t4 = t12 + t12 + t3;
    F t5 = t02 + t02 + t2;
    F t6 = t3 + t5;
    F t7 = t2 + t4;
    a = t6;
    b = t5;
    c = t7;
    d = t4;
}
template <class F>
inline void poseidon2_external_matrix(std::array<F, 16>& st) {  // circ(2 M4, M4, M4, M4), components/poseidon2.rs:89-113
    for (int i = 0; i < 4; i++) poseidon2_apply_m4(st[4 * i], st[4 * i + 1], st[4 * i + 2], st[4 * i + 3]);
    for (int j = 0; j < 4; j++) {
        F s = st[j] + st[j + 4] + st[j + 8] + st[j + 12];
        for (int i = 0; i < 4; i++) st[4 * i + j] = st[4 * i + j] + s;
    }
}
template <class F, class MulC>
inline void poseidon2_internal_matrix(std::array<F, 16>& st, MulC mulc) {  // components/poseidon2.rs:117-128
    F sum = st[0];
    for (int i = 1; i < 16; i++) sum = sum + st[i];
    const Poseidon2Constants& k = poseidon2_constants();
    for (int i = 0; i < 16; i++) st[i] = mulc(st[i], k.diag[i]) + sum;
}

// host permutation on canonical M31 values
struct M31Val {
    u32 v;
};
inline M31Val operator+(M31Val a, M31Val b) { return M31Val{m31_add(a.v, b.v)}; }
inline M31Val operator*(M31Val a, M31Val b) { return M31Val{m31_mul(a.v, b.v)}; }
inline std::array<u32, 16> poseidon2_permutation(const std::array<u32, 16>& input) {
    const Poseidon2Constants& k = poseidon2_constants();
    std::array<M31Val, 16> st;
    for (int i = 0; i < 16; i++) st[i] = M31Val{input[i] % P};
    auto full_round = [&](int r) {
        for (int i = 0; i < 16; i++) {
            M31Val x{m31_add(st[i].v, k.external[r][i])};
            M31Val x2 = x * x, x4 = x2 * x2;
            st[i] = x4 * x;
        }
        poseidon2_external_matrix(st);
    };
    poseidon2_external_matrix(st);
    for (int r = 0; r < POSEIDON2_FULL_ROUNDS / 2; r++) full_round(r);
    for (int r = 0; r < POSEIDON2_PARTIAL_ROUNDS; r++) {
        M31Val x{m31_add(st[0].v, k.internal[r])};
        M31Val x2 = x * x, x4 = x2 * x2;
        st[0] = x4 * x;
        poseidon2_internal_matrix(st, [](M31Val a, u32 c) { return M31Val{m31_mul(a.v, c)}; });
    }
    for (int r = 0; r < POSEIDON2_FULL_ROUNDS / 2; r++) full_round(POSEIDON2_FULL_ROUNDS / 2 + r);
    std::array<u32, 16> out;
    for (int i = 0; i < 16; i++) out[i] = st[i].v;
    return out;
}
// Poseidon2Hash::hash (poseidon2.rs:25-32): state = [left, right, 0, ...], digest = first element
inline u32 poseidon2_hash(u32 left, u32 right) {
    std::array<u32, 16> in{};
    in[0] = left;
    in[1] = right;
    return poseidon2_permutation(in)[0];
}

}  // namespace cm31

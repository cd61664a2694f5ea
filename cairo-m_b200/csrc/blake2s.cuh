// Blake2s-256 (RFC 7693), host + device, word oriented.
//
// The reference hashes through the RustCrypto `blake2 0.10.6` crate (vcs/blake2_merkle.rs:14-30,
// vcs/blake2_hash.rs) and restates the compression in-tree in backend/simd/blake2s.rs:352-400;
// this is an independent implementation of the same RFC used by the Merkle kernels, the grind
// kernel and the host-side Fiat-Shamir channel.
#pragma once
#include <cstdint>
#include <cstring>

#include "field.cuh"

namespace cm31 {

struct Blake2sState {
    u32 h[8];
};

#if defined(__CUDA_ARCH__)
#define CM_ROTR(x, n) __funnelshift_r((x), (x), (n))
#else
#define CM_ROTR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
#endif

CM_HD void blake2s_init(Blake2sState& s) {
    s.h[0] = 0x6A09E667u ^ 0x01010020u;  // digest 32 bytes, no key, fanout=depth=1
    s.h[1] = 0xBB67AE85u;
    s.h[2] = 0x3C6EF372u;
    s.h[3] = 0xA54FF53Au;
    s.h[4] = 0x510E527Fu;
    s.h[5] = 0x9B05688Cu;
    s.h[6] = 0x1F83D9ABu;
    s.h[7] = 0x5BE0CD19u;
}

// Pipe balance on sm_100a: xor/rotate (LOP3, SHF, PRMT) can only issue on the ALU pipe, which is
// the bottleneck of Blake2s (ncu r01: alu 92-96 %, fma 9 %).  Additions written as x*1+y with an
// opaque 1 (a __constant__ word ptxas cannot fold) become IMAD on the otherwise idle FMA pipe:
// per G 8 ALU + 6 FMA issue slots instead of 12 ALU.
#if defined(__CUDACC__)
static __constant__ u32 CM_BLAKE_ONE = 1;
#endif
#if defined(__CUDA_ARCH__)
#define CM_ADD(x, y) ((x) * CM_BLAKE_ONE + (y))
#else
#define CM_ADD(x, y) ((x) + (y))
#endif

#define CM_G(a, b, c, d, x, y)        \
    a = CM_ADD(x, CM_ADD(b, a));      \
    d = CM_ROTR(d ^ a, 16);           \
    c = CM_ADD(d, c);                 \
    b = CM_ROTR(b ^ c, 12);           \
    a = CM_ADD(y, CM_ADD(b, a));      \
    d = CM_ROTR(d ^ a, 8);            \
    c = CM_ADD(d, c);                 \
    b = CM_ROTR(b ^ c, 7);

#define CM_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    CM_G(v0, v4, v8, v12, m[s0], m[s1])                                                  \
    CM_G(v1, v5, v9, v13, m[s2], m[s3])                                                  \
    CM_G(v2, v6, v10, v14, m[s4], m[s5])                                                 \
    CM_G(v3, v7, v11, v15, m[s6], m[s7])                                                 \
    CM_G(v0, v5, v10, v15, m[s8], m[s9])                                                 \
    CM_G(v1, v6, v11, v12, m[s10], m[s11])                                               \
    CM_G(v2, v7, v8, v13, m[s12], m[s13])                                                \
    CM_G(v3, v4, v9, v14, m[s14], m[s15])

// One compression. `t` = total bytes hashed so far INCLUDING this block; `last` sets f0.
CM_HD void blake2s_compress(Blake2sState& s, const u32 m[16], u64 t, bool last) {
    u32 v0 = s.h[0], v1 = s.h[1], v2 = s.h[2], v3 = s.h[3];
    u32 v4 = s.h[4], v5 = s.h[5], v6 = s.h[6], v7 = s.h[7];
    u32 v8 = 0x6A09E667u, v9 = 0xBB67AE85u, v10 = 0x3C6EF372u, v11 = 0xA54FF53Au;
    u32 v12 = 0x510E527Fu ^ (u32)t, v13 = 0x9B05688Cu ^ (u32)(t >> 32);
    u32 v14 = last ? ~0x1F83D9ABu : 0x1F83D9ABu, v15 = 0x5BE0CD19u;
    CM_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    CM_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    CM_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    CM_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    CM_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    CM_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    CM_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    CM_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    CM_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    CM_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
    s.h[0] ^= v0 ^ v8;
    s.h[1] ^= v1 ^ v9;
    s.h[2] ^= v2 ^ v10;
    s.h[3] ^= v3 ^ v11;
    s.h[4] ^= v4 ^ v12;
    s.h[5] ^= v5 ^ v13;
    s.h[6] ^= v6 ^ v14;
    s.h[7] ^= v7 ^ v15;
}

// Host convenience: hash `len` bytes (little-endian word packing), out = 32 bytes.
inline void blake2s_hash_bytes(const uint8_t* data, size_t len, uint8_t out[32]) {
    Blake2sState s;
    blake2s_init(s);
    u32 m[16];
    size_t off = 0;
    while (len - off > 64) {
        memcpy(m, data + off, 64);
        off += 64;
        blake2s_compress(s, m, off, false);
    }
    uint8_t last[64];
    memset(last, 0, 64);
    memcpy(last, data + off, len - off);
    memcpy(m, last, 64);
    blake2s_compress(s, m, len, true);
    memcpy(out, s.h, 32);
}

}  // namespace cm31

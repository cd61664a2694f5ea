// Blake2sChannel + Blake2sMerkleChannel::mix_root (host, Fiat-Shamir transcript).
// Mirrors external/stwo/crates/prover/src/core/channel/blake2s.rs:15-116,
// channel/mod.rs:19-33, vcs/blake2_merkle.rs:36-46, queries.rs:21-50.
#pragma once
#include <algorithm>
#include <set>
#include <vector>

#include "../blake2s.cuh"
#include "qm_ops.hpp"

namespace cm31 {

class Blake2sChannel {
   public:
    Blake2sChannel() { memset(digest_.b, 0, 32); }
    const Hash32& digest() const { return digest_; }
    void update_digest(const Hash32& d) {
        digest_ = d;
        n_sent_ = 0;
        n_challenges_++;
    }
    void mix_u32s(const u32* data, size_t n) {
        std::vector<uint8_t> buf(32 + 4 * n);
        memcpy(buf.data(), digest_.b, 32);
        memcpy(buf.data() + 32, data, 4 * n);  // little-endian host
        Hash32 h;
        blake2s_hash_bytes(buf.data(), buf.size(), h.b);
        update_digest(h);
    }
    void mix_u64(u64 v) {
        u32 d[2] = {(u32)v, (u32)(v >> 32)};
        mix_u32s(d, 2);
    }
    void mix_felts(const std::vector<QM31>& felts) {
        std::vector<u32> w;
        w.reserve(felts.size() * 4);
        for (const QM31& f : felts) {
            w.push_back(f.a);
            w.push_back(f.b);
            w.push_back(f.c);
            w.push_back(f.d);
        }
        mix_u32s(w.data(), w.size());
    }
    void mix_root(const Hash32& root) {
        uint8_t buf[64];
        memcpy(buf, digest_.b, 32);
        memcpy(buf + 32, root.b, 32);
        Hash32 h;
        blake2s_hash_bytes(buf, 64, h.b);
        update_digest(h);
    }
    Hash32 draw_random_bytes() {
        uint8_t buf[36];
        memcpy(buf, digest_.b, 32);
        memcpy(buf + 32, &n_sent_, 4);
        n_sent_++;
        Hash32 h;
        blake2s_hash_bytes(buf, 36, h.b);
        return h;
    }
    QM31 draw_secure_felt() {
        u32 f[8];
        draw_base_felts(f);
        return qm_make(f[0], f[1], f[2], f[3]);
    }
    std::vector<QM31> draw_secure_felts(size_t n) {
        std::vector<QM31> out;
        while (out.size() < n) {
            u32 f[8];
            draw_base_felts(f);
            out.push_back(qm_make(f[0], f[1], f[2], f[3]));
            if (out.size() < n) out.push_back(qm_make(f[4], f[5], f[6], f[7]));
        }
        return out;
    }
    u32 trailing_zeros() const {
        for (int i = 0; i < 16; i++)
            if (digest_.b[i]) return 8 * i + __builtin_ctz(digest_.b[i]);
        return 128;
    }

   private:
    void draw_base_felts(u32 out[8]) {
        for (;;) {
            Hash32 h = draw_random_bytes();
            u32 w[8];
            memcpy(w, h.b, 32);
            bool ok = true;
            for (int i = 0; i < 8; i++) ok = ok && w[i] < 2 * P;
            if (ok) {
                for (int i = 0; i < 8; i++) out[i] = w[i] >= P ? w[i] - P : w[i];
                return;
            }
        }
    }
    Hash32 digest_;
    u32 n_sent_ = 0;
    size_t n_challenges_ = 0;
};

// CirclePoint::<SecureField>::get_random_point (circle.rs:169-181)
inline SecurePoint get_random_point(Blake2sChannel& channel) {
    QM31 t = channel.draw_secure_felt();
    QM31 t_square = qm_sqr(t);
    QM31 one_plus_tsquared_inv = qm_inv(t_square + qm_one());
    SecurePoint p;
    p.x = (qm_one() + (-t_square)) * one_plus_tsquared_inv;
    p.y = (t + t) * one_plus_tsquared_inv;
    return p;
}

// Queries::generate / fold (queries.rs:21-50)
struct Queries {
    std::vector<size_t> positions;
    u32 log_domain_size = 0;

    static Queries generate(Blake2sChannel& channel, u32 log_domain_size, size_t n_queries) {
        std::set<size_t> q;
        size_t cnt = 0;
        u32 max_query = (u32)(((u64)1 << log_domain_size) - 1);
        for (;;) {
            Hash32 h = channel.draw_random_bytes();
            for (int i = 0; i < 8; i++) {
                u32 bits;
                memcpy(&bits, h.b + 4 * i, 4);
                q.insert(bits & max_query);
                if (++cnt == n_queries) {
                    Queries out;
                    out.positions.assign(q.begin(), q.end());
                    out.log_domain_size = log_domain_size;
                    return out;
                }
            }
        }
    }
    Queries fold(u32 n_folds) const {
        Queries out;
        out.log_domain_size = log_domain_size - n_folds;
        for (size_t p : positions) {
            size_t f = p >> n_folds;
            if (out.positions.empty() || out.positions.back() != f) out.positions.push_back(f);
        }
        return out;
    }
};

}  // namespace cm31

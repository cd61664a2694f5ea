"""GPU parity: every Backend op of libcm31 (through the C ABI) against the CPU oracle, bit exact.

Mirrors the reference's "SIMD result == CPU result on seeded input" tests
(simd/fft/rfft.rs:725-754, ifft.rs:670-, simd/quotients.rs:345-394, simd/fri.rs:187-,
simd/blake2s.rs:462, simd/grind.rs:136-157) with the CUDA backend in SimdBackend's place.
"""
import numpy as np
import pytest

from tests import oracle_lib as orc

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

P = orc.P


def dev(a: np.ndarray):
    return torch.from_numpy(a.astype(np.int64).astype(np.int32) if a.dtype != np.int32 else a).cuda()


def to_dev_cols(mat: np.ndarray):
    """(n_cols, n) u32 -> list of contiguous int32 cuda tensors (one per column)."""
    return [torch.from_numpy(np.ascontiguousarray(mat[c]).view(np.int32)).cuda() for c in range(mat.shape[0])]


def host(t) -> np.ndarray:
    return t.cpu().numpy().view(np.uint32)


@pytest.fixture(scope="module")
def tw(cm):
    t = cm.Twiddles(27)
    yield t
    t.close()


def test_twiddle_tree_matches_oracle(cm):
    for L in [3, 5, 12]:
        t = cm.Twiddles(L)
        g_tw, g_itw = t.buffers()
        o_tw, o_itw = orc.twiddles(L)
        assert g_tw == o_tw.tolist()
        assert g_itw == o_itw.tolist()
        t.close()


@pytest.mark.parametrize("L", [1, 2, 3, 4, 5, 7, 8, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 25])
def test_interpolate_matches_oracle(cm, tw, L):
    n_cols = 3 if L <= 11 else (5 if L <= 20 else (2 if L <= 23 else 1))  # 5: one full column quad + a ragged one (fft4.cu)
    vals = orc.splitmix64(0xCA1120 + L, n_cols << L).reshape(n_cols, 1 << L)
    cols = to_dev_cols(vals)
    cm.interpolate_batch(cols, L, tw)
    cm.sync()
    expect = orc.interpolate(vals, L)
    for c in range(n_cols):
        assert np.array_equal(host(cols[c]), expect[c]), f"column {c}"


@pytest.mark.parametrize("L,LE", [(1, 1), (1, 2), (2, 2), (2, 3), (3, 3), (3, 4), (4, 5), (7, 8), (10, 11), (11, 12),
                                  (12, 13), (13, 14), (14, 15), (15, 16), (16, 17), (17, 18), (18, 19), (19, 20), (5, 9),
                                  (20, 21), (21, 22), (22, 23), (12, 12), (10, 14), (3, 13), (1, 12), (20, 22), (24, 25), (25, 26)])
def test_evaluate_matches_oracle(cm, tw, L, LE):
    n_cols = 3 if LE <= 11 else (5 if LE <= 20 else (2 if LE <= 23 else 1))
    coeffs = orc.splitmix64(0xBEEF + L, n_cols << L).reshape(n_cols, 1 << L)
    src = to_dev_cols(coeffs)
    out = [torch.empty(1 << LE, dtype=torch.int32, device="cuda") for _ in range(n_cols)]
    cm.evaluate_batch(src, out, L, LE, tw)
    cm.sync()
    expect = orc.evaluate(coeffs, L, LE)
    for c in range(n_cols):
        assert np.array_equal(host(out[c]), expect[c]), f"column {c}"


@pytest.mark.parametrize("L", [22, 24])
def test_fft_roundtrip_and_linearity_full_size(cm, tw, L):
    # size-independent properties at BASELINE sizes: interpolate∘evaluate = id, linearity,
    # and the LDE restricted to... (spot rows checked against eval_at_point)
    n = 1 << L
    a = orc.splitmix64(1, n)
    b = orc.splitmix64(2, n)
    s = ((a.astype(np.uint64) + b) % P).astype(np.uint32)
    cols = to_dev_cols(np.stack([a, b, s]))
    cm.interpolate_batch(cols, L, tw)
    ca, cb, cs = (host(c).astype(np.uint64) for c in cols)
    assert np.array_equal((ca + cb) % P, cs)
    out = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(3)]
    cm.evaluate_batch(cols, out, L, L, tw)
    cm.sync()
    assert np.array_equal(host(out[0]), a)
    assert np.array_equal(host(out[1]), b)
    # spot-check one LDE value against the definition (eval_at_point on the oracle)
    lde = [torch.empty(2 * n, dtype=torch.int32, device="cuda")]
    cm.evaluate_batch(cols[:1], lde, L, L + 1, tw)
    cm.sync()
    h = host(lde[0])
    coeffs = host(cols[0])
    for i in [0, 12345, 2 * n - 1]:
        d = int(format(i, f"0{L + 1}b")[::-1], 2)
        x, y = orc.domain_at(L + 1, d)
        got = cm.eval_at_point_batch(cols[:1], [L], [(x, 0, 0, 0, y, 0, 0, 0)], [0])[0]
        assert got == (int(h[i]), 0, 0, 0)


@pytest.mark.parametrize("L", [0, 1, 2, 3, 5, 10, 11, 12, 15, 20])
def test_eval_at_point_matches_oracle(cm, L):
    n_polys = 3
    coeffs = orc.splitmix64(0xE1 + L, n_polys << L).reshape(n_polys, 1 << L)
    cols = to_dev_cols(coeffs)
    pts = [tuple(int(v) for v in orc.splitmix64(50 + k, 8)) for k in range(2)]
    idx = [0, 1, 0]
    got = cm.eval_at_point_batch(cols, [L] * n_polys, pts, idx)
    for i in range(n_polys):
        assert got[i] == orc.eval_at_point(coeffs[i], L, pts[idx[i]])


def test_eval_at_point_mixed_sizes(cm):
    sizes = [4, 13, 0, 9, 16]
    mats = [orc.splitmix64(9 + i, 1 << L) for i, L in enumerate(sizes)]
    cols = [torch.from_numpy(m.view(np.int32)).cuda() for m in mats]
    pts = [tuple(int(v) for v in orc.splitmix64(77, 8))]
    got = cm.eval_at_point_batch(cols, sizes, pts, [0] * len(sizes))
    for i, L in enumerate(sizes):
        assert got[i] == orc.eval_at_point(mats[i], L, pts[0])


def test_bit_reverse(cm):
    L = 10
    a = orc.splitmix64(3, 1 << L)
    t = torch.from_numpy(a.view(np.int32)).cuda()
    cm.bit_reverse(t, L)
    cm.sync()
    idx = np.array([int(format(i, f"0{L}b")[::-1], 2) for i in range(1 << L)])
    assert np.array_equal(host(t), a[idx])


@pytest.mark.parametrize("L,n_cols,with_prev", [(0, 0, True), (0, 3, False), (3, 5, False), (3, 2, True), (6, 16, True),
                                                (6, 17, False), (9, 33, True), (12, 64, False), (10, 0, True),
                                                (14, 7, True)])
def test_commit_on_layer_matches_oracle(cm, L, n_cols, with_prev):
    n = 1 << L
    mat = orc.splitmix64(0xAB + L, max(1, n_cols) * n).reshape(max(1, n_cols), n)[:n_cols]
    cols = to_dev_cols(mat) if n_cols else []
    prev = orc.splitmix64(0xCD + L, 2 * n * 8).reshape(2 * n, 8) if with_prev else None
    # hashes are arbitrary u32 words, not field elements: keep full 32 bits
    if prev is not None:
        prev = (prev.astype(np.uint64) * 3 + 0x80000001).astype(np.uint32)
    dprev = torch.from_numpy(prev.view(np.int32)).cuda() if prev is not None else None
    out = torch.empty((n, 8), dtype=torch.int32, device="cuda")
    cm.blake2s_commit_layer(L, dprev, cols, out)
    cm.sync()
    expect = orc.commit_on_layer(L, prev, mat if n_cols else None)
    assert np.array_equal(host(out), expect)


def test_merkle_tree_root_full_size(cm):
    # 2^20 leaves x 8 columns: root equals the oracle's root (the oracle finishes this in seconds)
    L, n_cols = 16, 8
    mat = orc.splitmix64(4242, n_cols << L).reshape(n_cols, 1 << L)
    cols = to_dev_cols(mat)
    prev_d, prev_o = None, None
    for log in range(L, -1, -1):
        out = torch.empty((1 << log, 8), dtype=torch.int32, device="cuda")
        cm.blake2s_commit_layer(log, prev_d, cols if log == L else [], out)
        prev_o = orc.commit_on_layer(log, prev_o, mat if log == L else None)
        prev_d = out
    cm.sync()
    assert np.array_equal(host(prev_d), prev_o)


@pytest.mark.parametrize("L", [1, 2, 3, 4, 8, 13, 18])
def test_fold_line_matches_oracle(cm, tw, L):
    src = orc.splitmix64(0xF0 + L, 4 << L).reshape(4, 1 << L)
    alpha = (1, 3, 5, 7)  # simd/fri.rs:191
    s = to_dev_cols(src)
    d = [torch.empty(1 << (L - 1), dtype=torch.int32, device="cuda") for _ in range(4)]
    cm.fold_line(s, L, alpha, tw, d)
    cm.sync()
    expect = orc.fold_line(src, L, alpha)
    for k in range(4):
        assert np.array_equal(host(d[k]), expect[k])


@pytest.mark.parametrize("L", [1, 2, 3, 4, 5, 9, 14, 18])
def test_fold_circle_into_line_matches_oracle(cm, tw, L):
    src = orc.splitmix64(0xF1 + L, 4 << L).reshape(4, 1 << L)
    dst = orc.splitmix64(0xF2 + L, 4 << (L - 1)).reshape(4, 1 << (L - 1))
    alpha = tuple(int(v) for v in orc.splitmix64(0xF3, 4))
    s = to_dev_cols(src)
    d = to_dev_cols(dst)
    cm.fold_circle_into_line(d, s, L, alpha, tw)
    cm.sync()
    expect = orc.fold_circle_into_line(dst, src, L, alpha)
    for k in range(4):
        assert np.array_equal(host(d[k]), expect[k])


def test_fri_fold_chain_low_degree_full_size(cm, tw):
    # K2: LDE of a degree < 2^(k-1) secure poly folds down to a constant pair
    k = 21
    coeffs = orc.splitmix64(0x77, 4 << (k - 1)).reshape(4, 1 << (k - 1))
    c = to_dev_cols(coeffs)
    ev = [torch.empty(1 << k, dtype=torch.int32, device="cuda") for _ in range(4)]
    cm.evaluate_batch(c, ev, k - 1, k, tw)
    alpha = (1, 3, 5, 7)
    cur = [torch.zeros(1 << (k - 1), dtype=torch.int32, device="cuda") for _ in range(4)]
    cm.fold_circle_into_line(cur, ev, k, alpha, tw)
    log = k - 1
    while log > 1:
        nxt = [torch.empty(1 << (log - 1), dtype=torch.int32, device="cuda") for _ in range(4)]
        cm.fold_line(cur, log, alpha, tw, nxt)
        cur, log = nxt, log - 1
    cm.sync()
    last = np.stack([host(t) for t in cur])
    assert np.array_equal(last[:, 0], last[:, 1])


@pytest.mark.parametrize("L", [1, 4, 11])
def test_decompose_matches_oracle(cm, L):
    src = orc.splitmix64(0xD0 + L, 4 << L).reshape(4, 1 << L)
    s = to_dev_cols(src)
    d = [torch.empty(1 << L, dtype=torch.int32, device="cuda") for _ in range(4)]
    lam = cm.decompose(s, L, d)
    cm.sync()
    expect, elam = orc.decompose(src, L)
    assert lam == elam
    for k in range(4):
        assert np.array_equal(host(d[k]), expect[k])


def test_accumulate_and_powers(cm):
    n = 1000
    a = orc.splitmix64(1, 4 * n).reshape(4, n)
    b = orc.splitmix64(2, 4 * n).reshape(4, n)
    da, db = to_dev_cols(a), to_dev_cols(b)
    cm.accumulate(da, db, n)
    cm.sync()
    for k in range(4):
        assert np.array_equal(host(da[k]), ((a[k].astype(np.uint64) + b[k]) % P).astype(np.uint32))
    alpha = (5, 6, 7, 8)
    pw = cm.secure_powers(alpha, 6)
    acc = (1, 0, 0, 0)
    for i in range(6):
        assert pw[i] == acc
        acc = orc.qm31_mul(acc, alpha)


@pytest.mark.parametrize("L,n_cols", [(1, 1), (3, 2), (8, 5), (12, 9), (16, 3)])
def test_accumulate_quotients_matches_oracle(cm, L, n_cols):
    mat = orc.splitmix64(0x51 + L, n_cols << L).reshape(n_cols, 1 << L)
    cols = to_dev_cols(mat)
    rc = tuple(int(v) for v in orc.splitmix64(0x52, 4))
    p0 = tuple(int(v) for v in orc.splitmix64(0x53, 8))
    p1 = tuple(int(v) for v in orc.splitmix64(0x54, 8))
    vals = [tuple(int(v) for v in orc.splitmix64(0x60 + i, 4)) for i in range(2 * n_cols)]
    batches = [(p0, [(c, vals[c]) for c in range(n_cols)]), (p1, [(c, vals[n_cols + c]) for c in range(0, n_cols, 2)])]
    out = [torch.empty(1 << L, dtype=torch.int32, device="cuda") for _ in range(4)]
    cm.accumulate_quotients(L, cols, rc, batches, out)
    cm.sync()
    expect = orc.accumulate_quotients(L, mat, rc, batches)
    for k in range(4):
        assert np.array_equal(host(out[k]), expect[k])


@pytest.mark.parametrize("bits", [0, 1, 5, 12, 16, 20])
def test_grind_matches_oracle(cm, bits):
    digest = bytes((7 * i + bits) & 0xFF for i in range(32))
    assert cm.grind_blake2s(digest, bits) == orc.grind(digest, bits)


def test_gather(cm):
    mat = orc.splitmix64(5, 3 * 64).reshape(3, 64)
    cols = to_dev_cols(mat)
    idx = [0, 5, 63, 17]
    got = cm.gather_u32(cols, idx)
    for c in range(3):
        assert got[c] == mat[c, idx].tolist()


def test_gather_words(cm):
    mat = orc.splitmix64(6, 4 * 256).reshape(4, 256)
    cols = to_dev_cols(mat)
    rng = np.random.default_rng(3)
    sid = rng.integers(0, 4, 1000).tolist()
    widx = rng.integers(0, 256, 1000).tolist()
    got = cm.gather_words(cols, sid, widx)
    assert got == [int(mat[s, w]) for s, w in zip(sid, widx)]


@pytest.mark.parametrize("top,with_prev,col_layers", [(0, False, {0: 2}), (3, True, {}), (5, False, {5: 3, 4: 1, 2: 20}),
                                                      (10, True, {10: 4, 7: 17, 0: 1}), (10, False, {10: 33}),
                                                      # layers with many injected columns (1100 x 16 like cairo-m's padding
                                                      # components): long per-node hash chains
                                                      (4, True, {4: 1100, 3: 40, 0: 33}), (6, False, {6: 64, 5: 32, 4: 31, 1: 257}),
                                                      (7, True, {7: 500, 2: 1000})])
def test_commit_top_layers_matches_layer_by_layer_oracle(cm, top, with_prev, col_layers):
    # the fused top-of-tree launch == MerkleProver::commit's per-layer loop (vcs/prover.rs:52-64)
    mats = {l: orc.splitmix64(0x70 + l, k << l).reshape(k, 1 << l) for l, k in col_layers.items()}
    prev = None
    if with_prev:
        prev = (orc.splitmix64(0x99, (2 << top) * 8).astype(np.uint64) * 5 + 0x80000003).astype(np.uint32).reshape(2 << top, 8)
    dprev = torch.from_numpy(prev.view(np.int32)).cuda() if prev is not None else None
    dcols = {l: to_dev_cols(m) for l, m in mats.items()}
    outs = [torch.empty((1 << l, 8), dtype=torch.int32, device="cuda") for l in range(top + 1)]
    cm.blake2s_commit_top(top, dprev, [dcols.get(l, []) for l in range(top + 1)], outs)
    cm.sync()
    expect_prev = prev
    for l in range(top, -1, -1):
        expect_prev = orc.commit_on_layer(l, expect_prev, mats.get(l))
        assert np.array_equal(host(outs[l]), expect_prev), f"layer {l}"


@pytest.mark.parametrize("log_size,n_cols,with_prev,n_levels", [(8, 4, False, 9), (11, 4, False, 1), (13, 0, True, 3), (12, 21, True, 2),
                                                               (16, 4, False, 6), (17, 17, True, 9),
                                                               # >= 2^19 nodes: barrier-free warp subtrees (4 nodes per lane: 3
                                                               # levels per launch; 8 per lane from 2^21: 4 levels), then the rest
                                                               (19, 4, False, 2), (19, 0, True, 3), (20, 5, True, 9), (21, 4, False, 4),
                                                               (21, 18, True, 6), (22, 4, False, 9)])
def test_commit_multi_matches_layer_by_layer_oracle(cm, log_size, n_cols, with_prev, n_levels):
    # fused consecutive layers (first with columns / a previous layer, the rest column-free) == one commit_on_layer per layer
    mat = orc.splitmix64(0x51 + log_size, n_cols << log_size).reshape(n_cols, 1 << log_size) if n_cols else None
    prev = None
    if with_prev:
        prev = (orc.splitmix64(0x77, (2 << log_size) * 8).astype(np.uint64) * 7 + 0x80000005).astype(np.uint32).reshape(2 << log_size, 8)
    dprev = torch.from_numpy(prev.view(np.int32)).cuda() if prev is not None else None
    dcols = to_dev_cols(mat) if n_cols else []
    outs = [torch.empty((1 << (log_size - l), 8), dtype=torch.int32, device="cuda") for l in range(n_levels)]
    cm.blake2s_commit_multi(log_size, dprev, dcols, outs)
    cm.sync()
    expect = prev
    for l in range(n_levels):
        expect = orc.commit_on_layer(log_size - l, expect, mat if l == 0 else None)
        assert np.array_equal(host(outs[l]), expect), f"level {l}"


def test_gather_runs(cm):
    mat = orc.splitmix64(8, 3 * 512).reshape(3, 512)
    cols = to_dev_cols(mat)
    rng = np.random.default_rng(4)
    sid = rng.integers(0, 3, 300).tolist()
    counts = rng.choice([1, 8], 300).tolist()
    widx = [int(rng.integers(0, 512 - c + 1)) for c in counts]
    got = cm.gather_runs(cols, sid, widx, counts)
    want = [int(v) for s, w, c in zip(sid, widx, counts) for v in mat[s, w:w + c]]
    assert got == want


def test_gather_batch_runs_and_row_grids(cm):
    # the per-proof decommitment gather: interleaved run requests (hash nodes, single words) and row grids (all columns of a
    # Merkle layer at the visited nodes), including an empty-run-only and a grid-only call
    mat = orc.splitmix64(9, 5 * 1024).reshape(5, 1024)
    cols = to_dev_cols(mat)
    rng = np.random.default_rng(5)
    want, sid, widx, off, cnt, grids = [], [], [], [], [], []
    for it in range(40):
        if it % 3 == 2:
            cids = rng.permutation(5)[: int(rng.integers(1, 6))].tolist()
            rows = sorted(rng.choice(1024, int(rng.integers(1, 50)), replace=False).tolist())
            grids.append((cids, rows, len(want)))
            want += [int(mat[c, r]) for r in rows for c in cids]
        else:
            c = int(rng.choice([1, 8]))
            s, w = int(rng.integers(0, 5)), int(rng.integers(0, 1024 - c + 1))
            sid.append(s); widx.append(w); off.append(len(want)); cnt.append(c)
            want += [int(v) for v in mat[s, w:w + c]]
    assert cm.gather_batch(cols, sid, widx, off, cnt, grids, len(want)) == want
    g_only = [([0, 4], [3, 1000], 0)]
    assert cm.gather_batch(cols, [], [], [], [], g_only, 4) == [int(mat[0, 3]), int(mat[4, 3]), int(mat[0, 1000]), int(mat[4, 1000])]
    assert cm.gather_batch(cols, [2], [7], [0], [8], [], 8) == [int(v) for v in mat[2, 7:15]]


def test_sharded_commit_world1_equals_layer_by_layer_oracle(cm):
    # cairo-m_b200/sharded_commit.py on one GPU == interpolate -> LDE -> Merkle layers of the oracle
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("sharded_commit", Path(__file__).resolve().parent.parent / "cairo-m_b200" / "sharded_commit.py")
    sc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sc)
    L, n_cols = 10, 7
    trace = orc.splitmix64(0x5EED, n_cols << L).reshape(n_cols, 1 << L)
    root, rows = sc.sharded_commit(sc.CudaOps(L + 2), to_dev_cols(trace), n_cols, L, 1)
    lde = orc.evaluate(orc.interpolate(trace, L), L, L + 1)
    layer = orc.commit_on_layer(L + 1, None, lde)
    for log in range(L, -1, -1):
        layer = orc.commit_on_layer(log, layer, None)
    assert np.array_equal(host(root), layer.reshape(8))
    assert all(np.array_equal(host(rows[c]), lde[c]) for c in range(n_cols))


def test_sharded_commit_p2p_world1_matches_nccl_variant(cm):
    # the peer-memory (cudaIpc) layout of the fused exchange: same root as the all-to-all variant
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("sharded_commit", Path(__file__).resolve().parent.parent / "cairo-m_b200" / "sharded_commit.py")
    sc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sc)
    L, n_cols = 12, 6
    trace = orc.splitmix64(0xABCD, n_cols << L).reshape(n_cols, 1 << L)
    ops = sc.CudaOps(L + 2)
    root_a, _ = sc.sharded_commit(ops, to_dev_cols(trace), n_cols, L, 1)
    peer = sc.PeerLde(n_cols, L + 1, None, 0, 1)
    try:
        root_b = sc.sharded_commit_p2p(ops, peer, to_dev_cols(trace), L, 1)
        assert np.array_equal(host(root_a), host(root_b))
    finally:
        peer.close()


@pytest.mark.parametrize("L", [1, 4, 10, 11, 12, 13, 16, 19, 22])
def test_logup_finalize_last_matches_oracle(cm, L):
    # LogupTraceGenerator::finalize_last (constraint_framework/src/logup.rs:211-251): claimed sum of the last cumulative
    # column, shift by claimed_sum / n, inclusive prefix sum in coset order (simd/prefix_sum.rs:19; the oracle follows
    # inclusive_prefix_sum_slow).  L <= 11: single-CTA kernel; L >= 12: the segmented coset-order scan.
    import ctypes as C
    n = 1 << L
    src = orc.splitmix64(0x10C0 + L, 4 * n).reshape(4, n)
    cols = to_dev_cols(src)
    ptrs = (C.c_void_p * 4)(*[t.data_ptr() for t in cols])
    claimed = (C.c_uint32 * 4)()
    cm.check(cm.lib().cm31_logup_finalize_last(ptrs, C.c_uint32(L), claimed))
    n_inv = pow(n % P, P - 2, P)
    for k in range(4):
        total = int(src[k].astype(np.uint64).sum() % P)
        assert claimed[k] == total
        shift = total * n_inv % P
        shifted = ((src[k].astype(np.int64) - shift) % P).astype(np.uint32)
        assert np.array_equal(host(cols[k]), orc.prefix_sum(shifted, L)), f"coordinate {k}"

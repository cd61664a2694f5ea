"""An INDEPENDENT replay of the Fiat-Shamir transcript of a proof, written from the reference's Rust source only -- no C++ of this
repo is involved beyond producing the proof and printing it in the reference's serde JSON layout.

The oracle prover and the CUDA prover share one protocol driver, so their byte-identity says nothing about
the ORDER and ENCODING of what is mixed into the channel.  This test re-walks the verifier's side of the transcript exactly as

    crates/prover/src/verifier.rs:17-95                     verify_cairo_m
    external/stwo/.../core/pcs/mod.rs:42-49, fri.rs:80-89   PcsConfig / FriConfig ::mix_into
    crates/prover/src/public_data.rs:132-189, 401-412       PublicMemory / PublicData ::mix_into
    crates/prover/src/components/mod.rs:94-104, 298-308     Claim / InteractionClaim ::mix_into (opcodes in define_opcodes! order)
    crates/prover/src/components/mod.rs:311-323 + constraint_framework/src/logup.rs:82-95   Relations::draw
    external/stwo/.../core/prover/mod.rs:87-140             verify: random coefficient, composition root, OODS point
    external/stwo/.../core/pcs/verifier.rs:55-84            verify_values: sampled values, random coefficient, FRI commit, PoW
    external/stwo/.../core/fri.rs:370-432                   FriVerifier::commit: layer roots / folding alphas / last layer
    external/stwo/.../core/channel/blake2s.rs:15-116        Blake2sChannel;  vcs/blake2_merkle.rs:36-46  mix_root

prescribe it (Python + hashlib), and checks the two places where the transcript state is observable in the proof: the
interaction proof of work (relations.rs:47, 2 bits) after the execution-trace commitment, and the FRI proof of work (16 bits)
after the last FRI layer.  A prover that mixed anything in another order, width or endianness than the reference passes the
second check with probability 2^-16.  From that channel state the FRI query positions are drawn (queries.rs:21-50,
fri.rs:560-603) and the Merkle decommitments of all four commitment trees are verified against the committed roots by a Python
restatement of MerkleVerifier::verify (vcs/verifier.rs:53-171, hasher vcs/blake2_merkle.rs:14-30), with the per-tree column counts
taken from the reference SOURCE (N_TRACE_COLUMNS / N_*_LOOKUPS constants, tests/golden/air_shapes_reference.json): this pins the
query derivation, the leaf / node hashing, the column order by size inside a tree and the number of columns every component
commits, independently of the oracle.  Between the two, the closing sum of the lookup argument (verifier.rs:64-72:
InteractionClaim::claimed_sum with PublicData::initial_logup_sum, public_data.rs:287-399) is recomputed with a Python QM31 and
the relation elements in Relations::draw order: it is zero only if the relations are drawn in the reference's order and
`combine`, the register / Merkle-root / public-memory terms and every component's claimed sum are the reference's."""
import ctypes as C
import hashlib
import json
from pathlib import Path

import pytest

from tests import cairo_helpers as ch
from tests.test_proof_json import to_json

P = (1 << 31) - 1
INTERACTION_POW_BITS = 2  # crates/prover/src/relations.rs:47
N_RELATIONS = 8           # components/mod.rs:311-323: registers, memory, merkle, poseidon2, range_check_8/16/20, bitwise


class Blake2sChannel:  # channel/blake2s.rs:15-116
    def __init__(self):
        self.digest = bytes(32)
        self.n_sent = 0

    def _update(self, data: bytes):
        self.digest = hashlib.blake2s(self.digest + data, digest_size=32).digest()
        self.n_sent = 0  # ChannelTime::inc_challenges

    def mix_u32s(self, words):
        self._update(b"".join(int(w).to_bytes(4, "little") for w in words))

    def mix_u64(self, v):
        self.mix_u32s([v & 0xFFFFFFFF, v >> 32])

    def mix_felts(self, felts):  # QM31 as [[a, b], [c, d]] (serde of SecureField)
        self.mix_u32s([x for f in felts for half in f for x in half])

    def mix_root(self, root_bytes):  # Blake2sMerkleChannel::mix_root = concat_and_hash(digest, root)
        self._update(bytes(root_bytes))

    def draw_random_bytes(self):
        out = hashlib.blake2s(self.digest + self.n_sent.to_bytes(4, "little"), digest_size=32).digest()
        self.n_sent += 1
        return out

    def draw_base_felts(self):
        while True:
            b = self.draw_random_bytes()
            words = [int.from_bytes(b[4 * i:4 * i + 4], "little") for i in range(8)]
            if all(w < 2 * P for w in words):
                return [w % P for w in words]

    def draw_secure_felt(self):
        return self.draw_base_felts()[:4]

    def draw_secure_felts(self, n):
        out, pool = [], []
        while len(out) < n:
            if len(pool) < 4:
                pool += self.draw_base_felts()
            out.append(pool[:4])
            pool = pool[4:]
        return out

    def trailing_zeros(self):
        v = int.from_bytes(self.digest[:16], "little")
        return 128 if v == 0 else (v & -v).bit_length() - 1


def replay(proof):
    """verify_cairo_m's transcript; returns (trailing zeros after interaction_pow, trailing zeros after proof_of_work, channel)."""
    sp = proof["stark_proof"]
    ch_ = Blake2sChannel()
    cfg = sp["config"]
    ch_.mix_u64(cfg["pow_bits"])                                        # PcsConfig::mix_into
    ch_.mix_u64(cfg["fri_config"]["log_blowup_factor"])                 # FriConfig::mix_into: blowup, n_queries, last layer bound
    ch_.mix_u64(cfg["fri_config"]["n_queries"])
    ch_.mix_u64(cfg["fri_config"]["log_last_layer_degree_bound"])
    pd = proof["public_data"]                                           # PublicData::mix_into
    ch_.mix_u32s([pd["initial_registers"]["pc"], pd["initial_registers"]["fp"], pd["final_registers"]["pc"], pd["final_registers"]["fp"],
                  pd["clock"], pd["initial_root"], pd["final_root"]])
    pm = pd["public_memory"]
    ch_.mix_u32s([len(pm["program"]), len(pm["input"]), len(pm["output"])])
    for part in ("program", "input", "output"):                        # iter().flatten(): None entries are skipped
        ch_.mix_u32s([w for e in pm[part] if e is not None for w in (e[0], e[1][0][0], e[1][0][1], e[1][1][0], e[1][1][1], e[2])])
    ch_.mix_root(sp["commitments"][0])                                  # preprocessed trace
    claim = proof["claim"]                                              # Claim::mix_into
    for comp in claim["opcodes"].values():
        ch_.mix_u64(comp["log_size"])
    for name in ("memory", "merkle", "clock_update", "poseidon2", "range_check_8", "range_check_16", "range_check_20", "bitwise"):
        ch_.mix_u64(claim[name]["log_size"])
    ch_.mix_root(sp["commitments"][1])                                  # execution traces
    ch_.mix_u64(proof["interaction_pow"])
    tz_interaction = ch_.trailing_zeros()
    rels = []
    for _ in range(N_RELATIONS):                                        # Relations::draw: [z, alpha] = draw_secure_felts(2) each
        z, alpha = ch_.draw_secure_felts(2)
        rels.append((q_from(z), q_from(alpha)))
    ch_.relations = rels
    ic = proof["interaction_claim"]                                     # InteractionClaim::mix_into
    for comp in ic["opcodes"].values():
        ch_.mix_felts([comp["claimed_sum"]])
    for name in ("memory", "merkle", "clock_update", "poseidon2", "range_check_8", "range_check_16", "range_check_20", "bitwise"):
        ch_.mix_felts([ic[name]["claimed_sum"]])
    ch_.mix_root(sp["commitments"][2])                                  # interaction traces
    ch_.draw_secure_felt()                                              # verify(): random_coeff
    ch_.mix_root(sp["commitments"][3])                                  # composition polynomial
    ch_.draw_secure_felt()                                              # get_random_point: t
    ch_.mix_felts([v for tree in sp["sampled_values"] for col in tree for v in col])  # verify_values: flatten_cols
    ch_.draw_secure_felt()                                              # random_coeff of the quotients
    fri = sp["fri_proof"]                                               # FriVerifier::commit
    ch_.mix_root(fri["first_layer"]["commitment"])
    ch_.draw_secure_felt()
    for layer in fri["inner_layers"]:
        ch_.mix_root(layer["commitment"])
        ch_.draw_secure_felt()
    ch_.mix_felts(fri["last_layer_poly"]["coeffs"])
    ch_.mix_u64(sp["proof_of_work"])
    return tz_interaction, ch_.trailing_zeros(), ch_


# ---------------------------------------------------------------- the lookup argument's closing sum (verifier.rs:64-72)
# QM31 = CM31[u] / (u^2 - 2 - i), CM31 = M31[i] / (i^2 + 1)  (core/fields/cm31.rs, qm31.rs:14-129); values as [[a, b], [c, d]]
def m_inv(x):
    return pow(x, P - 2, P)


def c_mul(x, y):
    return [(x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P]


def c_add(x, y):
    return [(x[0] + y[0]) % P, (x[1] + y[1]) % P]


def c_sub(x, y):
    return [(x[0] - y[0]) % P, (x[1] - y[1]) % P]


def c_inv(x):
    n = m_inv((x[0] * x[0] + x[1] * x[1]) % P)
    return [x[0] * n % P, (-x[1]) * n % P]


R = [2, 1]  # qm31.rs:14


def q_mul(x, y):  # (a + bu)(c + du) = (ac + R bd) + (ad + bc) u   (qm31.rs:78-88)
    return [c_add(c_mul(x[0], y[0]), c_mul(R, c_mul(x[1], y[1]))), c_add(c_mul(x[0], y[1]), c_mul(x[1], y[0]))]


def q_add(x, y):
    return [c_add(x[0], y[0]), c_add(x[1], y[1])]


def q_neg(x):
    return [[(-x[0][0]) % P, (-x[0][1]) % P], [(-x[1][0]) % P, (-x[1][1]) % P]]


def q_inv(x):  # (a + bu)^-1 = (a - bu) / (a^2 - R b^2)   (qm31.rs:119-129)
    d = c_inv(c_sub(c_mul(x[0], x[0]), c_mul(R, c_mul(x[1], x[1]))))
    return [c_mul(x[0], d), c_mul([(-x[1][0]) % P, (-x[1][1]) % P], d)]


def q_from(felts4):
    return [[felts4[0], felts4[1]], [felts4[2], felts4[3]]]


def combine(rel, values):  # LookupElements::combine (constraint_framework/src/logup.rs:96-111): sum alpha^i v_i - z
    z, alpha = rel
    acc, power = [[0, 0], [0, 0]], [[1, 0], [0, 0]]
    for v in values:
        acc = q_add(acc, q_mul(power, [[v % P, 0], [0, 0]]))
        power = q_mul(power, alpha)
    return q_add(acc, q_neg(z))


TREE_HEIGHT = 30  # adapter/merkle.rs:62: MAX_MEMORY_LOG_SIZE (28) + QM31_LOG_SIZE (2)


def logup_total(proof, rels):
    """InteractionClaim::claimed_sum (components/mod.rs:283-296) with PublicData::initial_logup_sum (public_data.rs:287-399);
    rels = the 8 [z, alpha] pairs in Relations::draw order: registers, memory, merkle, poseidon2, rc8, rc16, rc20, bitwise."""
    registers, memory, merkle = rels[0], rels[1], rels[2]
    pd = proof["public_data"]
    one = [[1, 0], [0, 0]]
    terms = [
        combine(registers, [pd["initial_registers"]["pc"], pd["initial_registers"]["fp"], 1]),
        q_neg(combine(registers, [pd["final_registers"]["pc"], pd["final_registers"]["fp"], pd["clock"] + 1])),
        combine(merkle, [0, 0, pd["initial_root"], pd["initial_root"]]),
        combine(merkle, [0, 0, pd["final_root"], pd["final_root"]]),
    ]
    for part, mult, root in (("program", one, pd["initial_root"]), ("input", one, pd["initial_root"]), ("output", q_neg(one), pd["final_root"])):
        for e in pd["public_memory"][part]:
            if e is None:
                continue
            addr, value, clock = e[0], [value for half in e[1] for value in half], e[2]
            terms.append(q_mul(mult, combine(memory, [addr, clock] + value)))
            for k in range(4):
                terms.append(q_neg(combine(merkle, [4 * addr + k, TREE_HEIGHT, value[k], root])))
    total = [[0, 0], [0, 0]]
    for t in terms:
        total = q_add(total, q_inv(t))
    ic = proof["interaction_claim"]
    for comp in ic["opcodes"].values():
        total = q_add(total, comp["claimed_sum"])
    for name in ("memory", "merkle", "clock_update", "poseidon2", "range_check_8", "range_check_16", "range_check_20", "bitwise"):
        total = q_add(total, ic[name]["claimed_sum"])
    return total


# ---------------------------------------------------------------- Merkle decommitments (vcs/verifier.rs:53-171)
def hash_node(children, values):  # Blake2sMerkleHasher::hash_node (vcs/blake2_merkle.rs:14-30)
    h = hashlib.blake2s(digest_size=32)
    if children is not None:
        h.update(children[0])
        h.update(children[1])
    for v in values:
        h.update(int(v).to_bytes(4, "little"))
    return h.digest()


def merkle_verify(root, n_columns_per_log_size, queries_per_log_size, queried_values, decommitment):
    """MerkleVerifier::verify, statement for statement; returns None or the name of the reference's error."""
    qv, hw, cw = iter(queried_values), iter(decommitment["hash_witness"]), iter(decommitment["column_witness"])
    last = None
    for layer_log_size in range(max(n_columns_per_log_size), -1, -1):
        n_cols = n_columns_per_log_size.get(layer_log_size, 0)
        prev_q = [q for q, _ in last] if last is not None else []
        prev_h, ph = last, 0
        col_q = list(queries_per_log_size.get(layer_log_size, []))
        pi = ci = 0
        total = []
        while pi < len(prev_q) or ci < len(col_q):
            node = min(([prev_q[pi] // 2] if pi < len(prev_q) else []) + ([col_q[ci]] if ci < len(col_q) else []))  # next_decommitment_node
            while pi < len(prev_q) and prev_q[pi] // 2 == node:
                pi += 1
            children = None
            if prev_h is not None:
                pair = []
                for child in (2 * node, 2 * node + 1):
                    if ph < len(prev_h) and prev_h[ph][0] == child:
                        pair.append(prev_h[ph][1])
                        ph += 1
                    else:
                        w = next(hw, None)
                        if w is None:
                            return "WitnessTooShort"
                        pair.append(bytes(w))
                children = tuple(pair)
            if ci < len(col_q) and col_q[ci] == node:
                ci += 1
                src, err = qv, "TooFewQueriedValues"
            else:
                src, err = cw, "WitnessTooShort"
            values = [v for v in (next(src, None) for _ in range(n_cols)) if v is not None]
            if len(values) != n_cols:
                return err
            total.append((node, hash_node(children, values)))
        last = total
    if next(hw, None) is not None or next(cw, None) is not None:
        return "WitnessTooLong"
    if next(qv, None) is not None:
        return "TooManyQueriedValues"
    if len(last) != 1 or last[0][1] != bytes(root):
        return "RootMismatch"
    return None


def tree_column_counts(proof, shapes):
    """n_columns_per_log_size of the 4 commitment trees (extended sizes: log_size + log_blowup_factor), from the claim and the
    column counts the reference SOURCE declares (tests/golden/air_shapes_reference.json: N_TRACE_COLUMNS, N_*_LOOKUPS;
    interaction columns = SECURE_EXTENSION_DEGREE * ceil(lookups / 2), e.g. opcodes/store_fp_imm.rs:96-101, memory.rs:56-59)."""
    blowup = proof["stark_proof"]["config"]["fri_config"]["log_blowup_factor"]
    claim = proof["claim"]
    log_size = dict((name, c["log_size"]) for name, c in claim["opcodes"].items())
    log_size.update((k, v["log_size"]) for k, v in claim.items() if k != "opcodes")
    trees = [{}, {}, {}, {}]

    def add(tree, ls, n):
        trees[tree][ls + blowup] = trees[tree].get(ls + blowup, 0) + n
    # PreProcessedTraceBuilder::default() (preprocessed/mod.rs:75-83): bitwise(8) = 4 columns of 2*8+2 bits, range checks 8/16/20
    add(0, 18, 4)
    for bits in (8, 16, 20):
        add(0, bits, 1)
    for comp in shapes["components"]:
        ls = log_size[comp["name"]]
        add(1, ls, comp["n_trace_columns"])
        add(2, ls, 4 * ((sum(comp["lookups"].values()) + 1) // 2))
    add(3, max(log_size.values()) + 1, 4)  # composition_log_degree_bound = max(log_size + 1), 4 coordinate columns
    return trees


def check_decommitments(proof, channel, shapes):
    sp = proof["stark_proof"]
    trees = tree_column_counts(proof, shapes)
    column_log_sizes = sorted({ls for t in trees for ls in t})
    max_log = column_log_sizes[-1]
    # Queries::generate (queries.rs:21-40) + get_query_positions_by_log_size (fri.rs:592-603)
    n_queries = sp["config"]["fri_config"]["n_queries"]
    positions, cnt = set(), 0
    while cnt < n_queries:
        b = channel.draw_random_bytes()
        for i in range(8):
            positions.add(int.from_bytes(b[4 * i:4 * i + 4], "little") & ((1 << max_log) - 1))
            cnt += 1
            if cnt == n_queries:
                break
    positions = sorted(positions)
    per_log = {}
    for ls in column_log_sizes:
        folded = []
        for q in positions:
            f = q >> (max_log - ls)
            if not folded or folded[-1] != f:
                folded.append(f)
        per_log[ls] = folded
    for t in range(4):
        err = merkle_verify(sp["commitments"][t], trees[t], per_log, sp["queried_values"][t], sp["decommitments"][t])
        assert err is None, f"tree {t}: {err}"
    return trees, per_log


def check(cm, blob):
    proof = json.loads(to_json(cm, blob))
    tz_i, tz_pow, channel = replay(proof)
    assert tz_i >= INTERACTION_POW_BITS, "interaction proof of work: the transcript up to the execution-trace commitment differs"
    assert tz_pow >= proof["stark_proof"]["config"]["pow_bits"], "proof of work: the transcript up to the last FRI layer differs"
    # InvalidLogupSum (verifier.rs:64-72): the claimed sums plus the public data's terms cancel under the relations drawn above
    assert logup_total(proof, channel.relations) == [[0, 0], [0, 0]], "the lookup argument does not close"
    # the channel now stands where the reference verifier draws the FRI queries: the 4 trees' decommitments must verify at
    # exactly those positions, against the committed roots, with the column counts the reference source declares
    shapes = json.loads((Path(__file__).resolve().parent / "golden" / "air_shapes_reference.json").read_text())
    check_decommitments(proof, channel, shapes)
    return tz_i, tz_pow


@pytest.mark.parametrize("program,n", [(ch.FIB, 10), (ch.U32_MIX, 3), (ch.ARRAY_SUM, 7)])
def test_oracle_proof_transcript_replays_as_the_reference_verifier_prescribes(cm, program, n):
    blob, _ = ch.oracle_program_prove(program, n)
    tz_i, tz_pow = check(cm, blob)
    assert tz_pow >= 16


def test_a_reordered_transcript_fails_the_replay(cm):
    # sensitivity: swap two roots in the proof (what a prover mixing them in the wrong order would produce): the FRI proof
    # of work no longer verifies
    blob, _ = ch.oracle_program_prove(ch.FIB, 10)
    proof = json.loads(to_json(cm, blob))
    c = proof["stark_proof"]["commitments"]
    c[2], c[3] = c[3], c[2]
    assert replay(proof)[1] < 16
    proof = json.loads(to_json(cm, blob))
    proof["claim"]["memory"]["log_size"] += 1
    tz_i, tz_pow, _ = replay(proof)
    assert tz_pow < 16
    # the closing sum is zero only under the reference's relation order and public-data terms
    proof = json.loads(to_json(cm, blob))
    _, _, channel = replay(proof)
    rels = list(channel.relations)
    assert logup_total(proof, rels) == [[0, 0], [0, 0]]
    rels[1], rels[2] = rels[2], rels[1]  # memory <-> merkle
    assert logup_total(proof, rels) != [[0, 0], [0, 0]]
    proof["public_data"]["clock"] += 1
    assert logup_total(proof, channel.relations) != [[0, 0], [0, 0]]
    # ... and the decommitment check is not vacuous: one flipped queried value, or one column too many, is caught
    shapes = json.loads((Path(__file__).resolve().parent / "golden" / "air_shapes_reference.json").read_text())
    proof = json.loads(to_json(cm, blob))
    _, _, channel = replay(proof)
    proof["stark_proof"]["queried_values"][1][5] ^= 1
    with pytest.raises(AssertionError, match="tree 1: RootMismatch"):
        check_decommitments(proof, channel, shapes)
    proof = json.loads(to_json(cm, blob))
    _, _, channel = replay(proof)
    shapes["components"][0]["n_trace_columns"] += 1
    with pytest.raises(AssertionError, match="tree 1"):
        check_decommitments(proof, channel, shapes)


@pytest.mark.gpu
def test_gpu_proof_transcript_replays(cm):
    inp = ch.GpuFibInput(cm, 1000)
    try:
        blob, _ = inp.prove()
    finally:
        inp.close()
    check(cm, blob)

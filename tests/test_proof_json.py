"""The reference's proof wire format (serde JSON of `Proof<Blake2sMerkleHasher>`, crates/prover/src/lib.rs:61-73):
`cm31_proof_to_json` / `cm31_proof_from_json` against the struct layouts of the reference source -- field names and order of
Proof, Claim (nine sections, opcode components nested in define_opcodes! order), PublicData, CommitmentSchemeProof
(pcs/prover.rs:156-165), FriProof / FriLayerProof (fri.rs:675-699), MerkleDecommitment (vcs/prover.rs:163-173), and serde's
encoding of M31 / QM31 / Blake2sHash / Option -- and blob -> JSON -> blob identity."""
import ctypes as C
import json
from pathlib import Path

import pytest

from tests import cairo_helpers as ch

GOLDEN = Path(__file__).resolve().parent / "golden"
P = (1 << 31) - 1


def to_json(cm, blob: bytes) -> str:
    buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
    n = C.c_size_t()
    cm.check(cm.lib().cm31_proof_to_json(buf, C.c_size_t(len(blob)), None, C.c_size_t(0), C.byref(n)))
    out = C.create_string_buffer(n.value + 1)
    cm.check(cm.lib().cm31_proof_to_json(buf, C.c_size_t(len(blob)), out, C.c_size_t(n.value + 1), C.byref(n)))
    return out.value.decode()


def from_json(cm, text: str) -> bytes:
    raw = text.encode()
    n = C.c_size_t()
    cm.check(cm.lib().cm31_proof_from_json(raw, C.c_size_t(len(raw)), None, C.c_size_t(0), C.byref(n)))
    out = (C.c_uint8 * n.value)()
    cm.check(cm.lib().cm31_proof_from_json(raw, C.c_size_t(len(raw)), out, C.c_size_t(n.value), C.byref(n)))
    return bytes(out)


@pytest.fixture(scope="module", params=[(ch.FIB, 10), (ch.U32_MIX, 3)])
def blob(request):
    return ch.oracle_program_prove(*request.param)[0]


def is_qm31(v):
    return (isinstance(v, list) and len(v) == 2 and all(isinstance(c, list) and len(c) == 2 for c in v)
            and all(isinstance(x, int) and 0 <= x < P for c in v for x in c))


def is_hash(v):
    return isinstance(v, list) and len(v) == 32 and all(isinstance(b, int) and 0 <= b < 256 for b in v)


def test_json_has_the_reference_struct_layout(cm, blob):
    proof = json.loads(to_json(cm, blob))
    shapes = json.loads((GOLDEN / "air_shapes_reference.json").read_text())["components"]
    opcode_components = [c["name"] for c in shapes if c["opcodes"]]
    other = [c["name"] for c in shapes if not c["opcodes"]]
    assert list(proof) == ["claim", "interaction_claim", "public_data", "stark_proof", "interaction_pow"]        # lib.rs:62-73
    for section, field in (("claim", "log_size"), ("interaction_claim", "claimed_sum")):
        assert list(proof[section]) == ["opcodes"] + other                                                      # components/mod.rs:28-38
        assert list(proof[section]["opcodes"]) == opcode_components                                             # opcodes/mod.rs:223-268
        for comp in list(proof[section]["opcodes"].values()) + [proof[section][k] for k in other]:
            assert list(comp) == [field]
    assert all(isinstance(c["log_size"], int) and c["log_size"] >= 4 for c in proof["claim"]["opcodes"].values())
    assert is_qm31(proof["interaction_claim"]["memory"]["claimed_sum"])
    pd = proof["public_data"]
    assert list(pd) == ["initial_registers", "final_registers", "clock", "initial_root", "final_root", "public_memory"]  # public_data.rs:213-227
    assert list(pd["initial_registers"]) == ["pc", "fp"] and list(pd["public_memory"]) == ["program", "input", "output"]
    for entry in pd["public_memory"]["program"]:
        assert entry is None or (isinstance(entry[0], int) and is_qm31(entry[1]) and isinstance(entry[2], int))
    sp = proof["stark_proof"]                                                                                   # StarkProof is a newtype
    assert list(sp) == ["config", "commitments", "sampled_values", "decommitments", "queried_values", "proof_of_work", "fri_proof"]
    assert sp["config"] == {"pow_bits": 16, "fri_config": {"log_blowup_factor": 1, "log_last_layer_degree_bound": 0, "n_queries": 80}}
    assert len(sp["commitments"]) == 4 and all(is_hash(h) for h in sp["commitments"])
    assert len(sp["sampled_values"]) == 4 and len(sp["sampled_values"][0]) == 7 and len(sp["sampled_values"][3]) == 4
    assert sum(len(t) for t in sp["sampled_values"][1:3]) == 1006 + 1180
    assert all(is_qm31(v) for col in sp["sampled_values"][2] for v in col)
    assert all(list(d) == ["hash_witness", "column_witness"] for d in sp["decommitments"])
    fp = sp["fri_proof"]
    assert list(fp) == ["first_layer", "inner_layers", "last_layer_poly"]
    assert list(fp["first_layer"]) == ["fri_witness", "decommitment", "commitment"] and is_hash(fp["first_layer"]["commitment"])
    assert list(fp["last_layer_poly"]) == ["coeffs", "log_size"] and len(fp["last_layer_poly"]["coeffs"]) == 1 << fp["last_layer_poly"]["log_size"]
    assert isinstance(proof["interaction_pow"], int)


def test_blob_json_blob_round_trip(cm, blob):
    text = to_json(cm, blob)
    assert from_json(cm, text) == blob
    # any serde-compatible writer's formatting is accepted (whitespace, indentation)
    assert from_json(cm, json.dumps(json.loads(text), indent=2)) == blob
    # and the proof that went through the wire format still verifies
    assert ch.oracle_cairo_verify(from_json(cm, text)) == 0


def test_malformed_json_is_refused(cm, blob):
    text = to_json(cm, blob)
    for bad in (text[:-1], text.replace('"claim"', '"claims"', 1), text.replace('"log_size":', '"log_size":-', 1), text + "x"):
        raw = bad.encode()
        n = C.c_size_t()
        assert cm.lib().cm31_proof_from_json(raw, C.c_size_t(len(raw)), None, C.c_size_t(0), C.byref(n)) != 0
        assert b"proof json" in cm.lib().cm31_last_error()


def test_fibonacci_public_memory_contents(cm):
    # crates/prover/tests/prover.rs:373-449: the proof's public memory holds (1) the return value in the output range,
    # (2) the input argument in the input range, (3) the program, word for word, in the program range
    n = 5
    blob, _ = ch.oracle_program_prove(ch.FIB, n)
    pm = json.loads(to_json(cm, blob))["public_data"]["public_memory"]
    values = lambda entries: [e[1] for e in entries if e is not None]
    assert values(pm["output"]) == [[[ch.fib_mod_p(n), 0], [0, 0]]], "the return value of fibonacci_loop(n)"
    assert values(pm["input"]) == [[[n, 0], [0, 0]]], "the input argument"
    # the program as the runner preloaded it (the VM's initial memory below the input cell)
    lib = cm.lib()
    h = C.c_void_p()
    cm.check(lib.cm31_test_vm_trace_create(C.c_uint32(ch.FIB), C.c_uint32(n), C.byref(h)))
    try:
        init = C.POINTER(C.c_uint32)()
        ranges = (C.c_uint32 * 6)()
        cm.check(lib.cm31_test_vm_trace_data(h, None, None, C.byref(init), ranges))
        program_start, program_end = ranges[0], ranges[1]
        words = [[[init[4 * a], init[4 * a + 1]], [init[4 * a + 2], init[4 * a + 3]]] for a in range(program_start, program_end)]
    finally:
        lib.cm31_test_vm_trace_destroy(h)
    got = values(pm["program"])
    assert len(got) == len(words) == program_end - program_start and got == words, "program in public memory == program loaded by the runner"
    assert [e[0] for e in pm["program"] if e is not None] == list(range(program_start, program_end))

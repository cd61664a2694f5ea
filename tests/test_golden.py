"""Golden vectors (tests/golden/golden.json): the reference's own KATs against the oracle, and
digests of oracle proofs pinning the protocol + component set (regenerate with make_golden.py)."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from tests import cairo_helpers as ch
from tests import oracle_lib as orc

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "golden.json").read_text())


def test_reference_kats():
    k = GOLD["reference_kats"]
    assert orc.blake2s(k["blake2s_a"]["input_ascii"].encode()).hex() == k["blake2s_a"]["digest"]
    assert hashlib.blake2s(b"a").hexdigest() == k["blake2s_a"]["digest"]
    g = orc.point_from_index(1)
    assert (int(g[0]), int(g[1])) == (k["m31_circle_gen"]["x"], k["m31_circle_gen"]["y"])


@pytest.mark.parametrize("name", sorted(GOLD["oracle_proofs"]))
def test_oracle_proof_digest(name):
    e = GOLD["oracle_proofs"][name]
    proof, _ = ch.oracle_program_prove(e["program"], e["n"])
    assert len(proof) == e["bytes"]
    assert hashlib.sha256(proof).hexdigest() == e["sha256"], "protocol or AIR changed: rerun tests/golden/make_golden.py if intended"


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD["oracle_proofs"]))
def test_gpu_proof_matches_golden_digest(cm, name):
    e = GOLD["oracle_proofs"][name]
    inp = ch.GpuFibInput(cm, e["n"], program=e["program"])
    try:
        got, _ = inp.prove()
    finally:
        inp.close()
    assert hashlib.sha256(got).hexdigest() == e["sha256"]

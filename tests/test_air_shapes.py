"""Ties the AIR set of this repo to the reference SOURCE: the shapes captured from `csrc/air/cairo_components.hpp`
(cm31_air_shapes: the InfoEvaluator-style pass over each `evaluate`) must equal what the reference's own files declare --
N_TRACE_COLUMNS, N_<RELATION>_LOOKUPS, the number of add_constraint / add_to_relation / next_trace_mask call sites of each
`fn evaluate`, the opcodes each component serves, the claim order and the relation sizes
(crates/prover/src/components/**, preprocessed/**, relations.rs, opcodes/mod.rs:223-268).  The reference side is the committed
fixture tests/golden/air_shapes_reference.json (made by tests/golden/make_air_shapes.py from /root/reference)."""
import ctypes as C
import json
import subprocess
import sys
from pathlib import Path

import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def reference():
    return json.loads((GOLDEN / "air_shapes_reference.json").read_text())


@pytest.fixture(scope="module")
def captured(cm):
    n = C.c_size_t()
    cm.check(cm.lib().cm31_air_shapes(None, C.c_size_t(0), C.byref(n)))
    buf = C.create_string_buffer(n.value + 1)
    cm.check(cm.lib().cm31_air_shapes(buf, C.c_size_t(n.value + 1), C.byref(n)))
    return json.loads(buf.value.decode())


def test_fixture_is_current_with_the_reference_checkout():
    if not Path("/root/reference/crates/prover/src").exists():
        pytest.skip("the reference checkout only exists in the build container")
    assert subprocess.run([sys.executable, str(GOLDEN / "make_air_shapes.py"), "--check"]).returncode == 0, \
        "tests/golden/air_shapes_reference.json is stale: rerun tests/golden/make_air_shapes.py"


def test_relation_sizes(reference, captured):  # relations.rs:7-44
    assert captured["relations"] == reference["relations"]
    assert reference["interaction_pow_bits"] == 2 and reference["preprocessed_trace_log_size"] == 20


def test_component_set_and_claim_order(reference, captured):  # opcodes/mod.rs:223-268, components/mod.rs:94-104
    assert [c["name"] for c in captured["components"]] == [c["name"] for c in reference["components"]]
    assert len(captured["components"]) == 34


def test_opcodes_served_by_each_component(reference, captured):  # define_opcodes! + instruction.rs ids
    for got, want in zip(captured["components"], reference["components"]):
        assert sorted(got["opcodes"]) == sorted(want["opcodes"]), got["name"]
    served = [o for c in captured["components"] for o in c["opcodes"]]
    assert len(served) == len(set(served)) == sum(len(c["opcodes"]) for c in reference["components"])


def test_shapes_match_the_reference_source(reference, captured):
    for got, want in zip(captured["components"], reference["components"]):
        name = want["name"]
        assert got["n_trace_columns"] == want["n_trace_columns"], name                      # N_TRACE_COLUMNS
        assert {k: v[0] for k, v in got["lookups"].items()} == want["lookups"], name        # N_<RELATION>_LOOKUPS
        n_lookups = sum(want["lookups"].values())
        assert got["n_lookups"] == n_lookups == want["static_add_to_relation"] or want["has_loop"], name
        assert got["n_lookups"] == n_lookups, name
        assert got["n_interaction_columns"] == 4 * ((n_lookups + 1) // 2), name             # SECURE_EXTENSION_DEGREE * ceil(k / 2)
        assert got["n_cumsum_columns"] == 4, name                                           # [-1, 0] mask on the last batch only
        if want.get("n_preprocessed_columns") is not None:
            assert got["n_preprocessed_columns"] == want["n_preprocessed_columns"], name
        else:
            assert got["n_preprocessed_columns"] == 0, name
        if not want["has_loop"]:  # call sites == calls
            assert want["static_next_trace_mask"] == want["n_trace_columns"], name           # sanity of the extraction itself
            assert got["n_constraints"] == want["static_add_constraint"] + (n_lookups + 1) // 2, name
        # widest tuple never exceeds the relation's size (combine would panic: logup.rs:96-111)
        for rel, (_, widest) in got["lookups"].items():
            assert widest <= captured["relations"][rel], (name, rel)


def test_survey_totals(captured):  # 1006 trace + 1180 interaction columns
    assert sum(c["n_trace_columns"] for c in captured["components"]) == 1006
    assert sum(c["n_interaction_columns"] for c in captured["components"]) == 1180

#!/usr/bin/env python3
"""Extracts the AIR shapes of the reference's 34 components from its SOURCE (crates/prover/src/components/**,
preprocessed/**, relations.rs, components/opcodes/mod.rs, crates/common/src/instruction.rs) into
tests/golden/air_shapes_reference.json.  tests/test_air_shapes.py compares what this repo captures from its own
restatement of the AIRs (cm31_air_shapes) with that file, and -- when /root/reference is present -- re-runs this
extraction to make sure the fixture is current.

Per component:  N_TRACE_COLUMNS, every N_<RELATION>_LOOKUPS constant, and -- counted inside the body of
`fn evaluate` -- the number of `add_constraint(`, `add_to_relation(` and `next_trace_mask()` call sites
(`static_*`; exact when the body has no loop, which `has_loop` records), plus the opcodes the component serves
(define_opcodes! in opcodes/mod.rs, ids from instruction.rs) and the claim order.
"""
from __future__ import annotations

import json
import re
import sys
from pathlib import Path

REF = Path("/root/reference")
PROVER = REF / "crates" / "prover" / "src"
OUT = Path(__file__).resolve().parent / "air_shapes_reference.json"

RELATION_OF_CONST = {"MEMORY": "memory", "REGISTERS": "registers", "MERKLE": "merkle", "POSEIDON2": "poseidon2",
                     "RANGE_CHECK_8": "range_check_8", "RANGE_CHECK_16": "range_check_16", "RANGE_CHECK_20": "range_check_20",
                     "BITWISE": "bitwise"}
RELATION_OF_STRUCT = {"Memory": "memory", "Registers": "registers", "Merkle": "merkle", "Poseidon2": "poseidon2",
                      "RangeCheck8": "range_check_8", "RangeCheck16": "range_check_16", "RangeCheck20": "range_check_20",
                      "Bitwise": "bitwise"}


def evaluate_body(src: str) -> str:
    """Text of `fn evaluate<E: EvalAtRow>(...) { ... }` (brace matched)."""
    m = re.search(r"fn evaluate<E: EvalAtRow>", src)
    if not m:
        raise ValueError("no evaluate")
    i = src.index("{", m.end())
    depth, j = 0, i
    while True:
        c = src[j]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return src[i:j + 1]
        j += 1


def strip_comments(src: str) -> str:
    return re.sub(r"//[^\n]*", "", src)


def const_env(src: str, env: dict) -> dict:
    out = dict(env)
    for m in re.finditer(r"^(?:pub )?const (N_[A-Z0-9_]+): usize\s*=\s*([^;]+);", src, re.M):
        name, expr = m.group(1), m.group(2)
        expr = re.sub(r"(\w+)\.div_ceil\((\d+)\)", r"(-(-\1 // \2))", expr)
        try:
            out[name] = int(eval(expr, {"__builtins__": {}}, out))  # arithmetic over earlier constants only
        except Exception:
            pass
    return out


def component_shape(path: Path, env: dict) -> dict:
    src = strip_comments(path.read_text())
    consts = const_env(src, env)
    body = evaluate_body(src)
    lookups = {RELATION_OF_CONST[k[2:-8]]: v for k, v in consts.items() if k.endswith("_LOOKUPS") and k[2:-8] in RELATION_OF_CONST}
    return {"file": str(path.relative_to(REF)), "n_trace_columns": consts.get("N_TRACE_COLUMNS"), "lookups": lookups,
            "static_add_constraint": body.count("add_constraint("), "static_add_to_relation": body.count("add_to_relation("),
            "static_next_trace_mask": body.count("next_trace_mask()"),
            "has_loop": bool(re.search(r"\bfor\b|from_fn|\.iter\(\)|\.map\(", body))}


def main():
    env = {"SECURE_EXTENSION_DEGREE": 4}
    p2 = (PROVER / "poseidon2.rs").read_text()
    for name in ("T", "FULL_ROUNDS", "PARTIAL_ROUNDS"):
        env[name] = int(re.search(rf"pub const {name}: usize = (\d+);", p2).group(1))
    opcode_ids = {m.group(1): int(m.group(2))
                  for m in re.finditer(r"^\s*(\w+) = (\d+) \{", (REF / "crates/common/src/instruction.rs").read_text(), re.M)}
    mod = strip_comments((PROVER / "components/opcodes/mod.rs").read_text())
    macro = mod[mod.rindex("define_opcodes!("):]
    macro = macro[:macro.index(");\n")]
    components = []
    for m in re.finditer(r"\(\s*\[([^\]]*)\],\s*(\w+)\s*\)", macro):
        ops = [o.strip() for o in m.group(1).split(",") if o.strip()]
        shape = component_shape(PROVER / "components/opcodes" / f"{m.group(2)}.rs", env)
        shape.update(name=m.group(2), opcodes=[opcode_ids[o] for o in ops], opcode_names=ops)
        components.append(shape)
    for name in ("memory", "merkle", "clock_update", "poseidon2"):  # components/mod.rs:94-104
        shape = component_shape(PROVER / "components" / f"{name}.rs", env)
        shape.update(name=name, opcodes=[], opcode_names=[])
        components.append(shape)
    rc = strip_comments((PROVER / "preprocessed/range_check/range_check_macro.rs").read_text())
    rc_body = evaluate_body(rc)
    rc_trace = int(re.search(r"let trace = vec!\[self\.log_size; (\d+)\];", rc).group(1))
    for bits in (8, 16, 20):
        components.append({"name": f"range_check_{bits}", "file": "crates/prover/src/preprocessed/range_check/range_check_macro.rs",
                           "n_trace_columns": rc_trace, "lookups": {f"range_check_{bits}": rc_body.count("add_to_relation(")},
                           "static_add_constraint": rc_body.count("add_constraint("), "static_add_to_relation": rc_body.count("add_to_relation("),
                           "static_next_trace_mask": rc_body.count("next_trace_mask()"), "has_loop": False, "opcodes": [], "opcode_names": [],
                           "n_preprocessed_columns": rc_body.count("get_preprocessed_column(")})
    bw = strip_comments((PROVER / "preprocessed/bitwise.rs").read_text())
    bw_body = evaluate_body(bw)
    components.append({"name": "bitwise", "file": "crates/prover/src/preprocessed/bitwise.rs",
                       "n_trace_columns": int(re.search(r"let trace = vec!\[self\.log_size; (\d+)\];", bw).group(1)),
                       "lookups": {"bitwise": bw_body.count("add_to_relation(")}, "static_add_constraint": bw_body.count("add_constraint("),
                       "static_add_to_relation": bw_body.count("add_to_relation("), "static_next_trace_mask": bw_body.count("next_trace_mask()"),
                       "has_loop": False, "opcodes": [], "opcode_names": [],
                       "n_preprocessed_columns": int(re.search(r"pub fn ids\(&self\) -> \[PreProcessedColumnId; (\d+)\]", bw).group(1))})
    relations = {RELATION_OF_STRUCT[m.group(1)]: int(m.group(2))
                 for m in re.finditer(r"^relation!\((\w+), (\d+)\);", (PROVER / "relations.rs").read_text(), re.M)}
    cfg = (PROVER / "prover_config.rs").read_text()
    out = {"source": "kkrt-labs/cairo-m crates/prover/src (extracted by tests/golden/make_air_shapes.py)", "relations": relations,
           "interaction_pow_bits": int(re.search(r"INTERACTION_POW_BITS: u32 = (\d+)", (PROVER / "relations.rs").read_text()).group(1)),
           "preprocessed_trace_log_size": int(re.search(r"PREPROCESSED_TRACE_LOG_SIZE: u32 = (\d+)", (PROVER / "prover.rs").read_text()).group(1)),
           "pcs_config_text": re.sub(r"\s+", " ", cfg[cfg.index("REGULAR_96_BITS"):cfg.index("};", cfg.index("REGULAR_96_BITS"))])[:400],
           "components": components}
    text = json.dumps(out, indent=1) + "\n"
    if "--check" in sys.argv:
        return 0 if OUT.read_text() == text else 1
    OUT.write_text(text)
    print(f"wrote {OUT}: {len(components)} components")
    return 0


if __name__ == "__main__":
    sys.exit(main())

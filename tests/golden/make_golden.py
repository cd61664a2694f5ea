"""Regenerates the `oracle_proofs` digests of tests/golden/golden.json (run after an intentional change
of the protocol or of the component set; the reference KATs are never regenerated)."""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tests import cairo_helpers as ch  # noqa: E402

path = Path(__file__).resolve().parent / "golden.json"
gold = json.loads(path.read_text())
for name, e in gold["oracle_proofs"].items():
    proof, _ = ch.oracle_program_prove(e["program"], e["n"])
    e["bytes"], e["sha256"] = len(proof), hashlib.sha256(proof).hexdigest()
path.write_text(json.dumps(gold, indent=1) + "\n")
print(json.dumps(gold["oracle_proofs"], indent=1))

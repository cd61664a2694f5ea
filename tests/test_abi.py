"""CPU-side checks of the drop-in boundary: libcm31.so loads, exports every symbol include/cm31.h declares, and exports NO
cm31_* symbol the header does not declare (nothing the product's own host path calls is hidden from an integrator)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "cm31.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cm31_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert len(syms) >= 20
    for must in ["cm31_interpolate_batch", "cm31_evaluate_batch", "cm31_blake2s_commit_layer",
                 "cm31_accumulate_quotients", "cm31_fold_line", "cm31_fold_circle_into_line", "cm31_grind_blake2s"]:
        assert must in syms


def test_library_exports_every_declared_symbol(cm):
    lib = cm.lib()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/cm31.h but not exported: {missing}"


def exported_symbols():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", str(ROOT / "cairo-m_b200" / "libcm31.so")], capture_output=True, text=True, check=True).stdout
    return sorted({line.split()[-1] for line in out.splitlines() if line.split() and line.split()[-1].startswith("cm31_")})


def test_every_exported_symbol_is_declared(cm):
    cm.lib()
    undeclared = sorted(set(exported_symbols()) - set(declared_symbols()))
    assert not undeclared, f"exported by libcm31.so but not declared in include/cm31.h: {undeclared}"


def test_test_conveniences_are_prefixed():
    # built-in programs, the tamper hook and the bring-up AIR are not part of the drop-in surface
    syms = declared_symbols()
    for s in syms:
        if any(k in s for k in ("fib_input", "program_input", "tamper", "vm_trace", "wide_fibonacci")):
            assert s.startswith("cm31_test_"), s


def test_errors_are_reported_not_swallowed(cm):
    # contract violations come back as a status + message (no silent CPU fallback)
    lib = cm.lib()
    out = ctypes.c_void_p()
    status = lib.cm31_twiddles_create(ctypes.c_uint32(99), ctypes.byref(out))
    assert status != 0
    assert b"log_size" in lib.cm31_last_error()

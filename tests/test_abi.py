"""CPU-side checks of the drop-in boundary: libcm31.so loads and exports every symbol include/cm31.h declares."""
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


def test_errors_are_reported_not_swallowed(cm):
    # contract violations come back as a status + message (no silent CPU fallback)
    lib = cm.lib()
    out = ctypes.c_void_p()
    status = lib.cm31_twiddles_create(ctypes.c_uint32(99), ctypes.byref(out))
    assert status != 0
    assert b"log_size" in lib.cm31_last_error()

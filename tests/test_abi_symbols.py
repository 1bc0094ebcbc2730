"""The C-ABI library loads on a CPU-only box and exports every symbol include/crm_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "crm_b200.h")).read()
    return sorted(set(re.findall(r"CRM_API\s+[\w\s\*]+?\b(crm_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for must in ("crm_create", "crm_destroy", "crm_setup", "crm_scan_interaction", "crm_scan_association", "crm_gemm",
                 "crm_lmm_fit_rotated", "crm_davies_pvalues", "crm_lrt_pvalues", "crm_last_error", "crm_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from cellregmap_b200 import _lib, build

    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert set(_lib.SIGNATURES) == set(_declared())
    assert _lib.load().crm_version() >= 100


def test_no_product_import_of_the_oracle():
    pkg = os.path.join(ROOT, "cellregmap_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"

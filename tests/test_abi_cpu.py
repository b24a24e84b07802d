"""CPU: the C-ABI library loads and exports every symbol include/cmrag.h declares."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "cmrag.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmr_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_header_symbols():
    from classmate_rag_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert declared, "no symbols parsed from cmrag.h"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in cmrag.h but not exported"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes table and header disagree"
    assert lib.cmr_version() >= 100


def test_no_oracle_import_in_product():
    """The product package must never route through the oracle."""
    for p in (ROOT / "classmate_rag_b200").rglob("*.py"):
        for line in p.read_text().splitlines():
            assert not re.match(r"\s*(from|import)\s+oracle\b", line), f"{p} imports the oracle: {line}"

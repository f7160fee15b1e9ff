"""The C-ABI library loads and exports every symbol include/gmrfb.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest


def _declared_symbols(entry):
    hdr = open(os.path.join(entry.ROOT, "include", "gmrfb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gmrfb_[A-Za-z0-9_]+)\s*\(", hdr)))


def test_exports_match_header(entry, pkg):
    names = _declared_symbols(entry)
    assert len(names) >= 40
    L = ctypes.CDLL(pkg._lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in gmrfb.h but not exported: {missing}"
    # the ctypes host binds exactly the declared set
    assert sorted(pkg._lib.SIGNATURES) == names


def test_version(pkg):
    assert pkg._lib.lib().gmrfb_version() >= 100


def test_no_cpu_fallback(pkg):
    """Without a usable GPU the context constructor must fail loudly; with one it must succeed."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(pkg.GmrfbError) as ei:
        pkg.Context(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle(entry):
    """The product package must not reference oracle/ in any way."""
    for root, _, files in os.walk(entry.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(root, f), errors="ignore").read()
                assert "gmrf_oracle" not in src and "liboracle" not in src and "oracle/" not in src, f

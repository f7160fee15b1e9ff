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


def test_no_exception_crosses_the_abi(entry):
    """A host allocation failure inside the symbolic analysis (std::bad_alloc under an address-space limit) comes back
    as GMRFB_ERR_ALLOC with a message - the process is not terminated.  Run in a child process: the limit stays there."""
    import subprocess
    import sys
    import textwrap

    code = textwrap.dedent(f"""
        import os, resource, sys
        sys.path.insert(0, {entry.ROOT!r})
        import __graft_entry__ as g
        pkg = g.load_pkg()
        prob = pkg.workloads.matern_posterior(400, obs_frac=0.1, q_eps=1e2, corr_range=0.05, seed=0)
        pkg._lib.lib()
        os.environ["GMRFB_ND_THREADS"] = "1"
        vms = int(open("/proc/self/statm").read().split()[0]) * os.sysconf("SC_PAGE_SIZE")
        resource.setrlimit(resource.RLIMIT_AS, (vms + 8 * 1024 * 1024, resource.getrlimit(resource.RLIMIT_AS)[1]))
        try:
            pkg.Symbolic(prob["Qpost"], coords=prob["nodes"], host_only=True)
            print("NO-ERROR")
        except pkg.GmrfbError as e:
            print("STATUS", e.status, str(e))
        except MemoryError:
            print("PYTHON-MEMORYERROR")
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "STATUS 3" in out.stdout and "bad_alloc" in out.stdout, out.stdout + out.stderr[-500:]

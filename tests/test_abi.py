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


def test_ctypes_signatures_match_the_header_types(entry, pkg):
    """Every entry of the ctypes table has the header's argument kinds in the header's order (an int32 / int64 or value /
    pointer mismatch would otherwise only show up as garbage in the upper half of a register), the same return kind, and
    the Structure mirrors have the header's field order and widths."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(entry.ROOT, "include", "gmrfb.h")).read(), flags=re.S)

    def c_kind(t):
        t = re.sub(r"\bconst\b", "", t).strip()
        if "*" in t:
            return "ptr"
        return {"int32_t": "i32", "gmrfb_status": "i32", "int": "i32", "int64_t": "i64", "uint64_t": "u64", "double": "f64",
                "size_t": "u64"}.get(t.split()[0], "?" + t)

    def py_kind(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or issubclass(t, ctypes._Pointer):
            return "ptr"
        return {ctypes.c_int32: "i32", ctypes.c_int64: "i64", ctypes.c_uint64: "u64", ctypes.c_double: "f64"}[t]

    checked = 0
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(gmrfb_[A-Za-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", hdr, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        kinds = []
        if args not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                kinds.append(c_kind(a if a.endswith("*") else re.sub(r"[A-Za-z_][A-Za-z0-9_]*$", "", a)))
        restype, argtypes = pkg._lib.SIGNATURES[name]
        assert [py_kind(t) for t in argtypes] == kinds, (name, kinds)
        assert py_kind(restype) == c_kind(ret), (name, ret)
        checked += 1
    assert checked == len(pkg._lib.SIGNATURES)
    for cname, S in (("gmrfb_analyze_opts", pkg._lib.AnalyzeOpts), ("gmrfb_sym_info", pkg._lib.SymInfo),
                     ("gmrfb_fac_info", pkg._lib.FacInfo), ("gmrfb_btd_info", pkg._lib.BtdInfo)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), hdr, flags=re.S).group(1)
        cf = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                names = [x.strip() for x in decl.split(",")]
                first = re.search(r"([A-Za-z_][A-Za-z0-9_]*)$", names[0]).group(1)
                cf += [(nm, c_kind(names[0][:-len(first)])) for nm in [first] + names[1:]]
        assert [(n, py_kind(t)) for n, t in S._fields_] == cf, cname


def test_header_is_plain_c_and_a_c_program_links(entry, pkg, tmp_path):
    """include/gmrfb.h compiles as C99 (what cgo / ccall / any FFI generator consumes) and a C program linked against
    libgmrfb.so calls through it: the version, and a context creation that reports - not aborts - when there is no GPU."""
    import subprocess

    src = tmp_path / "driver.c"
    src.write_text(
        '#include <stdio.h>\n#include "gmrfb.h"\n'
        "int main(void) {\n"
        "  gmrfb_ctx* ctx = NULL;\n"
        "  gmrfb_status st = gmrfb_ctx_create(0, &ctx);\n"
        '  printf("version %d status %d msg %s\\n", (int)gmrfb_version(), (int)st, st ? gmrfb_last_error(NULL) : "ok");\n'
        "  if (ctx) gmrfb_ctx_destroy(ctx);\n"
        "  return 0;\n}\n")
    inc = os.path.join(entry.ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)])
    exe = tmp_path / "driver"
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", entry.PKG_DIR, "-lgmrfb",
                           "-Wl,-rpath," + entry.PKG_DIR])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("version ") and int(out.stdout.split()[1]) >= 100
    if " status 0 " not in out.stdout:
        assert "no CPU fallback" in out.stdout

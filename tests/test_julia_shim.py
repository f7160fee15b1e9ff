"""Static checks of the Julia `ccall` shim (no Julia in the image, so the shim cannot be executed here): every symbol it
binds is declared in include/gmrfb.h with the same number of arguments, and no function calls a module-level helper
through a name that one of its own keyword arguments shadows (the `check::Bool` / `check(ctx, st)` bug of round 1)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_prototypes():
    src = open(os.path.join(ROOT, "include", "gmrfb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(gmrfb_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        protos[name] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def split_top_level(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return parts


def ccalls(text):
    """(symbol, number of argument types, number of arguments passed) of every ccall((:sym, libgmrfb), ...)."""
    out = []
    text = re.sub(r"#=.*?=#", "", text, flags=re.S)  # inline comments may hold commas
    for m in re.finditer(r"ccall\(\(:(gmrfb_[a-z0-9_]+),\s*libgmrfb\)", text):
        i = m.end()
        depth, j = 1, i
        while depth > 0:  # to the parenthesis that closes the ccall
            c = text[j]
            depth += c in "([{"
            depth -= c in ")]}"
            j += 1
        parts = split_top_level(text[i:j - 1])
        parts = [p for p in parts if p.strip()]  # leading "" before the first comma
        rettype, argtypes, args = parts[0], parts[1].strip(), parts[2:]
        assert argtypes.startswith("(") and argtypes.endswith(")"), (m.group(1), argtypes)
        inner = argtypes[1:-1].strip()
        ntypes = len([p for p in split_top_level(inner) if p.strip()]) if inner else 0
        out.append((m.group(1), ntypes, len(args), rettype.strip()))
    return out


def julia_sources():
    return sorted(glob.glob(os.path.join(ROOT, "julia", "*.jl")) + glob.glob(os.path.join(ROOT, "julia", "ext", "*.jl")))


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    seen = set()
    for path in julia_sources():
        for sym, ntypes, nargs, _ in ccalls(open(path).read()):
            assert sym in protos, f"{os.path.basename(path)} binds {sym}, which include/gmrfb.h does not declare"
            assert ntypes == protos[sym], f"{sym}: {ntypes} argument types in the ccall, {protos[sym]} parameters in the header"
            assert nargs == ntypes, f"{sym}: {nargs} arguments passed for {ntypes} argument types"
            seen.add(sym)
    # the shim covers the core of the ABI
    for must in ("gmrfb_ctx_create", "gmrfb_analyze", "gmrfb_factorize", "gmrfb_solve", "gmrfb_sample", "gmrfb_var_selinv",
                 "gmrfb_var_rbmc", "gmrfb_btd_factor", "gmrfb_btd_solve", "gmrfb_gn_create", "gmrfb_gn_optimize"):
        assert must in seen, must


def test_no_keyword_argument_shadows_a_called_helper():
    for path in julia_sources():
        text = open(path).read()
        helpers = set(re.findall(r"^function\s+([A-Za-z_][A-Za-z0-9_!]*)\s*\(", text, flags=re.M))
        helpers |= set(re.findall(r"^([A-Za-z_][A-Za-z0-9_!]*)\([^)]*\)\s*=", text, flags=re.M))
        for m in re.finditer(r"^function\s+[A-Za-z_.][A-Za-z0-9_.!]*\s*\((.*?)\)\n(.*?)^end", text, flags=re.M | re.S):
            sig, body = m.group(1), m.group(2)
            if ";" not in sig:
                continue
            kws = re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:::[^=,]+)?=", sig.split(";", 1)[1])
            for kw in kws:
                if kw in helpers:
                    calls = re.findall(r"(?<![A-Za-z0-9_.!])%s\(" % re.escape(kw), body)
                    assert not calls, f"{os.path.basename(path)}: keyword argument `{kw}` shadows the helper `{kw}(...)` it calls"


def test_extension_targets_the_blueprint_types_of_the_shim():
    shim = open(os.path.join(ROOT, "julia", "GMRFB200.jl")).read()
    ext = open(os.path.join(ROOT, "julia", "ext", "GMRFB200GMRFExt.jl")).read()
    for name in ("B200CholeskySolverBlueprint", "B200GNCholeskySolverBlueprint", "b200_cholesky", "var_rbmc", "var_selinv", "sample"):
        assert name in shim and name in ext, name
    proj = open(os.path.join(ROOT, "julia", "Project.toml")).read()
    assert "GMRFB200GMRFExt" in proj and "GaussianMarkovRandomFields" in proj

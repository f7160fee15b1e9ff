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
    for m in re.finditer(r"\b(gmrfb_[A-Za-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
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
    for m in re.finditer(r"ccall\(\(:(gmrfb_[A-Za-z0-9_]+),\s*libgmrfb\)", text):
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


# ---- argument and struct TYPES (an Int32 / Int64 or value / pointer mismatch is a silent ABI bug in a ccall) ----
def _c_kind(t):
    t = re.sub(r"\bconst\b", "", t).strip()
    if "*" in t:
        return "ptr"
    return {"int32_t": "i32", "gmrfb_status": "i32", "int": "i32", "int64_t": "i64", "double": "f64", "uint8_t": "u8",
            "size_t": "u64"}.get(t.split()[0] if t else t, "?" + t)


def _jl_kind(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t in ("Cstring", "Ptr"):
        return "ptr"
    return {"Int32": "i32", "Cint": "i32", "Int64": "i64", "Float64": "f64", "UInt8": "u8", "Csize_t": "u64"}.get(t, "?" + t)


def header_signatures():
    src = open(os.path.join(ROOT, "include", "gmrfb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    sigs = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(gmrfb_[A-Za-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        kinds = []
        if args not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                ty = re.sub(r"[A-Za-z_][A-Za-z0-9_]*$", "", a).strip() if not a.endswith("*") else a  # drop the parameter name
                kinds.append(_c_kind(ty))
        sigs[name] = (_c_kind(ret), kinds)
    return sigs


def test_ccall_argument_types_match_the_header():
    sigs = header_signatures()
    checked = 0
    for path in julia_sources():
        text = re.sub(r"#=.*?=#", "", open(path).read(), flags=re.S)
        for m in re.finditer(r"ccall\(\(:(gmrfb_[A-Za-z0-9_]+),\s*libgmrfb\)", text):
            i = m.end()
            depth, j = 1, i
            while depth > 0:
                c = text[j]
                depth += c in "([{"
                depth -= c in ")]}"
                j += 1
            parts = [p for p in split_top_level(text[i:j - 1]) if p.strip()]
            ret, argtypes = parts[0].strip(), parts[1].strip()[1:-1]
            jl = [_jl_kind(p) for p in split_top_level(argtypes) if p.strip()]
            cret, ckinds = sigs[m.group(1)]
            assert not any(k.startswith("?") for k in jl + ckinds + [cret]), (m.group(1), jl, ckinds, cret)
            assert _jl_kind(ret) == cret, f"{m.group(1)}: returns {cret} in the header, {ret} in the ccall"
            assert jl == ckinds, f"{m.group(1)}: header {ckinds}, ccall {jl}"
            checked += 1
    assert checked > 40


def test_julia_struct_mirrors_match_the_header():
    """AnalyzeOpts / SymInfo / FacInfo / BtdInfo are passed by reference: field order and widths must be the header's."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "gmrfb.h")).read(), flags=re.S)
    shim = open(os.path.join(ROOT, "julia", "GMRFB200.jl")).read()
    for cname, jname in (("gmrfb_analyze_opts", "AnalyzeOpts"), ("gmrfb_sym_info", "SymInfo"), ("gmrfb_fac_info", "FacInfo"),
                         ("gmrfb_btd_info", "BtdInfo")):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), hdr, flags=re.S).group(1)
        cf = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                names = [x.strip() for x in decl.split(",")]  # "int64_t b, nblocks"
                first = re.search(r"([A-Za-z_][A-Za-z0-9_]*)$", names[0]).group(1)
                ty = names[0][:-len(first)]
                cf += [(nm, _c_kind(ty)) for nm in [first] + names[1:]]
        jbody = re.search(r"^struct %s\n(.*?)^end" % jname, shim, flags=re.S | re.M).group(1)
        jf = [(n, _jl_kind(t)) for n, t in re.findall(r"([A-Za-z_][A-Za-z0-9_]*)::([A-Za-z0-9_{}]+)", jbody)]
        assert jf == cf, f"{jname} vs {cname}: {jf} != {cf}"


def test_extension_is_keyed_to_the_package_the_reference_depends_on():
    """The weak dependency must carry the UUID of GaussianMarkovRandomFields in the reference's Project.toml:19, or Pkg
    never loads the extension."""
    proj = open(os.path.join(ROOT, "julia", "Project.toml")).read()
    assert 'GaussianMarkovRandomFields = "d5f06795-35bb-4323-9f0b-405ef76cfc5b"' in proj

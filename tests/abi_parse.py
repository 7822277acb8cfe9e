"""Parsers shared by the ABI tests: the struct layouts of include/mom6cu.h, the bind(C) types and interfaces of
fortran/mom6cu_interface.F90, and the public lists / dummy-argument lists of Fortran modules."""
import re


def _strip_c_comments(s):
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    return re.sub(r"//[^\n]*", " ", s)


def c_structs(path):
    """{name: [(kind, member, count)]} with kind in int | double | ptr | i64 | size_t | struct:<name>."""
    s = _strip_c_comments(open(path).read())
    out = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)?\s*\{(.*?)\}\s*(\w+)\s*;", s, flags=re.S):
        name, body = m.group(3), m.group(2)
        mem = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            mm = re.match(r"(const\s+)?(unsigned\s+)?(struct\s+)?(\w+(?:\s+long)?)\s*(.*)$", decl)
            base, rest = mm.group(4), mm.group(5)
            for item in [x.strip() for x in rest.split(",")]:
                ptr = item.count("*")
                item = re.sub(r"\bconst\b", "", item.replace("*", " ")).strip()
                am = re.match(r"(\w+)\s*(?:\[(\w+)\])?$", item)
                nm, cnt = am.group(1), am.group(2)
                if ptr:
                    kind = "ptr"
                elif base == "int":
                    kind = "int"
                elif base == "double":
                    kind = "double"
                elif base in ("int64_t", "long long", "uint64_t"):
                    kind = "i64"
                elif base == "size_t":
                    kind = "size_t"
                else:
                    kind = "struct:" + base
                mem.append((kind, nm, cnt or "1"))
        out[name] = mem
    return out


def _join_continuations(src):
    lines, cur = [], ""
    for ln in src.splitlines():
        ln = ln.split("!")[0].rstrip() if not ln.lstrip().startswith("!") else ""
        if ln.lstrip().startswith("#"):
            continue
        if cur:
            if not ln.strip():      # a comment-only or blank line inside a continued statement does not end it
                continue
            ln = ln.lstrip()
            if ln.startswith("&"):
                ln = ln[1:]
        if ln.endswith("&"):
            cur += ln[:-1] + " "
            continue
        lines.append(cur + ln)
        cur = ""
    return lines


def fortran_bindc_types(path):
    """{name: [(kind, member, count)]} of every `type, bind(C) :: name` in a Fortran source."""
    out, cur = {}, None
    for ln in _join_continuations(open(path).read()):
        t = ln.strip()
        m = re.match(r"type\s*,\s*bind\(C\)\s*::\s*(\w+)", t, flags=re.I)
        if m:
            cur = m.group(1); out[cur] = []
            continue
        if cur and re.match(r"end\s+type", t, flags=re.I):
            cur = None
            continue
        if cur and "::" in t:
            spec, names = t.split("::", 1)
            spec = spec.strip().lower().replace(" ", "")
            if spec.startswith("integer(c_int64_t)"):
                kind = "i64"
            elif spec.startswith("integer(c_int)"):
                kind = "int"
            elif spec.startswith("integer(c_size_t)"):
                kind = "size_t"
            elif spec.startswith("real(c_double)"):
                kind = "double"
            elif spec.startswith("type(c_ptr)") or spec.startswith("type(c_funptr)"):
                kind = "ptr"
            elif spec.startswith("type("):
                kind = "struct:" + re.match(r"type\((\w+)\)", spec).group(1)
            else:
                raise ValueError(f"{path}: unparsed member spec {spec!r}")
            for item in re.split(r",(?![^()]*\))", names):
                item = item.strip()
                am = re.match(r"(\w+)\s*(?:\((\w+)\))?$", item)
                out[cur].append((kind, am.group(1), am.group(2) or "1"))
    return out


def fortran_public_api(path):
    """(module name, public names, {procedure: [dummy argument names]}) of a Fortran module source."""
    lines = _join_continuations(open(path).read())
    mod, public, procs, default_public = None, [], {}, True
    depth_contains = False
    for ln in lines:
        t = ln.strip()
        m = re.match(r"module\s+(\w+)\s*$", t, flags=re.I)
        if m and mod is None and not re.match(r"module\s+procedure", t, flags=re.I):
            mod = m.group(1)
        if re.match(r"implicit\s+none\s*;\s*private", t, flags=re.I) or re.match(r"private\s*$", t, flags=re.I):
            default_public = False
        if re.match(r"contains\s*$", t, flags=re.I):
            depth_contains = True
        if not depth_contains:
            m = re.match(r"public\s*(?:::)?\s*(.+)$", t, flags=re.I)
            if m and not t.lower().startswith("public ::") or (m and t.lower().startswith("public")):
                public += [x.strip() for x in m.group(1).split(",") if x.strip()]
            m = re.match(r"type\s*,\s*public\s*::\s*(\w+)", t, flags=re.I)
            if m:
                public.append(m.group(1))
            m = re.match(r"interface\s+(\w+)", t, flags=re.I)
        m = re.match(r"(?:(?:pure|elemental|recursive|real|integer|logical)\s+)*(subroutine|function)\s+(\w+)\s*\(([^)]*)\)", t, flags=re.I)
        if m:
            procs[m.group(2)] = [a.strip() for a in m.group(3).split(",") if a.strip()]
    seen, pub = set(), []
    for p in public:
        if p.lower() not in seen:
            seen.add(p.lower()); pub.append(p)
    return mod, pub, procs


def c_functions(path):
    """Names of the functions include/mom6cu.h declares (everything `name(` at the start of a declaration that returns int / double / void / char*)."""
    s = _strip_c_comments(open(path).read())
    s = re.sub(r"typedef\s+struct.*?\}\s*\w+\s*;", " ", s, flags=re.S)
    return sorted(set(re.findall(r"\b(mom6cu_\w+)\s*\(", s)))


def fortran_bindc_interfaces(path):
    """{C name: fortran name} of every bind(C, name="...") interface."""
    out = {}
    for ln in _join_continuations(open(path).read()):
        m = re.search(r"(?:function|subroutine)\s+(\w+)\s*\(.*bind\(C\s*,\s*name\s*=\s*\"(\w+)\"\)", ln, flags=re.I)
        if m:
            out[m.group(2)] = m.group(1)
    return out


def fortran_types(paths):
    """{type name (lower): {member (lower): member's derived type name (lower) or None}} for every derived type defined in the sources."""
    db = {}
    for path in paths:
        cur = None
        for ln in _join_continuations(open(path, errors="replace").read()):
            t = ln.strip()
            m = re.match(r"type\s*(?:,\s*(?:public|private|bind\(C\)|abstract|extends\(\w+\)))*\s*(?:::)?\s*(\w+)\s*(?:;.*)?$", t, flags=re.I)
            if m and not re.match(r"type\s*\(", t, flags=re.I) and cur is None and m.group(1).lower() not in ("is",):
                cur = m.group(1).lower(); db.setdefault(cur, {})
                continue
            if cur and re.match(r"end\s*type", t, flags=re.I):
                cur = None
                continue
            if cur and "::" in t:
                spec, names = t.split("::", 1)
                dm = re.match(r"\s*(?:type|class)\s*\(\s*(\w+)\s*\)", spec, flags=re.I)
                dtype = dm.group(1).lower() if dm else None
                for item in re.split(r",(?![^()]*\))", names):
                    nm = re.match(r"\s*(\w+)", item)
                    if nm:
                        db[cur][nm.group(1).lower()] = dtype
    return db


def fortran_procedures(text):
    """[(name, [dummy names], {dummy (lower): derived type name (lower) or None}, body text)] of the subroutines / functions in a source text."""
    lines = _join_continuations(text)
    out, cur = [], None
    for ln in lines:
        t = ln.strip()
        m = re.match(r"(?:(?:pure|elemental|recursive|logical|integer|real)\s+)*(subroutine|function)\s+(\w+)\s*\(([^)]*)\)", t, flags=re.I)
        if m and cur is None:
            cur = [m.group(2), [a.strip() for a in m.group(3).split(",") if a.strip()], {}, []]
            continue
        if cur is not None:
            if re.match(rf"end\s+(subroutine|function)\s+{cur[0]}\b", t, flags=re.I):
                out.append((cur[0], cur[1], cur[2], "\n".join(cur[3]))); cur = None
                continue
            cur[3].append(t)
            if "::" in t:
                spec, names = t.split("::", 1)
                dm = re.match(r"\s*(?:type|class)\s*\(\s*(\w+)\s*\)", spec, flags=re.I)
                for item in re.split(r",(?![^()]*\))", names):
                    nm = re.match(r"\s*(\w+)", item)
                    if nm:
                        cur[2][nm.group(1).lower()] = dm.group(1).lower() if dm else None
    return out

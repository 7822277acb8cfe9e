"""Fortran-subset -> Python translator for the reference's own sources.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).

There is no Fortran compiler in this image, so the reference cannot be built here.  This module is the stand-in: it reads the
.F90 files where they lie under /root/reference (nothing is copied into the repo; the generated Python lives in memory), runs the few cpp directives MOM6 uses (MOM_memory.h, symmetric dynamic memory), parses the subset of
free-form Fortran 90 the hot-path modules are written in, and emits Python in which every floating-point expression is
evaluated in binary64 in exactly the order the source text prescribes.  tests/test_reference_f90.py runs those translated
reference routines on seeded inputs and compares the C++ oracle with them bit for bit; that is what pins the oracle.

Subset: modules with contained subroutines/functions; real/integer/logical/character/type declarations with dimension,
intent, optional, pointer, parameter; do / do while / if / select case / exit / cycle / return / call / assignment (scalar,
element, section, whole array); derived-type components; optional and keyword arguments; scalar intent(out) arguments
(copy-out); array sections as actual arguments (views); generic interfaces resolved by rank/presence at run time.
Anything else is translated into a statement that raises when (and only when) it is executed."""
import os
import re

from .rt import mangle

# ---------------------------------------------------------------------------------------------------------------------------
# cpp
# ---------------------------------------------------------------------------------------------------------------------------


class Cpp:
    def __init__(self, include_dirs, defines=None):
        self.inc = list(include_dirs)
        self.macros = dict(defines or {})  # name -> (params or None, body)

    def _find(self, name, cur):
        for d in [os.path.dirname(cur)] + self.inc:
            p = os.path.join(d, name)
            if os.path.exists(p):
                return p
        return None

    def expand(self, line):
        if not self.macros:
            return line
        for _ in range(12):
            changed = False
            for name, (params, body) in self.macros.items():
                if name not in line:
                    continue
                if params is None:
                    new = re.sub(r"\b" + re.escape(name) + r"\b", lambda m: body, line)
                else:
                    new = self._expand_fn(line, name, params, body)
                if new != line:
                    line, changed = new, True
            if not changed:
                break
        return line

    @staticmethod
    def _expand_fn(line, name, params, body):
        out, pos = "", 0
        for m in re.finditer(r"\b" + re.escape(name) + r"\(", line):
            if m.start() < pos:
                continue
            depth, i = 1, m.end()
            while i < len(line) and depth:
                depth += (line[i] == "(") - (line[i] == ")")
                i += 1
            args = _split_top(line[m.end():i - 1]) if len(params) > 1 else [line[m.end():i - 1].strip()]
            rep = body
            for p, a in zip(params, args):
                rep = re.sub(r"\b" + re.escape(p) + r"\b", lambda mm: a, rep)
            out += line[pos:m.start()] + rep
            pos = i
        return out + line[pos:]

    def run(self, path):
        """-> list of (lineno, text) of path with directives applied (line numbers of the including file for included text)"""
        out = []
        self._run(path, out, None)
        return out

    def _run(self, path, out, at):
        stack = []  # (taking, taken_before)
        with open(path) as f:
            lines = f.read().split("\n")
        for n, raw in enumerate(lines, 1):
            ln = at if at is not None else n
            s = raw.strip()
            live = all(t for t, _ in stack)
            if s.startswith("#"):
                d = s[1:].strip()
                m = re.match(r"(\w+)\s*(.*)", d)
                if not m:
                    continue
                cmd, rest = m.group(1), m.group(2)
                if cmd in ("ifdef", "ifndef"):
                    on = (rest.split()[0] in self.macros) == (cmd == "ifdef")
                    stack.append((on, on))
                elif cmd == "if":
                    mm = re.match(r"defined\s*\(?\s*(\w+)\s*\)?\s*$", rest)
                    on = bool(mm) and mm.group(1) in self.macros
                    stack.append((on, on))
                elif cmd == "else":
                    t, tb = stack.pop()
                    stack.append((not tb, True))
                elif cmd == "endif":
                    stack.pop()
                elif not live:
                    continue
                elif cmd == "define":
                    mm = re.match(r"(\w+)\(([^)]*)\)\s*(.*)", rest)
                    if mm:
                        self.macros[mm.group(1)] = ([p.strip() for p in mm.group(2).split(",")], mm.group(3).strip())
                    else:
                        mm = re.match(r"(\w+)\s*(.*)", rest)
                        self.macros[mm.group(1)] = (None, mm.group(2).strip())
                elif cmd == "undef":
                    self.macros.pop(rest.split()[0], None)
                elif cmd == "include":
                    name = rest.strip().strip('<>"')
                    p = self._find(name, path)
                    if p is not None:
                        self._run(p, out, ln)
                continue
            if live:
                out.append((ln, self.expand(raw)))


# ---------------------------------------------------------------------------------------------------------------------------
# statements
# ---------------------------------------------------------------------------------------------------------------------------

def _strip_comment(line):
    q = None
    for i, c in enumerate(line):
        if q:
            if c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c == "!":
            return line[:i]
    return line


def _lower_outside_strings(s):
    out, q = [], None
    for c in s:
        if q:
            out.append(c)
            if c == q:
                q = None
        else:
            if c in "'\"":
                q = c
                out.append(c)
            else:
                out.append(c.lower())
    return "".join(out)


def _split_semicolons(s):
    parts, cur, q = [], [], None
    for c in s:
        if q:
            cur.append(c)
            if c == q:
                q = None
        elif c in "'\"":
            q = c
            cur.append(c)
        elif c == ";":
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(c)
    parts.append("".join(cur))
    return [p.strip() for p in parts if p.strip()]


def statements(lines):
    """(lineno, text) source lines -> list of (lineno, statement) with continuations joined, comments removed, lower-cased"""
    out, cur, start = [], "", None
    for ln, raw in lines:
        s = _strip_comment(raw).rstrip()
        if not s.strip():
            continue
        t = s.strip()
        if cur:
            if t.startswith("&"):
                t = t[1:]
            cur += " " + t.lstrip() if not cur.endswith(("'", '"')) or True else t
        else:
            cur, start = t, ln
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        for st in _split_semicolons(cur):
            out.append((start, _lower_outside_strings(st)))
        cur = ""
    if cur:
        out.append((start, _lower_outside_strings(cur)))
    return out


# ---------------------------------------------------------------------------------------------------------------------------
# expressions
# ---------------------------------------------------------------------------------------------------------------------------
_TOK = re.compile(r"""
    (?P<ws>\s+)
  | (?P<num>(?:\d+\.(?![a-z]+\.)\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_\w+)?)
  | (?P<dot>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
  | (?P<name>[a-z_]\w*)
  | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<op>\*\*|//|==|/=|<=|>=|=>|\(/|/\)|[-+*/<>=(),:%\[\]])
""", re.X)


def tokenize(s):
    toks, i = [], 0
    while i < len(s):
        m = _TOK.match(s, i)
        if not m:
            raise SyntaxError(f"cannot tokenize {s[i:i + 20]!r} in {s!r}")
        i = m.end()
        k = m.lastgroup
        if k == "ws":
            continue
        v = m.group(k)
        if k == "op" and v == "(/":
            # "(/" is an array-constructor bracket only if it is not "(" followed by a "/=" or a division; MOM6 uses [ ] or (/ /)
            # with the slash directly attached, and never writes "(/=": accept
            pass
        toks.append((k, v))
    toks.append(("end", ""))
    return toks


_DOTREL = {".eq.": "==", ".ne.": "/=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i]

    def next(self):
        t = self.t[self.i]
        self.i += 1
        return t

    def accept(self, k, v=None):
        t = self.t[self.i]
        if t[0] == k and (v is None or t[1] == v):
            self.i += 1
            return t
        return None

    def expect(self, k, v=None):
        t = self.accept(k, v)
        if t is None:
            raise SyntaxError(f"expected {v or k}, found {self.t[self.i]} in {' '.join(x[1] for x in self.t)}")
        return t

    def at_end(self):
        return self.t[self.i][0] == "end"

    # expr := equiv
    def expr(self):
        return self.p_eqv()

    def p_eqv(self):
        l = self.p_or()
        while self.peek() in (("dot", ".eqv."), ("dot", ".neqv.")):
            op = self.next()[1]
            l = ("bin", op, l, self.p_or())
        return l

    def p_or(self):
        l = self.p_and()
        while self.accept("dot", ".or."):
            l = ("bin", ".or.", l, self.p_and())
        return l

    def p_and(self):
        l = self.p_not()
        while self.accept("dot", ".and."):
            l = ("bin", ".and.", l, self.p_not())
        return l

    def p_not(self):
        if self.accept("dot", ".not."):
            return ("un", ".not.", self.p_not())
        return self.p_rel()

    def p_rel(self):
        l = self.p_concat()
        t = self.peek()
        if t[0] == "op" and t[1] in ("==", "/=", "<", "<=", ">", ">="):
            self.next()
            return ("bin", t[1], l, self.p_concat())
        if t[0] == "dot" and t[1] in _DOTREL:
            self.next()
            return ("bin", _DOTREL[t[1]], l, self.p_concat())
        return l

    def p_concat(self):
        l = self.p_add()
        while self.accept("op", "//"):
            l = ("bin", "//", l, self.p_add())
        return l

    def p_add(self):
        t = self.peek()
        if t == ("op", "-"):
            self.next()
            l = ("un", "-", self.p_mul())
        elif t == ("op", "+"):
            self.next()
            l = self.p_mul()
        else:
            l = self.p_mul()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.next()[1]
            l = ("bin", op, l, self.p_mul())
        return l

    def p_mul(self):
        l = self.p_pow()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.next()[1]
            l = ("bin", op, l, self.p_pow())
        return l

    def p_pow(self):
        b = self.p_primary()
        if self.accept("op", "**"):
            # right associative; a unary minus may follow ** in practice ("x**-2" is non-standard but "x**(-2)" is parsed below)
            if self.peek() == ("op", "-"):
                self.next()
                return ("bin", "**", b, ("un", "-", self.p_pow()))
            return ("bin", "**", b, self.p_pow())
        return b

    def p_args(self, close=")"):
        args = []
        if self.accept("op", close):
            return args
        while True:
            args.append(self.p_arg())
            if self.accept("op", ","):
                continue
            self.expect("op", close)
            return args

    def p_arg(self):
        # keyword argument
        if self.peek()[0] == "name" and self.t[self.i + 1] == ("op", "="):
            name = self.next()[1]
            self.next()
            return ("kw", name, self.expr())
        lo = None
        if self.peek() != ("op", ":"):
            lo = self.expr()
            if self.peek() != ("op", ":"):
                return lo
        self.expect("op", ":")
        hi = st = None
        if self.peek() not in (("op", ","), ("op", ")"), ("op", ":")):
            hi = self.expr()
        if self.accept("op", ":"):
            st = self.expr()
        return ("slice", lo, hi, st)

    def p_primary(self):
        t = self.next()
        k, v = t
        if k == "num":
            return ("num", v)
        if k == "str":
            return ("str", v)
        if k == "dot" and v in (".true.", ".false."):
            return ("log", v == ".true.")
        if k == "op" and v == "(":
            e = self.expr()
            self.expect("op", ")")
            return ("paren", e)
        if k == "op" and v in ("(/", "["):
            items = self.p_args("/)" if v == "(/" else "]")
            return ("arr", items)
        if k == "op" and v in ("-", "+"):
            x = self.p_primary()
            return ("un", v, x) if v == "-" else x
        if k == "name":
            parts = []
            name = v
            while True:
                args = None
                if self.accept("op", "("):
                    args = self.p_args()
                    # a second parenthesised list (substring or array-of-array) is not used by the reference
                parts.append((name, args))
                if self.accept("op", "%"):
                    name = self.expect("name")[1]
                    continue
                break
            return ("ref", parts)
        raise SyntaxError(f"unexpected token {t} in {' '.join(x[1] for x in self.t)}")


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if not p.at_end():
        raise SyntaxError(f"trailing tokens in expression {s!r}")
    return e


# ---------------------------------------------------------------------------------------------------------------------------
# program structure
# ---------------------------------------------------------------------------------------------------------------------------
def _split_top(s, sep=","):
    parts, cur, depth, q = [], [], 0, None
    for c in s:
        if q:
            cur.append(c)
            if c == q:
                q = None
            continue
        if c in "'\"":
            q = c
        elif c in "([":
            depth += 1
        elif c in ")]":
            depth -= 1
        if c == sep and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(c)
    parts.append("".join(cur).strip())
    return parts


def _match_paren(s, i):
    """s[i] == '(' -> index just past the matching ')'"""
    depth, q = 0, None
    for j in range(i, len(s)):
        c = s[j]
        if q:
            if c == q:
                q = None
            continue
        if c in "'\"":
            q = c
        elif c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return j + 1
    raise SyntaxError("unbalanced parentheses in " + s)


_DECL = re.compile(r"^(real|integer|logical|character|type|class|double\s+precision|complex|procedure)\b")


class Var:
    __slots__ = ("name", "base", "tname", "dims", "intent", "optional", "pointer", "parameter", "init", "allocatable", "dummy", "target")

    def __init__(self, name, base, tname=None):
        self.name, self.base, self.tname = name, base, tname
        self.dims = None
        self.intent = None
        self.optional = self.pointer = self.parameter = self.allocatable = self.dummy = self.target = False
        self.init = None

    @property
    def kind(self):
        return {"real": "r", "integer": "i", "logical": "l", "character": "s", "type": "o", "class": "o"}.get(self.base, "r")


def parse_decl(st):
    """a declaration statement -> list of Var, or None if st is not a declaration"""
    m = _DECL.match(st)
    if not m:
        return None
    base = m.group(1)
    if base.startswith("double"):
        base = "real"
    i = m.end()
    rest = st[i:].lstrip()
    tname = None
    if rest.startswith("("):
        j = _match_paren(rest, 0)
        spec = rest[1:j - 1].strip()
        if base in ("type", "class"):
            tname = spec
        rest = rest[j:].lstrip()
    elif base in ("type", "class"):
        return None  # "type :: name" / "type, public :: name" is a type definition, handled elsewhere
    if "::" in rest:
        # split at the first top-level '::'
        depth, k = 0, None
        for p in range(len(rest) - 1):
            c = rest[p]
            depth += (c == "(") - (c == ")")
            if depth == 0 and rest[p:p + 2] == "::":
                k = p
                break
        attrs, ents = rest[:k], rest[k + 2:]
    else:
        if rest.startswith(","):
            return None
        attrs, ents = "", rest
        if not re.match(r"[a-z_]", ents):
            return None
    dims = None
    intent = None
    flags = set()
    for a in _split_top(attrs):
        a = a.strip()
        if not a:
            continue
        if a.startswith("dimension"):
            dims = _split_top(a[a.index("(") + 1:a.rindex(")")])
        elif a.startswith("intent"):
            intent = a[a.index("(") + 1:a.rindex(")")].replace(" ", "")
        else:
            flags.add(a.split("(")[0].strip())
    out = []
    for e in _split_top(ents):
        init = None
        mm = re.match(r"([a-z_]\w*)\s*(.*)$", e, re.S)
        name, tail = mm.group(1), mm.group(2).strip()
        edims = dims
        if tail.startswith("("):
            j = _match_paren(tail, 0)
            edims = _split_top(tail[1:j - 1])
            tail = tail[j:].strip()
        if tail.startswith("*"):
            tail = re.sub(r"^\*\s*(\(\s*[^)]*\)|\w+)", "", tail).strip()
        if tail.startswith("=>"):
            init = ("ptr", tail[2:].strip())
        elif tail.startswith("="):
            init = ("val", tail[1:].strip())
        v = Var(name, base, tname)
        v.dims = edims
        v.intent = intent
        v.optional = "optional" in flags
        v.pointer = "pointer" in flags
        v.parameter = "parameter" in flags
        v.allocatable = "allocatable" in flags
        v.init = init
        out.append(v)
    return out


class Proc:
    def __init__(self, name, kind, args, result, module):
        self.name, self.kind, self.args, self.result, self.module = name, kind, args, result, module
        self.vars = {}      # name -> Var
        self.body = []      # nested statement tree
        self.line = 0
        self.uses = []
        self.elemental = False
        self.internal = {}  # name -> Proc: the internal procedures after this one's CONTAINS
        self.host = None    # the Proc an internal procedure lies in

    def out_scalars(self):
        """dummy arguments that are scalars of intrinsic type and may be defined by the procedure (copied out to the caller)"""
        if self.kind == "function":
            return []
        out = []
        for a in self.args:
            v = self.vars.get(a)
            if v is None or v.dims is not None or v.base in ("type", "class", "procedure") or v.pointer:
                continue
            if v.intent == "in":
                continue
            out.append(a)
        return out


class Module:
    def __init__(self, name, path):
        self.name, self.path = name, path
        self.vars = {}
        self.procs = {}
        self.generics = {}  # generic name -> [specific names]
        self.uses = []      # (module, only-list or None as [(local, remote)])
        self.types = {}     # type name -> {component name -> Var}
        self.type_ext = {}  # type name -> parent type name or None
        self.type_binds = {}  # type name -> {binding name -> procedure name}
        self.operators = {}   # '+', '-', '=' ... -> [module procedures that overload it for derived types]


def find_assign_simple(st):
    depth = 0
    for k, c in enumerate(st):
        depth += (c == "(") - (c == ")")
        if c == "=" and depth == 0 and st[k:k + 2] != "=>" and st[k - 1:k + 1] not in ("==", "/=", "<=", ">="):
            return k
    return -1


_HDR = re.compile(r"^(?:(?:pure|elemental|recursive|impure|real|integer|logical)(?:\s*\([^)]*\))?\s+)*(subroutine|function)\s+([a-z_]\w*)\s*(\(.*)?$")


def parse_use(st):
    m = re.match(r"use\s*(?:,\s*intrinsic\s*)?(?:::)?\s*([a-z_]\w*)\s*(?:,\s*only\s*:\s*(.*))?$", st)
    if not m:
        return None
    only = None
    if m.group(2) is not None:
        only = []
        for it in _split_top(m.group(2)):
            if "=>" in it:
                l, r = it.split("=>")
                only.append((l.strip(), r.strip()))
            elif it:
                only.append((it.strip(), it.strip()))
    return (m.group(1), only)


def parse_file(path, cpp):
    sts = statements(cpp.run(path))
    mods = []
    i = 0
    n = len(sts)
    mod = None
    while i < n:
        ln, st = sts[i]
        m = re.match(r"module\s+([a-z_]\w*)$", st)
        if m and not st.startswith("module procedure"):
            mod = Module(m.group(1), path)
            mods.append(mod)
            i = _parse_module_spec(sts, i + 1, mod)
            continue
        if re.match(r"end\s*module", st):
            mod = None
            i += 1
            continue
        h = _HDR.match(st)
        if h and mod is not None:
            i = _parse_proc(sts, i, mod)
            continue
        i += 1
    return mods


def _parse_module_spec(sts, i, mod):
    n = len(sts)
    while i < n:
        ln, st = sts[i]
        if st == "contains" or re.match(r"end\s*module", st):
            return i + (st == "contains")
        u = parse_use(st) if st.startswith("use") else None
        if u:
            mod.uses.append(u)
            i += 1
            continue
        m = re.match(r"type\s*(?:,\s*[^:]*)?(?:::)?\s*([a-z_]\w*)$", st)
        if m and not st.startswith("type("):
            comps = {}
            ext = re.search(r"extends\s*\(\s*([a-z_]\w*)\s*\)", st)
            binds = {}
            i += 1
            in_bind = False
            while not re.match(r"end\s*type", sts[i][1]):
                t = sts[i][1]
                if t == "contains":
                    in_bind = True
                elif in_bind:
                    mm = re.match(r"procedure\s*(?:\([^)]*\))?\s*(?:,[^:]*)?::\s*(.*)$", t)
                    if mm and "deferred" not in t:
                        for it in _split_top(mm.group(1)):
                            if "=>" in it:
                                l, r = it.split("=>")
                                binds[l.strip()] = r.strip()
                            else:
                                binds[it.strip()] = it.strip()
                else:
                    d = parse_decl(t)
                    if d:
                        for v in d:
                            comps[v.name] = v
                i += 1
            mod.types[m.group(1)] = comps
            mod.type_ext[m.group(1)] = ext.group(1) if ext else None
            mod.type_binds[m.group(1)] = binds
            i += 1
            continue
        mo = re.match(r"interface\s+(operator|assignment)\s*\(\s*([^)]+?)\s*\)$", st)
        if mo:
            specs = []
            i += 1
            while not re.match(r"end\s*interface", sts[i][1]):
                mm = re.match(r"module\s+procedure\s+(.*)$", sts[i][1])
                if mm:
                    specs += [x.strip() for x in mm.group(1).split(",")]
                i += 1
            mod.operators.setdefault(mo.group(2), []).extend(specs)
            i += 1
            continue
        m = re.match(r"(?:abstract\s+)?interface\s*([a-z_]\w*)?$", st)
        if m:
            gname = m.group(1)
            specs = []
            i += 1
            while not re.match(r"end\s*interface", sts[i][1]):
                mm = re.match(r"module\s+procedure\s+(.*)$", sts[i][1])
                if mm:
                    specs += [x.strip() for x in mm.group(1).split(",")]
                i += 1
            if gname:
                mod.generics.setdefault(gname, []).extend(specs)
            i += 1
            continue
        d = parse_decl(st)
        if d:
            for v in d:
                mod.vars[v.name] = v
        i += 1
    return i


def _parse_proc(sts, i, mod, into=None):
    ln, st = sts[i]
    h = _HDR.match(st)
    kind, name, tail = h.group(1), h.group(2), (h.group(3) or "").strip()
    args, result = [], None
    if tail.startswith("("):
        j = _match_paren(tail, 0)
        args = [a.strip() for a in tail[1:j - 1].split(",") if a.strip()]
        rest = tail[j:].strip()
        m = re.match(r"result\s*\(\s*([a-z_]\w*)\s*\)", rest)
        if m:
            result = m.group(1)
    if kind == "function" and result is None:
        result = name
    P = Proc(name, kind, args, result, mod)
    P.line = ln
    P.elemental = bool(re.search(r"\belemental\b", st.split("(")[0]))
    m = re.search(r"\b(real|integer|logical)\b(?=.*\bfunction\b)", st.split("(")[0])
    if kind == "function" and m and result not in P.vars:
        P.vars[result] = Var(result, m.group(1))
    i += 1
    n = len(sts)
    pending_attrs = []
    # specification part
    while i < n:
        ln, st = sts[i]
        if st.startswith("use"):
            u = parse_use(st)
            if u:
                P.uses.append(u)
                i += 1
                continue
        if st.startswith("implicit") or st.startswith("external") or st.startswith("save") or st.startswith("intrinsic"):
            i += 1
            continue
        ma = re.match(r"(optional|pointer|target|allocatable)\s*(?:::)?\s*(.*)$", st)
        if ma and find_assign_simple(st) < 0:
            for nm in _split_top(ma.group(2)):
                nm = nm.strip()
                if nm in P.vars:
                    setattr(P.vars[nm], ma.group(1), True)
                else:
                    pending_attrs.append((nm, ma.group(1)))
            i += 1
            continue
        d = parse_decl(st)
        if d is None:
            break
        for v in d:
            P.vars[v.name] = v
        i += 1
    for nm, attr in pending_attrs:
        if nm in P.vars:
            setattr(P.vars[nm], attr, True)
    for a in args:
        if a in P.vars:
            P.vars[a].dummy = True
    # executable part, up to the matching end
    body, i = _parse_block(sts, i, ("endproc",))
    if body and body[-1][0] == "internal":
        _, _, first, last = body.pop()
        j = first
        while j < last:
            if _HDR.match(sts[j][1]):
                j = _parse_proc(sts, j, mod, into=P.internal)
            else:
                j += 1
        for q in P.internal.values():
            q.host = P
    P.body = body
    (mod.procs if into is None else into)[name] = P
    return i


_END_PROC = re.compile(r"^end\s*(subroutine|function)?\b(\s+[a-z_]\w*)?$")


def _parse_block(sts, i, stop):
    """-> (list of nodes, index just past the terminator).  Nodes: ('stmt', ln, text) | ('if', ln, [(cond, block)...], else) |
    ('do', ln, var, lo, hi, step, block) | ('dowhile', ln, cond, block) | ('doforever', ln, block) | ('select', ln, expr, cases)"""
    out = []
    n = len(sts)
    while i < n:
        ln, st = sts[i]
        if "endproc" in stop and (_END_PROC.match(st) and not re.match(r"end\s*(if|do|select|type|interface|where|module)", st)):
            return out, i + 1
        if st == "contains":
            # internal procedures: note where they lie (the host's _parse_proc parses them) and skip to the end of the host
            depth = 0
            first = i + 1
            i += 1
            while i < n:
                s2 = sts[i][1]
                if _HDR.match(s2):
                    depth += 1
                elif _END_PROC.match(s2) and not re.match(r"end\s*(if|do|select|type|interface|where|module)", s2):
                    if depth == 0:
                        out.append(("internal", ln, first, i))
                        return out, i + 1
                    depth -= 1
                i += 1
            return out, i
        if "endif" in stop and re.match(r"(end\s*if|else\b|else\s*if\b|elseif\b)", st):
            return out, i
        if "enddo" in stop and re.match(r"end\s*do\b", st):
            return out, i + 1
        if "endselect" in stop and (re.match(r"end\s*select", st) or re.match(r"case\b", st)):
            return out, i
        if "endseltype" in stop and (re.match(r"end\s*select", st) or re.match(r"(type|class)\s+is\b", st) or re.match(r"class\s+default$", st)):
            return out, i
        # --- if construct
        if st.startswith("if") and re.match(r"if\s*\(", st):
            j = _match_paren(st, st.index("("))
            cond, rest = st[st.index("(") + 1:j - 1], st[j:].strip()
            if rest == "then":
                arms, els = [], None
                blk, i = _parse_block(sts, i + 1, ("endif",))
                arms.append((cond, blk))
                while True:
                    ln2, s2 = sts[i]
                    m = re.match(r"(?:else\s*if|elseif)\s*\(", s2)
                    if m:
                        j2 = _match_paren(s2, s2.index("("))
                        c2 = s2[s2.index("(") + 1:j2 - 1]
                        blk, i = _parse_block(sts, i + 1, ("endif",))
                        arms.append((c2, blk))
                    elif re.match(r"else$", s2):
                        els, i = _parse_block(sts, i + 1, ("endif",))
                    elif re.match(r"end\s*if", s2):
                        i += 1
                        break
                    else:
                        raise SyntaxError(f"line {ln2}: unexpected {s2!r} inside if construct")
                out.append(("if", ln, arms, els))
                continue
            # one-line if: the action statement may itself be any simple statement
            out.append(("if", ln, [(cond, [("stmt", ln, rest)])], None))
            i += 1
            continue
        # --- do constructs
        m = re.match(r"do\s+([a-z_]\w*)\s*=\s*(.*)$", st)
        if m:
            parts = _split_top(m.group(2))
            blk, i = _parse_block(sts, i + 1, ("enddo",))
            out.append(("do", ln, m.group(1), parts[0], parts[1], parts[2] if len(parts) > 2 else None, blk))
            continue
        m = re.match(r"do\s+while\s*\(", st)
        if m:
            j = _match_paren(st, st.index("("))
            blk, i = _parse_block(sts, i + 1, ("enddo",))
            out.append(("dowhile", ln, st[st.index("(") + 1:j - 1], blk))
            continue
        if st == "do":
            blk, i = _parse_block(sts, i + 1, ("enddo",))
            out.append(("doforever", ln, blk))
            continue
        m = re.match(r"select\s*case\s*\(", st)
        if m:
            j = _match_paren(st, st.index("("))
            sel = st[st.index("(") + 1:j - 1]
            cases = []
            i += 1
            while True:
                ln2, s2 = sts[i]
                if re.match(r"end\s*select", s2):
                    i += 1
                    break
                mm = re.match(r"case\s*(default|\(.*\))$", s2)
                if not mm:
                    raise SyntaxError(f"line {ln2}: unexpected {s2!r} inside select case")
                lab = mm.group(1)
                blk, i = _parse_block(sts, i + 1, ("endselect",))
                cases.append((None if lab == "default" else lab[1:-1], blk))
            out.append(("select", ln, sel, cases))
            continue
        m = re.match(r"select\s*type\s*\(", st)
        if m:
            # select type ([assoc =>] selector): ('seltype', ln, associate name, selector, [(guard 'is' | 'class' | None, type name, block)])
            j = _match_paren(st, st.index("("))
            sel = st[st.index("(") + 1:j - 1]
            assoc, _, expr = sel.partition("=>") if "=>" in sel else (sel, "", sel)
            arms = []
            i += 1
            while True:
                ln2, s2 = sts[i]
                if re.match(r"end\s*select", s2):
                    i += 1
                    break
                mm = re.match(r"(type|class)\s+is\s*\(\s*([a-z_]\w*)\s*\)$", s2) or re.match(r"(class)\s+(default)$", s2)
                if not mm:
                    raise SyntaxError(f"line {ln2}: unexpected {s2!r} inside select type")
                blk, i = _parse_block(sts, i + 1, ("endseltype",))
                if mm.group(2) == "default":
                    arms.append((None, None, blk))
                else:
                    arms.append(("is" if mm.group(1) == "type" else "class", mm.group(2), blk))
            out.append(("seltype", ln, assoc.strip(), expr.strip(), arms))
            continue
        out.append(("stmt", ln, st))
        i += 1
    return out, i

"""Python code generation for oracle/f90run/translate.py's parse trees.  TEST INFRASTRUCTURE ONLY.

Every arithmetic node is emitted fully parenthesised, so Python evaluates it in the order of the Fortran parse tree: explicit
parentheses are kept, equal-precedence operators associate left to right, '**' right to left.  '/' goes through rt.div
(integer truncation, IEEE real division), '**' through rt.powi / rt.power (libgfortran's integer powering; libm pow)."""
import re

from .rt import INTRINSICS, mangle
from .translate import Parser, Var, parse_expr, tokenize, _split_top, _match_paren

_ARITH = {"+", "-", "*", "/", "**"}
_REAL_FN = {"sqrt", "exp", "log", "sin", "cos", "tan", "atan", "atan2", "tanh", "real", "dble", "float", "epsilon", "tiny",
            "asin", "acos", "sinh", "cosh", "log10", "aint"}
_INT_FN = {"int", "nint", "floor", "ceiling", "size", "lbound", "ubound", "len", "len_trim", "count"}
_SAME_FN = {"abs", "max", "min", "sign", "mod", "modulo", "merge", "sum", "maxval", "minval"}


def _num(text):
    t = text
    if "_" in t:
        t = t[:t.index("_")]
    t = t.replace("d", "e")
    if "." in t or "e" in t:
        return repr(float(t)), "r"
    return str(int(t)), "i"


def find_assign(st):
    """index of the top-level assignment '=' (or '=>') in a statement, or -1"""
    depth, q, i, n = 0, None, 0, len(st)
    while i < n:
        c = st[i]
        if q:
            if c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c in "([":
            depth += 1
        elif c in ")]":
            depth -= 1
        elif c == "=" and depth == 0:
            prev = st[i - 1] if i else ""
            nxt = st[i + 1] if i + 1 < n else ""
            if nxt == "=" or prev in "=/<>":
                i += 2 if nxt == "=" else 1
                continue
            return i
        i += 1
    return -1


# library routines that are stubbed (oracle/f90run/stubs.py, stages.py) and define scalar actual arguments: the positions of those
# arguments; the stub returns their values as a tuple
_STUB_OUTS = {"get_time": (1, 2)}   # get_time(Time, seconds, days), FMS time_manager


class ProcGen:
    def __init__(self, prog, mod, proc):
        self.prog, self.mod, self.P = prog, mod, proc
        self.lines = []
        self.globals_assigned = set()
        self.host_assigned = set()   # an internal procedure's assignments to its host's variables (-> nonlocal)
        self.tmp = 0
        self.ret = None
        # scalar locals with an initialisation in their declaration (or the SAVE attribute) keep their value between calls
        self.save_vars = [mangle(n) for n, v in proc.vars.items()
                          if not v.dummy and not v.parameter and v.dims is None and v.base in ("real", "integer", "logical")
                          and v.init is not None and v.init[0] == "val"]

    # ---- symbols
    def var(self, name):
        v = self.P.vars.get(name)
        if v is not None:
            return v
        if self.P.host is not None and name in self.P.host.vars:
            return self.P.host.vars[name]
        return self.mod.vars.get(name)

    def is_local(self, name):
        if name in self.P.vars:
            return True
        if self.P.host is not None and name in self.P.host.vars:   # host association: a variable of the enclosing function
            self.host_assigned.add(mangle(name))
            return True
        return False

    def find_proc(self, name):
        """what a call to name resolves to: an internal procedure of this one (or a sibling, from inside one), else as Program.find_proc"""
        for P in (self.P, self.P.host):
            if P is not None and name in P.internal:
                return P.internal[name]
        return self.prog.find_proc(self.mod, name)

    def newtmp(self, stem="t"):
        self.tmp += 1
        return f"_{stem}{self.tmp}"

    # ---- expression types
    def etype(self, e):
        k = e[0]
        if k == "num":
            return _num(e[1])[1]
        if k == "log":
            return "l"
        if k == "str":
            return "s"
        if k == "paren":
            return self.etype(e[1])
        if k == "un":
            return "l" if e[1] == ".not." else self.etype(e[2])
        if k == "bin":
            op = e[1]
            if op in _ARITH:
                a, b = self.etype(e[2]), self.etype(e[3])
                if op == "**":
                    return a
                if a == "r" or b == "r":
                    return "r"
                if a == "i" and b == "i":
                    return "i"
                return None
            if op == "//":
                return "s"
            return "l"
        if k == "ref":
            parts = e[1]
            name, args = parts[0]
            if len(parts) == 1:
                v = self.var(name)
                if v is not None:
                    if v.base in ("type", "class"):
                        return "o"
                    if v.dims is not None and (args is None or any(a[0] == "slice" for a in args)):
                        return None  # an array value
                    return v.kind
                if args is not None:
                    if name in _REAL_FN:
                        return "r"
                    if name in _INT_FN:
                        return "i"
                    if name in ("present", "associated", "allocated", "any", "all", "is_root_pe"):
                        return "l"
                    if name in _SAME_FN and args:
                        ts = [self.etype(a) for a in args if a[0] != "kw"]
                        if all(t == "r" for t in ts):
                            return "r"
                        if all(t == "i" for t in ts):
                            return "i"
                        return None
                    f = self.find_proc(name)
                    if isinstance(f, list):  # a generic: a type only if every specific agrees
                        ks = {(s.vars[s.result].kind if s.kind == "function" and s.vars.get(s.result) is not None
                               and s.vars[s.result].dims is None else None) for s in f}
                        return ks.pop() if len(ks) == 1 else None
                    if f is not None and f.kind == "function":
                        rv = f.vars.get(f.result)
                        if rv is not None and rv.dims is None:
                            return rv.kind
            return None
        return None

    # ---- expressions
    def ex(self, e):
        k = e[0]
        if k == "num":
            return _num(e[1])[0]
        if k == "log":
            return "True" if e[1] else "False"
        if k == "str":
            s = e[1]
            q = s[0]
            return repr(s[1:-1].replace(q + q, q))
        if k == "paren":
            return "(" + self.ex(e[1]) + ")"
        if k == "un":
            if e[1] == ".not.":
                return "(not " + self.ex(e[2]) + ")"
            return "(-" + self.ex(e[2]) + ")"
        if k == "bin":
            op, a, b = e[1], e[2], e[3]
            A, B = self.ex(a), self.ex(b)
            if op == "/":
                if b[0] == "num" and _num(b[1])[1] == "r" and float(_num(b[1])[0]) != 0.0:
                    return f"({A} / {B})"
                return f"_rt.div({A}, {B})"
            if op == "**":
                tb = self.etype(b)
                if tb == "i":
                    return f"_rt.powi({A}, {B})"
                return f"_rt.power({A}, {B})"
            if op in ("+", "-", "*"):
                return f"({A} {op} {B})"
            if op == "//":
                return f"(str({A}) + str({B}))"
            if op == ".and.":
                return f"({A} and {B})"
            if op == ".or.":
                return f"({A} or {B})"
            if op in (".eqv.",):
                return f"(bool({A}) == bool({B}))"
            if op in (".neqv.",):
                return f"(bool({A}) != bool({B}))"
            if op == "/=":
                return f"({A} != {B})"
            return f"({A} {op} {B})"
        if k == "arr":
            return "[" + ", ".join(self.ex(x) for x in e[1]) + "]"
        if k == "slice":
            return self.slice_(e)
        if k == "kw":
            return f"{mangle(e[1])}={self.ex(e[2])}"
        if k == "ref":
            return self.ref(e[1])
        raise NotImplementedError(k)

    def slice_(self, e):
        f = lambda x: "None" if x is None else self.ex(x)
        return f"({f(e[1])}, {f(e[2])}, {f(e[3])})"

    def index(self, base, args):
        """element or section of the array expression base"""
        if any(a[0] == "slice" for a in args):
            return base + ".sec(" + ", ".join(self.ex(a) for a in args) + ")"
        n = len(args)
        m = f"g{n}" if n <= 3 else "g"
        return f"{base}.{m}(" + ", ".join(self.ex(a) for a in args) + ")"

    def ref(self, parts):
        name, args = parts[0]
        v = self.var(name)
        if v is not None or len(parts) > 1:
            s = mangle(name)
            if args is not None:
                if v is not None and v.dims is None and v.base == "character":
                    pass  # substring: the reference only does this in messages
                else:
                    s = self.index(s, args)
        elif args is not None:
            s = self.call_expr(name, args)
        else:
            s = mangle(name)
        for cname, cargs in parts[1:]:
            s = s + "." + mangle(cname)
            if cargs is not None:
                if self.prog.is_bound(cname):   # a type-bound function
                    s = s + "(" + ", ".join(self.ex(a) for a in cargs) + ")"
                else:
                    s = self.index(s, cargs)
        return s

    def call_expr(self, name, args):
        if name == "present" or name == "allocated":
            return "(" + self.ex(args[0]) + " is not None)"
        if name == "associated":
            if len(args) == 2:
                return "(" + self.ex(args[0]) + " is " + self.ex(args[1]) + ")"
            return "(" + self.ex(args[0]) + " is not None)"
        f = self.find_proc(name)
        if f is None and name in INTRINSICS:
            return f"_i_{name}(" + ", ".join(self.ex(a) for a in args) + ")"
        return mangle(name) + "(" + ", ".join(self.ex(a) for a in args) + ")"

    # ---- statements
    def emit(self, ind, text, ln=None):
        self.lines.append("    " * ind + text + (f"  # L{ln}" if ln is not None else ""))

    def coerce(self, v, e):
        s = self.ex(e)
        if v is None or v.dims is not None:
            return s
        t = self.etype(e)
        if v.base == "real" and t != "r":
            return f"_rt.tofloat({s})"
        if v.base == "integer" and t != "i":
            return f"_rt.toint({s})"
        return s

    def assign(self, ind, lhs, rhs, ln, pointer=False):
        L = parse_expr(lhs)
        if L[0] != "ref":
            raise SyntaxError("bad assignment target " + lhs)
        R = parse_expr(rhs)
        parts = L[1]
        name, args = parts[0]
        if len(parts) == 1:
            v = self.var(name)
            if v is not None and not self.is_local(name):
                self.globals_assigned.add(mangle(name))
            if pointer:
                self.emit(ind, f"{mangle(name)} = {self.ex(R)}", ln)
                return
            if args is None:
                if v is not None and v.dims is not None:
                    self.emit(ind, f"{mangle(name)} = _rt.assign_whole({mangle(name)}, {self.ex(R)})", ln)
                elif v is not None and v.base in ("type", "class") and not v.pointer:
                    # intrinsic (or overloaded, component-wise) assignment of a derived type: the components are copied into the
                    # existing object, so that a dummy argument's actual and every other reference to it see the new value
                    self.emit(ind, f"{mangle(name)} = _rt.assign_derived({mangle(name)}, {self.ex(R)})", ln)
                else:
                    self.emit(ind, f"{mangle(name)} = {self.coerce(v, R)}", ln)
                return
            if v is not None and v.base == "character" and v.dims is None:   # substring assignment: only used for labels
                self.emit(ind, f"{mangle(name)} = {self.ex(R)}", ln)
                return
            self.store(ind, mangle(name), args, R, ln)
            return
        obj = self.ref(parts[:-1])
        cname, cargs = parts[-1]
        if pointer:
            self.emit(ind, f"{obj}.{mangle(cname)} = {self.ex(R)}", ln)
        elif cargs is None:
            self.emit(ind, f"_rt.assign_attr({obj}, {mangle(cname)!r}, {self.ex(R)})", ln)
        else:
            self.store(ind, f"{obj}.{mangle(cname)}", cargs, R, ln)

    def store(self, ind, base, args, R, ln):
        if any(a[0] == "slice" for a in args):
            self.emit(ind, f"{base}.sec(" + ", ".join(self.ex(a) for a in args) + f").assign({self.ex(R)})", ln)
        else:
            n = len(args)
            m = f"s{n}" if n <= 3 else "s_"
            self.emit(ind, f"{base}.{m}(" + ", ".join(self.ex(a) for a in args) + f", {self.ex(R)})", ln)

    def designator_store(self, ind, actual, value, ln):
        """store value into the designator expression 'actual' (copy-out of a scalar dummy argument)"""
        if actual[0] == "kw":
            actual = actual[2]
        if actual[0] != "ref":
            return
        parts = actual[1]
        name, args = parts[0]
        if len(parts) == 1:
            v = self.var(name)
            if v is None:
                return
            if not self.is_local(name):
                self.globals_assigned.add(mangle(name))
            if args is None:
                if v.dims is None:
                    self.emit(ind, f"{mangle(name)} = {value}", ln)
                return
            if any(a[0] == "slice" for a in args):
                return
            n = len(args)
            m = f"s{n}" if n <= 3 else "s_"
            self.emit(ind, f"{mangle(name)}.{m}(" + ", ".join(self.ex(a) for a in args) + f", {value})", ln)
            return
        obj = self.ref(parts[:-1])
        cname, cargs = parts[-1]
        if cargs is None:
            self.emit(ind, f"{obj}.{mangle(cname)} = {value}", ln)
        elif not any(a[0] == "slice" for a in cargs):
            n = len(cargs)
            m = f"s{n}" if n <= 3 else "s_"
            self.emit(ind, f"{obj}.{mangle(cname)}.{m}(" + ", ".join(self.ex(a) for a in cargs) + f", {value})", ln)

    def designator_or_array(self, ind, actual, value, ln):
        """value is stored into a scalar designator; an array (or section) is filled in place by the callee"""
        n0 = len(self.lines)
        self.designator_store(ind, actual, value, ln)
        if len(self.lines) == n0:
            self.emit(ind, value, ln)

    def call(self, ind, text, ln):
        m = re.match(r"call\s+([a-z_][\w%]*)\s*(\(.*\))?\s*$", text, re.S)
        if not m:
            raise SyntaxError("bad call: " + text)
        name = m.group(1)
        args = []
        if m.group(2):
            p = Parser(tokenize(m.group(2)))
            p.expect("op", "(")
            args = p.p_args()
        if "%" in name:
            tgt = ".".join(mangle(x) for x in name.split("%"))
            argtxt = ", ".join(self.ex(a) for a in args)
            impls = self.prog.bound(name.split("%")[-1])
            outs = impls[0].out_scalars() if impls else []
            if not outs:
                self.emit(ind, f"{tgt}({argtxt})", ln)
                return
            f = impls[0]
            r = self.newtmp("r")
            self.emit(ind, f"{r} = {tgt}({argtxt})", ln)
            pos = [a for a in args if a[0] != "kw"]
            kws = {a[1]: a for a in args if a[0] == "kw"}
            for n, d in enumerate(outs):
                k = f.args.index(d) - 1   # the passed-object dummy comes first
                actual = pos[k] if 0 <= k < len(pos) else kws.get(d)
                if actual is not None:
                    self.designator_store(ind, actual, f"{r}[{n}]", ln)
            return
        if name == "random_number" and len(args) == 1:
            self.designator_or_array(ind, args[0], f"_rt.random_number({self.ex(args[0])})", ln)
            return
        if name == "random_seed":
            for a in args:
                if a[0] == "kw" and a[1] == "size":
                    self.designator_store(ind, a[2], "8", ln)
            self.emit(ind, "_rt.random_seed()", ln)
            return
        f = self.find_proc(name)
        argtxt = ", ".join(self.ex(a) for a in args)
        if f is None and name in _STUB_OUTS:   # a stubbed library routine that defines scalar arguments: the stub returns them
            r = self.newtmp("r")
            self.emit(ind, f"{r} = {mangle(name)}({argtxt})", ln)
            pos = [a for a in args if a[0] != "kw"]
            for n, k in enumerate(_STUB_OUTS[name]):
                if k < len(pos):
                    self.designator_store(ind, pos[k], f"{r}[{n}]", ln)
            return
        if isinstance(f, list) and any(s.out_scalars() for s in f):
            # a generic subroutine: which specific runs (and so which scalar arguments it defines) is known at run time; the
            # dispatcher returns them by dummy-argument name and position, stored into the actuals that can take them
            r, v = self.newtmp("r"), self.newtmp("v")
            self.emit(ind, f"{r} = {mangle(name)}({argtxt})", ln)
            npos = 0
            for a in args:
                kw = a[1] if a[0] == "kw" else None
                actual = a[2] if a[0] == "kw" else a
                if actual[0] == "ref" and not any(x is not None and any(y[0] == "slice" for y in x) for _, x in actual[1]):
                    n0 = len(self.lines)
                    self.designator_store(ind + 1, actual, v, ln)
                    body = self.lines[n0:]
                    del self.lines[n0:]
                    if body:
                        self.emit(ind, f"{v} = _rt.genout({r}, {npos if kw is None else None!r}, {mangle(kw) if kw else None!r})")
                        self.emit(ind, f"if {v} is not _rt.NOOUT:")
                        self.lines += body
                if kw is None:
                    npos += 1
            return
        if f is None or isinstance(f, list):
            self.emit(ind, f"{mangle(name)}({argtxt})", ln)
            return
        outs = f.out_scalars()
        if not outs:
            self.emit(ind, f"{mangle(name)}({argtxt})", ln)
            return
        r = self.newtmp("r")
        self.emit(ind, f"{r} = {mangle(name)}({argtxt})", ln)
        pos = [a for a in args if a[0] != "kw"]
        kws = {a[1]: a for a in args if a[0] == "kw"}
        for n, d in enumerate(outs):
            k = f.args.index(d)
            actual = pos[k] if k < len(pos) else kws.get(d)
            if actual is not None:
                self.designator_store(ind, actual, f"{r}[{n}]", ln)

    def formatted_write(self, ind, st, internal, ln):
        """write(unit, fmt) items with an explicit format: the record is built by _rt.fwrite and either assigned to the character
        variable (an internal write) or handed to _rt.unit_write, which keeps it if that unit is being recorded"""
        if not st.startswith("write"):
            raise NotImplementedError(st)
        j = _match_paren(st, st.index("("))
        ctl, items = _split_top(st[st.index("(") + 1:j - 1]), st[j:].strip()
        if len(ctl) != 2 or "=" in ctl[1].replace("==", ""):
            raise NotImplementedError("control list")
        vals = [self.ex(parse_expr(it.strip())) for it in _split_top(items)] if items else []
        if ctl[1].strip() == "*":   # list-directed: messages only (numbers are not edited the way a Fortran processor would)
            if internal:
                raise NotImplementedError("list-directed internal write")
            rec = f"_rt.list_write([{', '.join(vals)}])"
        else:
            rec = f"_rt.fwrite({self.ex(parse_expr(ctl[1].strip()))}, [{', '.join(vals)}])"
        if internal:
            e = parse_expr(ctl[0].strip())
            self.designator_store(ind, e, rec, ln)
        else:
            self.emit(ind, f"_rt.unit_write({self.ex(parse_expr(ctl[0].strip()))}, {rec})", ln)

    def block(self, ind, nodes):
        if not nodes:
            self.emit(ind, "pass")
            return
        for nd in nodes:
            self.node(ind, nd)

    def node(self, ind, nd):
        k = nd[0]
        try:
            if k == "stmt":
                self.stmt(ind, nd[2], nd[1])
            elif k == "if":
                for n, (cond, blk) in enumerate(nd[2]):
                    self.emit(ind, ("if " if n == 0 else "elif ") + self.ex(parse_expr(cond)) + ":", nd[1])
                    self.block(ind + 1, blk)
                if nd[3] is not None:
                    self.emit(ind, "else:")
                    self.block(ind + 1, nd[3])
            elif k == "do":
                _, ln, var, lo, hi, step, blk = nd
                v = mangle(var)
                a, b = self.newtmp("a"), self.newtmp("b")
                self.emit(ind, f"{a} = {self.ex(parse_expr(lo))}; {b} = {self.ex(parse_expr(hi))}", ln)
                if step is None:
                    self.emit(ind, f"{v} = {a}")
                    self.emit(ind, f"for {v} in range({a}, {b} + 1):")
                    self.block(ind + 1, blk)
                    self.emit(ind, "else:")
                    self.emit(ind + 1, f"{v} = {a} if {b} < {a} else {b} + 1")
                else:
                    c = self.newtmp("c")
                    self.emit(ind, f"{c} = {self.ex(parse_expr(step))}")
                    self.emit(ind, f"{v} = {a}")
                    self.emit(ind, f"for {v} in range({a}, ({b} + 1) if {c} > 0 else ({b} - 1), {c}):")
                    self.block(ind + 1, blk)
                    self.emit(ind, "else:")
                    self.emit(ind + 1, f"{v} = {a} + max(0, ({b} - {a} + {c}) // {c}) * {c}")
            elif k == "dowhile":
                self.emit(ind, "while " + self.ex(parse_expr(nd[2])) + ":", nd[1])
                self.block(ind + 1, nd[3])
            elif k == "doforever":
                self.emit(ind, "while True:", nd[1])
                self.block(ind + 1, nd[2])
            elif k == "seltype":
                _, ln, assoc, expr, arms = nd
                sel = self.ex(parse_expr(expr))
                name = mangle(assoc)
                if assoc != expr:
                    self.emit(ind, f"{name} = {sel}", ln)
                first = True
                for guard, tname, blk in arms:
                    # inside the arm the associate name has the guard's type (what its type-bound calls resolve against)
                    had = self.P.vars.get(assoc)
                    if assoc != expr:
                        self.P.vars[assoc] = Var(assoc, "class", tname)
                    if guard is None:
                        self.emit(ind, "if True:" if first else "else:")
                    else:
                        self.emit(ind, ("if " if first else "elif ") + f"_rt.type_is({name}, {tname!r}, {guard == 'class'!r}, globals()):")
                    self.block(ind + 1, blk)
                    if assoc != expr:
                        if had is None:
                            del self.P.vars[assoc]
                        else:
                            self.P.vars[assoc] = had
                    first = False
            elif k == "select":
                s = self.newtmp("sel")
                self.emit(ind, f"{s} = {self.ex(parse_expr(nd[2]))}", nd[1])
                first = True
                default = None
                for lab, blk in nd[3]:
                    if lab is None:
                        default = blk
                        continue
                    conds = []
                    for it in _split_top(lab):
                        if ":" in it and not it.strip().startswith(("'", '"')):
                            lo, hi = it.split(":")
                            c = []
                            if lo.strip():
                                c.append(f"{s} >= {self.ex(parse_expr(lo))}")
                            if hi.strip():
                                c.append(f"{s} <= {self.ex(parse_expr(hi))}")
                            conds.append("(" + " and ".join(c) + ")")
                        else:
                            conds.append(f"{s} == {self.ex(parse_expr(it))}")
                    self.emit(ind, ("if " if first else "elif ") + " or ".join(conds) + ":")
                    self.block(ind + 1, blk)
                    first = False
                if default is not None:
                    if first:
                        self.emit(ind, "if True:")
                    else:
                        self.emit(ind, "else:")
                    self.block(ind + 1, default)
            else:
                raise NotImplementedError(k)
        except (SyntaxError, NotImplementedError, ValueError, IndexError, AttributeError) as err:
            msg = f"{self.mod.name}:{nd[1]}: untranslatable ({type(err).__name__}: {err})"
            self.emit(ind, f"raise NotImplementedError({msg!r})", nd[1])

    def stmt(self, ind, st, ln):
        if st.startswith("call ") or st.startswith("call\t"):
            self.call(ind, st, ln)
            return
        if st == "return":
            for n in self.save_vars:
                self.emit(ind, f"_SAVE[{self.P.name + '.' + n!r}] = {n}", ln)
            self.emit(ind, "return " + self.ret, ln)
            return
        if st == "exit":
            self.emit(ind, "break", ln)
            return
        if st == "cycle":
            self.emit(ind, "continue", ln)
            return
        if st == "continue":
            self.emit(ind, "pass", ln)
            return
        if re.match(r"(write|read|open|close|format|flush|rewind)\b\s*[(*]", st) or re.match(r"print\b\s*['\"(*]", st):
            m = re.match(r"write\s*\(\s*([a-z_]\w*)\s*,", st)
            v = self.var(m.group(1)) if m else None
            internal = v is not None and v.base == "character" and v.dims is None
            try:
                self.formatted_write(ind, st, internal, ln)
            except (SyntaxError, NotImplementedError, ValueError, IndexError, KeyError, AttributeError):
                # list-directed output, implied DO loops, namelists ...: the text is not reproduced
                self.emit(ind, f"{mangle(m.group(1))} = ''" if internal else "pass", ln)
            return
        if re.match(r"(error\s+)?stop\b", st):
            self.emit(ind, f"raise _rt.FortranStop({st!r})", ln)
            return
        m = re.match(r"allocate\s*\((.*)\)$", st, re.S)
        if m:
            mt = re.match(r"\s*([a-z_]\w*)\s*::\s*(.*)$", m.group(1), re.S)
            if mt:  # typed allocation of a polymorphic object
                e = parse_expr(mt.group(2).strip())
                parts = e[1]
                tgt = self.ref(parts[:-1]) + "." + mangle(parts[-1][0]) if len(parts) > 1 else mangle(parts[-1][0])
                self.emit(ind, f"{tgt} = _rt.new_type(globals(), {mt.group(1)!r})", ln)
                return
            for it in _split_top(m.group(1)):
                if re.match(r"(stat|source|mold|errmsg)\s*=", it):
                    mm = re.match(r"source\s*=\s*(.*)$", it)
                    if mm:
                        self.emit(ind, f"_alloc_last.assign({self.ex(parse_expr(mm.group(1)))})", ln)
                    continue
                e = parse_expr(it)
                parts = e[1]
                cname, cargs = parts[-1]
                bounds = []
                for a in cargs or []:
                    if a[0] == "slice":
                        bounds.append(f"({self.ex(a[1])}, {self.ex(a[2])})")
                    else:
                        bounds.append(f"(1, {self.ex(a)})")
                tgt = self.ref(parts[:-1] + [(cname, None)]) if len(parts) > 1 else mangle(cname)
                v = self.var(parts[0][0]) if len(parts) == 1 else None
                kind = v.kind if v is not None else "r"
                if len(parts) == 1 and v is not None and not self.is_local(cname):
                    self.globals_assigned.add(mangle(cname))
                if cargs is None:
                    self.emit(ind, f"{tgt} = _rt.NS()", ln)
                else:
                    self.emit(ind, f"_alloc_last = {tgt} = _rt.alloc({kind!r}, ({', '.join(bounds)},))", ln)
            return
        m = re.match(r"(deallocate|nullify)\s*\((.*)\)$", st, re.S)
        if m:
            for it in _split_top(m.group(2)):
                if re.match(r"stat\s*=", it):
                    continue
                e = parse_expr(it)
                parts = e[1]
                if len(parts) == 1:
                    self.emit(ind, f"{mangle(parts[0][0])} = None", ln)
                else:
                    self.emit(ind, f"{self.ref(parts[:-1])}.{mangle(parts[-1][0])} = None", ln)
            return
        k = find_assign(st)
        if k > 0:
            if st[k:k + 2] == "=>":
                self.assign(ind, st[:k].strip(), st[k + 2:].strip(), ln, pointer=True)
            else:
                self.assign(ind, st[:k].strip(), st[k + 1:].strip(), ln)
            return
        raise NotImplementedError("statement: " + st)

    # ---- declarations
    def bounds(self, dims):
        """-> (list of lower-bound code, list of extent code or None, list of (lo,hi) code or None when deferred)"""
        lbs, exts, full = [], [], []
        for d in dims:
            d = d.strip()
            parts = _split_top(d, ":")
            if d == "*" or d == "..":
                lbs.append("1"); exts.append("None"); full.append(None)
            elif len(parts) == 1:
                hi = self.ex(parse_expr(parts[0]))
                lbs.append("1"); exts.append(hi); full.append(("1", hi))
            else:
                lo = parts[0].strip()
                hi = parts[1].strip()
                lo_c = self.ex(parse_expr(lo)) if lo else "1"
                if hi and hi != "*":
                    hi_c = self.ex(parse_expr(hi))
                    lbs.append(lo_c); exts.append(f"({hi_c}) - ({lo_c}) + 1"); full.append((lo_c, hi_c))
                else:
                    lbs.append(lo_c); exts.append("None"); full.append(None)
        return lbs, exts, full

    def generate(self):
        P = self.P
        outs = P.out_scalars()
        self.fn_outs = []
        if P.kind == "function":
            self.ret = mangle(P.result)
            # A function's value is all the translator hands back: what the function gives to its scalar intent(out) / (inout)
            # arguments does not reach the caller.  That must not happen silently, so the value at every RETURN is compared with
            # the value at entry and a change stops the run (_rt.fn_ret).
            self.fn_outs = [mangle(a) for a in P.args if a in P.vars and P.vars[a].dims is None and P.vars[a].intent in ("out", "inout")
                            and P.vars[a].base in ("real", "integer", "logical", "character")]
            if self.fn_outs:
                self.ret = (f"_rt.fn_ret({mangle(P.result)}, {P.name!r}, [" +
                            ", ".join(f"({a!r}, _e_{a}, {a})" for a in self.fn_outs) + "])")
        else:
            self.ret = "(" + "".join(mangle(a) + ", " for a in outs) + ")" if outs else "None"
        # prologue
        pro = []
        for name, v in P.vars.items():
            n = mangle(name)
            try:
                if v.dummy:
                    if v.dims is not None and (v.pointer or v.allocatable):
                        pass  # a pointer / allocatable dummy keeps the bounds of its actual argument
                    elif v.dims is not None and v.base != "character":
                        lbs, exts, _ = self.bounds(v.dims)
                        pro.append(f"{n} = _rt.rebase({n}, ({', '.join(lbs)},), ({', '.join(exts)},), {P.name + ':' + name!r})")
                    elif v.dims is None and v.base == "real":
                        pro.append(f"if {n} is not None: {n} = float({n})")
                    continue
                if v.parameter:
                    if v.dims is not None:
                        _, _, full = self.bounds(v.dims)
                        pro.append(f"{n} = _rt.alloc({v.kind!r}, ({', '.join('(%s, %s)' % b for b in full)},))")
                        pro.append(f"{n}.assign({self.ex(parse_expr(v.init[1]))})")
                    else:
                        pro.append(f"{n} = {self.coerce(v, parse_expr(v.init[1]))}")
                    continue
                if name == P.result and P.kind == "function" and v.dims is None and v.base not in ("type", "class"):
                    continue
                if v.dims is not None and v.base not in ("character", "type", "class"):
                    _, _, full = self.bounds(v.dims)
                    if any(b is None for b in full) or v.pointer or v.allocatable:
                        pro.append(f"{n} = None")
                    else:
                        pro.append(f"{n} = _rt.alloc({v.kind!r}, ({', '.join('(%s, %s)' % b for b in full)},))")
                        if v.init is not None and v.init[0] == "val":
                            pro.append(f"{n}.assign({self.ex(parse_expr(v.init[1]))})")
                elif v.base in ("type", "class"):
                    if v.dims is not None:
                        _, _, full = self.bounds(v.dims)
                        if any(b is None for b in full) or v.pointer or v.allocatable:
                            pro.append(f"{n} = None")
                        else:
                            pro.append(f"{n} = _rt.alloc_types(globals(), {v.tname!r}, ({', '.join('(%s, %s)' % b for b in full)},))")
                    else:
                        pro.append(f"{n} = None" if v.pointer else f"{n} = _rt.new_type(globals(), {v.tname!r})")
                elif v.init is not None and v.init[0] == "val":
                    if n in self.save_vars:
                        pro.append(f"{n} = _SAVE[{P.name + '.' + n!r}] if {P.name + '.' + n!r} in _SAVE else "
                                   f"{self.coerce(v, parse_expr(v.init[1]))}")
                    else:
                        pro.append(f"{n} = {self.coerce(v, parse_expr(v.init[1]))}")
                elif v.pointer:
                    pro.append(f"{n} = None")
                elif v.dims is not None and v.base == "character" and not v.allocatable:
                    _, _, full = self.bounds(v.dims)
                    if all(b is not None for b in full):   # an explicit-shape array of strings, blank until assigned
                        pro.append(f"{n} = _rt.alloc('o', ({', '.join('(%s, %s)' % b for b in full)},))")
                        pro.append(f"{n}.v[:] = [''] * len({n}.v)")
                elif v.dims is None and v.base in ("integer", "real", "logical", "character"):
                    # undefined until assigned; given a value so that it can be passed to an intent(out) dummy argument
                    # (a real starts as NaN, so a use before definition still shows in the results)
                    pro.append(f"{n} = " + {"integer": "0", "real": "_rt.NAN", "logical": "False", "character": "''"}[v.base])
            except (SyntaxError, NotImplementedError, ValueError, IndexError) as err:
                pro.append(f"{n} = None  # declaration not translated: {err}")
        self.lines = []
        self.block(1, P.body)
        body = self.lines
        head = [f"def {mangle(P.name)}(" + ", ".join(mangle(a) + "=None" for a in P.args) + f"):  # {self.mod.name}:{P.line}"]
        head += [f"    _e_{a} = {a}" for a in self.fn_outs]
        if self.globals_assigned:
            head.append("    global " + ", ".join(sorted(self.globals_assigned)))
        if self.host_assigned:
            head.append("    nonlocal " + ", ".join(sorted(self.host_assigned)))
        for q in P.internal.values():   # internal procedures: nested functions, defined once the host's locals exist
            pro += ProcGen(self.prog, self.mod, q).generate().split("\n")
        tail = [f"    _SAVE[{P.name + '.' + n!r}] = {n}" for n in self.save_vars]
        if P.name in self.prog.expose:   # for tests that compare a procedure's local variables with the oracle's
            tail.append(f"    _SAVE[{P.name + '.__locals__'!r}] = dict(locals())")
        out = head + ["    " + p for p in pro] + body + tail + ["    return " + self.ret, ""]
        if P.elemental:
            n = mangle(P.name)
            out.append(f"{n} = _rt.elemental({n}, {[mangle(a) for a in P.args]!r}, {[mangle(a) for a in outs]!r}, {P.kind == 'function'!r})")
            out.append("")
        return "\n".join(out)


class Program:
    def __init__(self):
        self.modules = {}
        self.expose = set()   # procedures whose local variables are kept (in _SAVE['<name>.__locals__']) when they return

    def bound(self, name):
        """the procedures a type-bound name may resolve to (one per type that binds it); [] if it is not a binding name"""
        out = []
        for m in self.modules.values():
            for t, binds in m.type_binds.items():
                if name in binds and binds[name] in m.procs:
                    out.append(m.procs[binds[name]])
        return out

    def is_bound(self, name):
        for m in self.modules.values():
            for binds in m.type_binds.values():
                if name in binds:
                    return True
        return False

    def add(self, mods):
        for m in mods:
            self.modules[m.name] = m

    def find_proc(self, mod, name):
        """the Proc a call to name from mod resolves to; a list of Procs for a generic; None if it is not a translated procedure"""
        if name in mod.procs:
            return mod.procs[name]
        if name in mod.generics:
            return [mod.procs[s] for s in mod.generics[name] if s in mod.procs]
        for uname, only in mod.uses:
            um = self.modules.get(uname)
            if um is None:
                continue
            remote = name
            if only is not None:
                hit = [r for l, r in only if l == name]
                if not hit:
                    continue
                remote = hit[0]
            if remote in um.procs:
                return um.procs[remote]
            if remote in um.generics:
                return [um.procs[s] for s in um.generics[remote] if s in um.procs]
        return None

    def gen_module(self, mod):
        out = [f"# generated from {mod.path} by oracle/f90run -- do not commit", "import oracle.f90run.rt as _rt", "_SAVE = {}"]
        for k in INTRINSICS:
            out.append(f"_i_{k} = _rt.INTRINSICS[{k!r}]")
        out.append("")
        # module variables
        for name, v in mod.vars.items():
            n = mangle(name)
            if v.init is not None and v.init[0] == "val" and v.dims is None:
                pg = ProcGen(self, mod, _EmptyProc(mod))
                try:
                    code = pg.coerce(v, parse_expr(v.init[1]))
                    out.append(f"try:\n    {n} = {code}\nexcept (NameError, AttributeError, TypeError):\n    {n} = None")
                except (SyntaxError, NotImplementedError):
                    out.append(f"{n} = None")
            elif v.init is not None and v.init[0] == "val" and v.base in ("real", "integer", "logical"):
                pg = ProcGen(self, mod, _EmptyProc(mod))
                try:
                    _, _, full = pg.bounds(v.dims)
                    code = pg.ex(parse_expr(v.init[1]))
                    out.append(f"try:\n    {n} = _rt.alloc({v.kind!r}, ({', '.join('(%s, %s)' % b for b in full)},))\n"
                               f"    {n}.assign({code})\nexcept (NameError, AttributeError, TypeError):\n    {n} = None")
                except (SyntaxError, NotImplementedError, TypeError):
                    out.append(f"{n} = None")
            else:
                out.append(f"{n} = None")
        out.append("")
        for P in mod.procs.values():
            out.append(ProcGen(self, mod, P).generate())
        for op, specs in mod.operators.items():
            for sname in specs:
                P = mod.procs.get(sname)
                if P is None or not P.args:
                    continue
                v0 = P.vars.get(P.args[0])
                if v0 is not None and v0.tname:
                    out.append(f"_rt.register_op({op!r}, {mangle(v0.tname)!r}, {mangle(sname)})")
        for tname, comps in mod.types.items():
            out.append(f"def _new_{mangle(tname)}():")
            ext = mod.type_ext.get(tname)
            out.append(f"    o = _rt.new_type(globals(), {ext!r})" if ext else "    o = _rt.NS()")
            if ext:
                out.append(f"    o._bases = tuple(o.__dict__.get('_bases') or ()) + ({ext!r},)")
            out.append(f"    o._type = {tname!r}")
            for b, impl in mod.type_binds.get(tname, {}).items():
                out.append(f"    o.{mangle(b)} = _rt.bind(globals(), {mangle(impl)!r}, o)")
            pg = ProcGen(self, mod, _EmptyProc(mod))
            for cname, v in comps.items():
                if v.init is not None and v.init[0] == "val" and v.dims is None:
                    try:
                        out.append(f"    o.{mangle(cname)} = {pg.coerce(v, parse_expr(v.init[1]))}")
                    except (SyntaxError, NotImplementedError):
                        pass
                elif v.base in ("type", "class") and v.dims is None and not v.pointer and not v.allocatable:
                    out.append(f"    o.{mangle(cname)} = _rt.new_type(globals(), {v.tname!r})")
                elif v.dims is not None and not v.pointer and not v.allocatable and v.base in ("real", "integer", "logical"):
                    try:   # an explicit-shape array component with constant bounds
                        _, _, full = pg.bounds(v.dims)
                        if all(b is not None for b in full):
                            out.append(f"    o.{mangle(cname)} = _rt.alloc({v.kind!r}, ({', '.join('(%s, %s)' % b for b in full)},))")
                            if v.init is not None and v.init[0] == "val":
                                out.append(f"    o.{mangle(cname)}.assign({pg.ex(parse_expr(v.init[1]))})")
                    except (SyntaxError, NotImplementedError):
                        pass
            out.append("    return o")
            out.append("")
        for g, specs in mod.generics.items():
            specs = [s for s in specs if s in mod.procs]
            if not specs:
                continue
            sig = []
            for s in specs:
                P = mod.procs[s]
                sig.append("(" + mangle(s) + ", [" + ", ".join(
                    "(%r, %r, %r)" % (mangle(a), (len(P.vars[a].dims) if (a in P.vars and P.vars[a].dims is not None) else 0),
                                      bool(a in P.vars and P.vars[a].optional)) for a in P.args) + "], " +
                           repr([mangle(a) for a in P.out_scalars()]) + ")")
            out.append(f"{mangle(g)} = _rt.generic({g!r}, [" + ", ".join(sig) + "])")
        return "\n".join(out) + "\n"


class _EmptyProc:
    def __init__(self, mod):
        self.vars, self.args, self.kind, self.result, self.name, self.line, self.body = {}, [], "subroutine", None, "<module>", 0, []
        self.internal, self.host = {}, None

    def out_scalars(self):
        return []

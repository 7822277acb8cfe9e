"""Run-time support for the Fortran-subset translator (oracle/f90run/translate.py).  TEST INFRASTRUCTURE ONLY.

The translator turns the reference's own .F90 sources -- read where they lie under /root/reference, never copied into this repo --
into Python that evaluates every expression in IEEE binary64 in exactly the order the Fortran text prescribes (parentheses kept,
equal-precedence operators left to right, no FMA contraction: what gfortran -O0 -ffp-contract=off would do).  This module holds
what the generated code calls: Fortran arrays with arbitrary lower bounds and section views, derived-type instances, the
intrinsics and the integer/real division and power rules."""
import math

NAN = float("nan")
INF = float("inf")
LENIENT_READS = False


class FortranStop(Exception):
    pass


OPS = {}


def register_op(op, tname, f):
    OPS[(op, tname)] = f


class NS:
    """A derived-type instance: components are attributes (lower case); an unset pointer / absent component reads as None."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, mangle(k), v)

    def __getattr__(self, name):  # only called when the attribute is missing
        if name.startswith("__"):
            raise AttributeError(name)
        return None

    def __repr__(self):
        return "NS(" + ", ".join(sorted(self.__dict__)) + ")"

    # operators overloaded for a derived type (interface operator(+) ...): dispatched on the type of the left operand
    def __add__(self, o):
        return OPS[("+", self.__dict__.get("_type"))](self, o)

    def __sub__(self, o):
        return OPS[("-", self.__dict__.get("_type"))](self, o)

    def __mul__(self, o):
        return OPS[("*", self.__dict__.get("_type"))](self, o)


_PYKW = {"is", "in", "as", "or", "and", "not", "if", "else", "for", "while", "def", "del", "from", "global", "lambda", "pass",
         "class", "return", "yield", "import", "with", "try", "raise", "none", "true", "false", "assert", "break", "continue",
         "elif", "except", "finally", "nonlocal", "async", "await", "print", "exec", "type", "len", "range", "float", "int",
         "tuple", "list", "math", "isinstance", "bool", "str", "abs", "min", "max", "sum"}


def mangle(name):
    n = name.lower()
    return n + "_" if n in _PYKW else n


def _prod(xs):
    p = 1
    for x in xs:
        p *= x
    return p


class FArray:
    """A Fortran array (or a section of one): shared flat storage, strides, lower bounds.  Element (i,j,k) lives at
    v[b + i*s0 + j*s1 + k*s2]."""
    __slots__ = ("v", "b", "s", "lb", "shape", "kind")

    def __init__(self, v, b, s, lb, shape, kind):
        self.v, self.b, self.s, self.lb, self.shape, self.kind = v, b, tuple(s), tuple(lb), tuple(shape), kind

    # ---- construction
    @staticmethod
    def alloc(kind, bounds, fill=None):
        lb = tuple(b[0] for b in bounds)
        shape = tuple(max(0, b[1] - b[0] + 1) for b in bounds)
        n = _prod(shape)
        if fill is None:
            fill = NAN if kind == "r" else (0 if kind == "i" else (False if kind == "l" else None))
        if kind == "o":
            v = [NS() for _ in range(n)]
        else:
            v = [fill] * n
        s, acc = [], 1
        for e in shape:
            s.append(acc)
            acc *= e
        return FArray(v, -sum(l * st for l, st in zip(lb, s)), s, lb, shape, kind)

    @staticmethod
    def from_numpy(a, lb=None, kind=None):
        """a is C-ordered with the Fortran dimensions reversed ([k,j,i] for a Fortran (i,j,k) array)."""
        import numpy as np
        a = np.ascontiguousarray(a)
        shape = tuple(reversed(a.shape))
        if kind is None:
            kind = "r" if a.dtype.kind == "f" else ("l" if a.dtype.kind == "b" else "i")
        v = a.ravel().tolist()
        if lb is None:
            lb = (1,) * len(shape)
        s, acc = [], 1
        for e in shape:
            s.append(acc)
            acc *= e
        return FArray(v, -sum(l * st for l, st in zip(lb, s)), s, lb, shape, kind)

    def to_numpy(self):
        import numpy as np
        dt = {"r": np.float64, "i": np.int64, "l": np.bool_}[self.kind]
        out = np.array(self.tolist(), dtype=dt)
        return out.reshape(tuple(reversed(self.shape)))

    def tolist(self):
        """elements in Fortran array-element order"""
        if len(self.shape) == 1:
            b, s0, l0 = self.b, self.s[0], self.lb[0]
            return [self.v[b + (l0 + i) * s0] for i in range(self.shape[0])]
        out = []
        self._walk(len(self.shape) - 1, self.b, out)
        return out

    def _walk(self, d, off, out):
        s, l, n = self.s[d], self.lb[d], self.shape[d]
        if d == 0:
            v = self.v
            out.extend(v[off + (l + i) * s] for i in range(n))
        else:
            for i in range(n):
                self._walk(d - 1, off + (l + i) * s, out)

    def _offsets(self):
        out = []

        def walk(d, off):
            s, l, n = self.s[d], self.lb[d], self.shape[d]
            if d == 0:
                out.extend(off + (l + i) * s for i in range(n))
            else:
                for i in range(n):
                    walk(d - 1, off + (l + i) * s)
        if self.shape:
            walk(len(self.shape) - 1, self.b)
        return out

    def rebase(self, lbs):
        """the same elements seen through a dummy argument declared with lower bounds lbs (explicit-shape / assumed-shape rules)"""
        lbs = tuple(lbs)
        if len(lbs) != len(self.shape):
            if _prod(self.shape) == 0:
                return self
            raise TypeError(f"rank mismatch in argument association: actual rank {len(self.shape)}, dummy rank {len(lbs)}")
        if lbs == self.lb:
            return self
        b = self.b + sum((l - nl) * s for l, nl, s in zip(self.lb, lbs, self.s))
        return FArray(self.v, b, self.s, lbs, self.shape, self.kind)

    def check_extent(self, d, n, what=""):
        if n is not None and self.shape[d] != n:
            raise TypeError(f"{what}: dimension {d + 1} has extent {self.shape[d]}, the dummy argument declares {n}")

    # ---- element access (bounds-checked: a silent wrap-around would hide an indexing error in the caller's adapter)
    def _chk(self, d, i):
        l = self.lb[d]
        if i < l or i >= l + self.shape[d]:
            raise IndexError(f"index {i} outside {l}:{l + self.shape[d] - 1} in dimension {d + 1}")

    def _oob(self, d, i):
        """an out-of-bounds READ: an error, unless LENIENT_READS is set (the reference's PPM_hybgen unit test reads u(n+1) and
        discards it), in which case the undefined value is NaN"""
        if LENIENT_READS:
            return NAN
        self._chk(d, i)

    def g1(self, i):
        l = self.lb[0]
        if i < l or i >= l + self.shape[0]: return self._oob(0, i)
        return self.v[self.b + i * self.s[0]]

    def g2(self, i, j):
        lb, sh = self.lb, self.shape
        if i < lb[0] or i >= lb[0] + sh[0]: self._chk(0, i)
        if j < lb[1] or j >= lb[1] + sh[1]: self._chk(1, j)
        s = self.s
        return self.v[self.b + i * s[0] + j * s[1]]

    def g3(self, i, j, k):
        lb, sh = self.lb, self.shape
        if i < lb[0] or i >= lb[0] + sh[0]: self._chk(0, i)
        if j < lb[1] or j >= lb[1] + sh[1]: self._chk(1, j)
        if k < lb[2] or k >= lb[2] + sh[2]: self._chk(2, k)
        s = self.s
        return self.v[self.b + i * s[0] + j * s[1] + k * s[2]]

    def g(self, *idx):
        off = self.b
        if len(idx) != len(self.shape):
            raise IndexError(f"rank {len(self.shape)} array referenced with {len(idx)} subscripts")
        for d, i in enumerate(idx):
            self._chk(d, i)
            off += i * self.s[d]
        return self.v[off]

    def _cv(self, x):
        k = self.kind
        if k == "r":
            return float(x)
        if k == "i":
            return int(x)
        return x

    def s1(self, i, x):
        self._chk(0, i)
        self.v[self.b + i * self.s[0]] = self._cv(x)

    def s2(self, i, j, x):
        self._chk(0, i); self._chk(1, j)
        s = self.s
        self.v[self.b + i * s[0] + j * s[1]] = self._cv(x)

    def s3(self, i, j, k, x):
        self._chk(0, i); self._chk(1, j); self._chk(2, k)
        s = self.s
        self.v[self.b + i * s[0] + j * s[1] + k * s[2]] = self._cv(x)

    def s_(self, *a):
        idx, x = a[:-1], a[-1]
        off = self.b
        if len(idx) != len(self.shape):
            raise IndexError(f"rank {len(self.shape)} array assigned with {len(idx)} subscripts")
        for d, i in enumerate(idx):
            self._chk(d, i)
            off += i * self.s[d]
        self.v[off] = self._cv(x)

    # ---- sections
    def sec(self, *items):
        """items: an int (that dimension collapses) or a (lo, hi, step) triplet with None for an omitted bound"""
        if len(items) != len(self.shape):
            raise IndexError(f"rank {len(self.shape)} array sectioned with {len(items)} subscripts")
        b, s, lb, shape = self.b, [], [], []
        for d, it in enumerate(items):
            if isinstance(it, tuple):
                lo, hi, st = it
                if lo is None: lo = self.lb[d]
                if hi is None: hi = self.lb[d] + self.shape[d] - 1
                if st is None: st = 1
                n = max(0, (hi - lo + st) // st)
                if n > 0:
                    self._chk(d, lo); self._chk(d, lo + (n - 1) * st)
                # new dimension has lower bound 1: element m (1-based) is old index lo + (m-1)*st
                b += (lo - st) * self.s[d]
                s.append(self.s[d] * st); lb.append(1); shape.append(n)
            else:
                self._chk(d, it)
                b += it * self.s[d]
        return FArray(self.v, b, s, lb, shape, self.kind)

    def assign(self, x):
        """array assignment: self = x (a scalar, or a conformable array; the right side is fully evaluated first)"""
        offs = self._offsets()
        v, cv = self.v, self._cv
        if isinstance(x, FArray):
            vals = x.tolist()
            if len(vals) != len(offs):
                raise ValueError(f"array assignment of non-conformable shapes {x.shape} -> {self.shape}")
            for o, val in zip(offs, vals):
                v[o] = cv(val)
        elif isinstance(x, (list, tuple)):
            if len(x) != len(offs):
                raise ValueError("array constructor of the wrong size")
            for o, val in zip(offs, x):
                v[o] = cv(val)
        else:
            if self.kind == "o":
                raise TypeError("scalar assignment to an array of derived type")
            val = cv(x)
            for o in offs:
                v[o] = val

    # ---- elementwise expressions
    def _new(self, vals, kind=None):
        out = FArray.alloc(kind or self.kind, [(1, n) for n in self.shape])
        out.v[:] = vals
        return out

    def _bin(self, o, f, kind=None):
        a = self.tolist()
        if isinstance(o, FArray):
            b = o.tolist()
            if len(a) != len(b):
                raise ValueError("non-conformable array operands")
            return self._new([f(x, y) for x, y in zip(a, b)], kind)
        return self._new([f(x, o) for x in a], kind)

    def __add__(self, o): return self._bin(o, lambda x, y: x + y)
    def __radd__(self, o): return self._bin(o, lambda x, y: y + x)
    def __sub__(self, o): return self._bin(o, lambda x, y: x - y)
    def __rsub__(self, o): return self._bin(o, lambda x, y: y - x)
    def __mul__(self, o): return self._bin(o, lambda x, y: x * y)
    def __rmul__(self, o): return self._bin(o, lambda x, y: y * x)
    def __truediv__(self, o): return div(self, o)
    def __rtruediv__(self, o): return div(o, self)
    def __neg__(self): return self._new([-x for x in self.tolist()])
    def __pos__(self): return self
    def __lt__(self, o): return self._bin(o, lambda x, y: x < y, "l")
    def __le__(self, o): return self._bin(o, lambda x, y: x <= y, "l")
    def __gt__(self, o): return self._bin(o, lambda x, y: x > y, "l")
    def __ge__(self, o): return self._bin(o, lambda x, y: x >= y, "l")


def alloc(kind, bounds, fill=None):
    return FArray.alloc(kind, bounds, fill)


def rebase(x, lbs, extents=None, what=""):
    if x is None:
        return None
    if isinstance(x, (list, tuple)):   # an array constructor as the actual argument
        flat = []
        for e in x:
            flat.extend(e.tolist() if type(e) is FArray else [e])
        kind = "r" if any(type(e) is float for e in flat) else ("l" if flat and type(flat[0]) is bool else "i")
        a = FArray.alloc(kind, [(1, len(flat))])
        a.v[:] = [float(e) for e in flat] if kind == "r" else flat
        x = a
    if not isinstance(x, FArray):
        raise TypeError(f"{what}: array dummy argument associated with a {type(x).__name__}")
    if len(lbs) != len(x.shape) and extents is not None and _prod(x.shape) != 0:
        # sequence association with a change of rank: the dummy's explicit shape is laid over the actual's elements in order
        first = x.b + sum(l * st for l, st in zip(x.lb, x.s))
        contiguous, acc = True, x.s[0]
        for n, st in zip(x.shape, x.s):
            contiguous = contiguous and st == acc
            acc *= n
        if not contiguous:
            raise TypeError(f"{what}: rank-changing association with a non-contiguous actual argument")
        total, shape = _prod(x.shape), []
        for n in extents:
            shape.append(n if n is not None else max(1, total // max(1, _prod(shape))))
        st, acc = [], x.s[0]
        for n in shape:
            st.append(acc)
            acc *= n
        return FArray(x.v, first - sum(l * q for l, q in zip(lbs, st)), st, lbs, shape, x.kind)
    y = x.rebase(lbs)
    if extents is not None and len(y.shape) == len(extents):
        for d, n in enumerate(extents):
            # an explicit-shape dummy may be associated with a longer actual (sequence association) in its last dimension
            if n is not None and _prod(y.shape) != 0 and (y.shape[d] < n or (y.shape[d] != n and d != len(extents) - 1)):
                raise TypeError(f"{what}: dimension {d + 1} of the actual argument has extent {y.shape[d]}, the dummy declares {n}")
    return y


# ---- arithmetic rules ----------------------------------------------------------------------------------------------------
def div(a, b):
    """Fortran '/': integer division truncates toward zero; real division is IEEE (x/0 -> inf or NaN, no exception)."""
    ta, tb = type(a), type(b)
    if ta is int and tb is int:
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    if ta is FArray:
        return a._bin(b, div)
    if tb is FArray:
        return b._bin(a, lambda y, x: div(x, y))
    try:
        return a / b
    except ZeroDivisionError:
        a = float(a)
        if a != a or a == 0.0:
            return NAN
        neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, float(b)) < 0)
        return -INF if neg else INF


def powi(x, n):
    """x**n for an integer n the way libgfortran's pow_r8_i4 does it (binary powering from the low bit up)."""
    if type(x) is int:
        if n >= 0:
            return x ** n
        return 1 if x == 1 else ((-1) ** n if x == -1 else 0)
    if type(x) is FArray:
        return x._new([powi(e, n) for e in x.tolist()])
    u = -n if n < 0 else n
    if n < 0:
        x = div(1.0, x)
    p = 1.0
    if u == 0:
        return p
    while True:
        if u & 1:
            p = p * x
        u >>= 1
        if u:
            x = x * x
        else:
            break
    return p


def power(x, y):
    if type(y) is int:
        return powi(x, y)
    if type(x) is FArray:
        return x._new([power(e, y) for e in x.tolist()])
    x = float(x)
    try:
        return math.pow(x, y)  # libm pow, the routine the compiled reference calls
    except (OverflowError, ValueError):
        if x == 0.0 and y < 0:
            return INF
        if x < 0:
            return NAN
        return INF


# ---- intrinsics ----------------------------------------------------------------------------------------------------------
def _elemental(f):
    def g(*a):
        for x in a:
            if type(x) is FArray:
                lists = [y.tolist() if type(y) is FArray else None for y in a]
                n = len(next(l for l in lists if l is not None))
                vals = [f(*[(l[m] if l is not None else y) for l, y in zip(lists, a)]) for m in range(n)]
                return x._new(vals)
        return f(*a)
    return g


def _max(*a):
    # gfortran: MAX(a,b) = (a > b || isnan(b)) ? a : b evaluated left to right; equal operands keep the LATER one only for NaN
    m = a[0]
    for x in a[1:]:
        if x > m or m != m:
            m = x
    return m


def _min(*a):
    m = a[0]
    for x in a[1:]:
        if x < m or m != m:
            m = x
    return m


def _sign(a, b):
    if type(a) is int and type(b) is int:
        return abs(a) if b >= 0 else -abs(a)
    return math.copysign(a, b)


def _mod(a, b):
    if type(a) is int and type(b) is int:
        return a - div(a, b) * b
    return math.fmod(a, b)


def _modulo(a, b):
    if type(a) is int and type(b) is int:
        return a % b
    r = math.fmod(a, b)
    if r != 0 and (r < 0) != (b < 0):
        r += b
    return r


def _sqrt(x):
    try:
        return math.sqrt(x)
    except ValueError:
        return NAN


def _exp(x):
    try:
        return math.exp(x)
    except OverflowError:
        return INF


def _log(x):
    try:
        return math.log(x)
    except ValueError:
        return -INF if x == 0 else NAN


def _nint(x):
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def _real(x, kind=None):
    return float(x)


def _int(x, kind=None):
    return int(x)  # truncates toward zero


def _size(a, dim=None):
    if dim is None:
        return _prod(a.shape)
    return a.shape[dim - 1]


def _lbound(a, dim=None):
    return a.lb[dim - 1] if dim is not None else list(a.lb)


def _ubound(a, dim=None):
    if dim is not None:
        return a.lb[dim - 1] + a.shape[dim - 1] - 1
    return [l + n - 1 for l, n in zip(a.lb, a.shape)]


def _sum(a, dim=None, mask=None):
    t = 0.0 if a.kind == "r" else 0
    if mask is not None:
        for x, m in zip(a.tolist(), mask.tolist()):
            if m:
                t = t + x
        return t
    for x in a.tolist():  # array-element order, as a scalar loop would add them
        t = t + x
    return t


def _maxval(a, dim=None, mask=None):
    l = a.tolist()
    return _max(*l) if l else -1.7976931348623157e308


def _minval(a, dim=None, mask=None):
    l = a.tolist()
    return _min(*l) if l else 1.7976931348623157e308


def _any(a):
    return any(a.tolist()) if type(a) is FArray else bool(a)


def _all(a):
    return all(a.tolist()) if type(a) is FArray else bool(a)


def _merge(t, f, m):
    return t if m else f


def _trim(s):
    return s.rstrip()


def _transfer(x, mold):
    """TRANSFER between a real(8) and an integer of the same size (the only use on the path: the bit pattern of a real)"""
    import struct
    if type(x) is float and type(mold) is int:
        return struct.unpack("<q", struct.pack("<d", x))[0]
    if type(x) is int and type(mold) is float:
        return struct.unpack("<d", struct.pack("<q", x))[0]
    raise NotImplementedError("transfer: only real(8) <-> integer(8) scalars")


INTRINSICS = {
    "abs": _elemental(abs), "max": _elemental(_max), "min": _elemental(_min), "sqrt": _elemental(_sqrt),
    "sign": _elemental(_sign), "mod": _mod, "modulo": _modulo, "exp": _elemental(_exp), "log": _elemental(_log),
    "sin": math.sin, "cos": math.cos, "tan": math.tan, "atan": math.atan, "atan2": math.atan2, "tanh": math.tanh,
    "asin": math.asin, "acos": math.acos, "sinh": math.sinh, "cosh": math.cosh, "log10": math.log10,
    "real": _real, "dble": _real, "float": _real, "int": _int, "nint": _nint, "floor": lambda x, kind=None: int(math.floor(x)),
    "ceiling": lambda x, kind=None: int(math.ceil(x)), "aint": lambda x: float(int(x)),
    "size": _size, "lbound": _lbound, "ubound": _ubound, "sum": _sum, "maxval": _maxval, "minval": _minval,
    "any": _any, "all": _all, "merge": _merge, "trim": _trim, "adjustl": lambda s: s.lstrip(), "len_trim": lambda s: len(s.rstrip()),
    "len": len, "epsilon": lambda x: 2.220446049250313e-16, "huge": lambda x: (1.7976931348623157e308 if type(x) is float else 2147483647),
    "tiny": lambda x: 2.2250738585072014e-308, "isnan": lambda x: x != x, "ieee_is_nan": lambda x: x != x,
    "count": lambda a: sum(1 for x in a.tolist() if x), "null": lambda: None,
    "index": lambda s_, sub, back=False: (s_.rfind(sub) if back else s_.find(sub)) + 1, "scan": lambda s_, set_: next((n + 1 for n, c in enumerate(s_) if c in set_), 0),
    "char": chr, "ichar": ord, "achar": chr, "iachar": ord, "repeat": lambda s_, n: s_ * n, "new_line": lambda c: "\n",
    "btest": lambda i, pos: bool((i >> pos) & 1), "ibset": lambda i, pos: i | (1 << pos), "ibclr": lambda i, pos: i & ~(1 << pos),
    "iand": lambda i, j: i & j, "ior": lambda i, j: i | j, "ishft": lambda i, s: (i << s) if s >= 0 else (i >> -s),
    "kind": lambda x: 8 if type(x) is float else 4, "popcnt": lambda i: bin(i & 0xFFFFFFFFFFFFFFFF).count("1"), "transfer": _transfer,
}


def assign_attr(obj, name, x):
    """obj%name = x where name may be an array component (array assignment) or a scalar / pointer component"""
    cur = obj.__dict__.get(name)
    if type(cur) is FArray and not (type(x) is FArray and x is cur):
        cur.assign(x)
    else:
        setattr(obj, name, x)


def tofloat(x):
    return x if type(x) is FArray else float(x)


def toint(x):
    return x if type(x) is FArray else int(x)


def assign_whole(cur, x):
    """name = x for a declared array name: array assignment into the existing storage (an unallocated allocatable takes x's shape)"""
    if cur is None:
        if type(x) is FArray:
            out = FArray.alloc(x.kind, [(l, l + n - 1) for l, n in zip(x.lb, x.shape)])
            out.assign(x)
            return out
        raise ValueError("assignment to an unallocated array")
    cur.assign(x)
    return cur


def fn_ret(value, fname, outs):
    """the value of a function; outs = [(dummy name, its value at entry, its value now)] for its scalar intent(out) / (inout) arguments,
    which the translator cannot hand back: a changed one would be lost, so it stops the run instead"""
    for name, before, now in outs:
        if before is not None and now is not before and not (now == before):
            raise NotImplementedError(f"{fname}: the function changed its scalar argument {name} ({before!r} -> {now!r}); "
                                      "values a function gives to its arguments are not returned by the translator")
    return value


class GenOut:
    """what a generic subroutine hands back: the scalar arguments the chosen specific defined, by dummy name, and its dummy order"""
    __slots__ = ("names", "values")

    def __init__(self, names, values):
        self.names, self.values = names, values


NOOUT = object()


def genout(res, pos, kw):
    """the value the specific procedure gave to its dummy argument at position pos (or named kw), NOOUT if it defines none there"""
    if type(res) is not GenOut:
        return NOOUT
    name = kw if kw is not None else (res.names[pos] if pos is not None and pos < len(res.names) else None)
    return res.values.get(name, NOOUT)


def generic(name, specifics):
    """a generic interface: the first specific whose dummy arguments fit the actual ones (count, keywords, array rank) is called"""
    def rank(x):
        return len(x.shape) if type(x) is FArray else 0

    def call(*a, **kw):
        for f, sig, *more in specifics:
            outs = more[0] if more else []
            if len(a) > len(sig):
                continue
            names = [s[0] for s in sig]
            if any(k not in names for k in kw):
                continue
            given = dict(zip(names, a))
            given.update(kw)
            ok = True
            for n, r, opt in sig:
                if n not in given or given[n] is None:
                    if not opt and n not in given:
                        ok = False
                        break
                    continue
                x = given[n]
                if (r > 0) != (type(x) is FArray) or (r > 0 and rank(x) != r):
                    ok = False
                    break
            if ok:
                res = f(*a, **kw)
                if outs and isinstance(res, tuple) and len(res) == len(outs):   # a subroutine that defines scalar arguments
                    return GenOut(names, dict(zip(outs, res)))
                return res
        raise TypeError(f"no specific procedure of generic {name} matches the call")
    return call


def new_type(ns, tname):
    """an instance of derived type tname with its default component initialisations, if the type's module was translated"""
    f = ns.get("_new_" + mangle(tname))
    return f() if f is not None else NS()


def type_is(obj, tname, or_extension, ns):
    """the guards of SELECT TYPE: TYPE IS (tname) -- the dynamic type is exactly tname -- or CLASS IS (tname) -- tname or an extension"""
    t = getattr(obj, "_type", None)
    if t is None:
        return False
    if mangle(t) == mangle(tname):
        return True
    if not or_extension:
        return False
    chain = getattr(obj, "__dict__", {}).get("_bases") or ()
    return mangle(tname) in [mangle(b) for b in chain]


def alloc_types(ns, tname, bounds):
    a = FArray.alloc("o", bounds)
    a.v[:] = [new_type(ns, tname) for _ in a.v]
    return a


def bind(ns, fname, obj):
    """a type-bound procedure: the module procedure fname called with obj as its passed-object dummy argument"""
    def call(*a, **k):
        return ns[fname](obj, *a, **k)
    return call


def elemental(f, argnames, outs, is_function):
    """an elemental procedure: applied element by element when any actual argument is an array"""
    def call(*a, **k):
        args = dict(zip(argnames, a))
        args.update(k)
        arrs = {n: x for n, x in args.items() if type(x) is FArray}
        if not arrs:
            return f(*a, **k)
        first = next(iter(arrs.values()))
        lists = {n: x.tolist() for n, x in arrs.items()}
        offs = {n: arrs[n]._offsets() for n in outs if n in arrs}
        res = []
        for m in range(_prod(first.shape)):
            r = f(**{n: (lists[n][m] if n in lists else x) for n, x in args.items()})
            if is_function:
                res.append(r)
            else:
                for q, o in enumerate(outs):
                    if o in offs:
                        arrs[o].v[offs[o][m]] = arrs[o]._cv(r[q])
        if is_function:
            return first._new(res, "r" if res and type(res[0]) is float else None)
        return tuple(args.get(o) for o in outs)
    return call


_RNG = [12345]


def random_seed():
    _RNG[0] = 12345


def _rand():
    _RNG[0] = (_RNG[0] * 6364136223846793005 + 1442695040888963407) % (1 << 64)
    return (_RNG[0] >> 11) / float(1 << 53)


def random_number(x):
    """RANDOM_NUMBER: uniform numbers in [0,1) (the reference only uses them for self-consistency checks in its unit tests)"""
    if type(x) is FArray:
        for o in x._offsets():
            x.v[o] = _rand()
        return x
    return _rand()


def _copy_value(x):
    if type(x) is FArray:
        out = FArray.alloc(x.kind, [(l, l + n - 1) for l, n in zip(x.lb, x.shape)])
        out.v[:] = [(_copy_value(e) if isinstance(e, NS) else e) for e in x.tolist()]
        return out
    if isinstance(x, NS):
        out = NS()
        for k, v in x.__dict__.items():
            out.__dict__[k] = v if callable(v) else _copy_value(v)
        return out
    return x


def assign_derived(cur, x):
    """a = b for derived types: value semantics.  Array and nested-type components are copied; pointer components (which the
    translator cannot tell from allocatable ones at run time) are copied too, which is right wherever the copy is not modified
    through the pointer afterwards."""
    if x is None or not isinstance(x, NS):
        return x
    new = _copy_value(x)
    if cur is None or not isinstance(cur, NS):
        return new
    # type-bound procedures are bound to the object they were created for: keep the target's own bindings
    keep = {k: v for k, v in cur.__dict__.items() if callable(v)}
    cur.__dict__.clear()
    cur.__dict__.update(new.__dict__)
    cur.__dict__.update(keep)
    return cur


# ---------------------------------------------------------------------------------------------------------- formatted output
# Fortran format-directed output for the edit descriptors the reference's diagnostics use (A, I, F, E, ES, L, Z, X, /, :, character
# string literals, repeat counts and groups), following the Fortran 2008 standard's section 10 the way gfortran prints: plus signs
# suppressed, a leading "0" before the decimal point of F / E only where the width allows, asterisks on overflow, two exponent
# digits with "E" (three without it).  This is a restatement of the standard -- libgfortran is not executed -- so what a test of
# formatted text pins is the reference's format string and output list, not its I/O library.
UNITS = {}   # unit number -> list of records, for the units a test wants to read back (see unit_write)


def unit_write(unit, text):
    if type(unit) is int and unit in UNITS:
        UNITS[unit] += text.split("\n")


def list_write(items):
    """WRITE(unit, *) items, for messages: strings as they are, numbers by repr (NOT a Fortran processor's list-directed editing)"""
    out = []
    for x in items:
        out += [repr(v) if type(v) is float else str(v) for v in (x.tolist() if type(x) is FArray else [x])]
    return " " + " ".join(out)


def _fmt_tokens(f):
    """the items of a format specification (without its outer parentheses), as a nested list"""
    out, i, n = [], 0, len(f)
    while i < n:
        c = f[i]
        if c in " ,":
            i += 1
        elif c in "'\"":
            j, lit = i + 1, []
            while True:
                if f[j] == c:
                    if j + 1 < n and f[j + 1] == c:
                        lit.append(c); j += 2; continue
                    break
                lit.append(f[j]); j += 1
            out.append(("lit", "".join(lit)))
            i = j + 1
        elif c == "/":
            out.append(("nl",)); i += 1
        elif c == ":":
            out.append(("colon",)); i += 1
        else:
            j = i
            while j < n and f[j].isdigit():
                j += 1
            rep = int(f[i:j]) if j > i else None
            if j < n and f[j] == "(":
                depth, k = 0, j
                while True:
                    if f[k] in "'\"":   # skip a literal inside the group
                        q = f[k]; k += 1
                        while f[k] != q:
                            k += 1
                    elif f[k] == "(":
                        depth += 1
                    elif f[k] == ")":
                        depth -= 1
                        if depth == 0:
                            break
                    k += 1
                out.append(("group", rep or 1, _fmt_tokens(f[j + 1:k])))
                i = k + 1
                continue
            k = j
            while k < n and f[k].isalpha():
                k += 1
            name = f[j:k].lower()
            m = k
            while m < n and (f[m].isdigit() or f[m] == "."):
                m += 1
            spec = f[k:m]
            e = None
            if m < n and f[m] in "eE" and name in ("e", "es", "en", "g"):
                m2 = m + 1
                while m2 < n and f[m2].isdigit():
                    m2 += 1
                e, m = int(f[m + 1:m2]), m2
            if name == "x":
                out.append(("x", rep or 1))
            elif name in ("a", "i", "f", "e", "es", "l", "z"):
                w, _, d = spec.partition(".")
                out.append(("edit", rep or 1, name, int(w) if w else None, int(d) if d else None, e))
            else:
                raise NotImplementedError(f"format item {f[i:m]!r}")
            i = m
    return out


def _fit(s, w):
    if w is None or w == 0:
        return s
    return "*" * w if len(s) > w else s.rjust(w)


def _edit_real(name, x, w, d, e):
    x = float(x)
    if x != x or x in (INF, -INF):
        s = "NaN" if x != x else ("-Infinity" if x < 0 else "Infinity")
        if w and len(s) > w:
            s = s.replace("inity", "")
        return _fit(s, w)
    neg = math.copysign(1.0, x) < 0
    if name == "f":
        s = "%.*f" % (d, abs(x))
        if w and s.startswith("0.") and len(s) + neg > w and d > 0:
            s = s[1:]   # the optional leading zero goes first
        return _fit(("-" if neg else "") + s, w)
    if name == "es":
        m, _, ex = ("%.*E" % (d, abs(x))).partition("E")
        ex = int(ex)
    else:   # E: 0.ddddE+ee, d significant digits
        m, _, ex = ("%.*E" % (max(d - 1, 0), abs(x))).partition("E")
        ex = int(ex) + (0 if abs(x) == 0.0 else 1)
        m = "0." + m.replace(".", "")
    ne = 2 if e is None else e
    if abs(ex) >= 10 ** ne and e is None:
        tail = "%+04d" % ex   # three exponent digits: the letter E is dropped
    else:
        tail = "E%+0*d" % (ne + 1, ex)
    s = m + tail
    if name == "e" and w and len(s) + neg > w:
        s = s[1:]
    return _fit(("-" if neg else "") + s, w)


def _edit(name, x, w, d, e):
    if name == "a":
        s = str(x)
        return s if w is None else (s[:w] if len(s) > w else s.rjust(w))
    if name == "l":
        return _fit("T" if x else "F", w or 2)
    if name == "i":
        x = int(x)
        s = str(abs(x))
        if d is not None:
            s = "" if (d == 0 and x == 0) else s.rjust(d, "0")
        return _fit(("-" if x < 0 else "") + s, w)
    if name == "z":
        import struct
        v = struct.unpack("<Q", struct.pack("<d", x))[0] if type(x) is float else (int(x) & 0xFFFFFFFF)
        s = "%X" % v
        if d is not None:
            s = s.rjust(d, "0")
        return _fit(s, w)
    return _edit_real(name, x, w, d, e)


def fwrite(fmt, items):
    """the record(s) WRITE(unit, fmt) items produces, records joined by newlines; a format with an edit descriptor that is not
    implemented here gives a marker text instead (such writes are messages, never results)"""
    try:
        return _fwrite(fmt, items)
    except (NotImplementedError, ValueError, TypeError, IndexError) as err:
        return f"<formatted output not reproduced: {err}>"


def _fwrite(fmt, items):
    fmt = fmt.strip()
    if not (fmt.startswith("(") and fmt.endswith(")")):
        raise NotImplementedError("format " + fmt)
    toks = _fmt_tokens(fmt[1:-1])
    vals = []
    for x in items:
        vals += x.tolist() if type(x) is FArray else [x]
    out, pos = [], [0]

    class _Done(Exception):
        pass

    def run(ts):
        for t in ts:
            if t[0] == "lit":
                out.append(t[1])
            elif t[0] == "x":
                out.append(" " * t[1])
            elif t[0] == "nl":
                out.append("\n")
            elif t[0] == "colon":
                if pos[0] >= len(vals):
                    raise _Done
            elif t[0] == "group":
                for _ in range(t[1]):
                    run(t[2])
            else:
                for _ in range(t[1]):
                    if pos[0] >= len(vals):
                        raise _Done
                    out.append(_edit(t[2], vals[pos[0]], t[3], t[4], t[5]))
                    pos[0] += 1

    try:
        run(toks)
        guard = 0
        while pos[0] < len(vals):   # format reversion: a new record, from the last top-level group (or the start)
            guard += 1
            if guard > 10000 or not any(t[0] in ("edit", "group") for t in toks):
                raise NotImplementedError("format reversion without a data edit descriptor")
            out.append("\n")
            last = [n for n, t in enumerate(toks) if t[0] == "group"]
            run(toks[last[-1]:] if last else toks)
    except _Done:
        pass
    return "".join(out)

"""Struct layouts across the boundary: include/mom6cu.h == fortran/mom6cu_interface.F90 (bind(C) types) == the ctypes mirrors of
mom6_b200/_lib.py == sizeof in the built library (mom6cu_sizeof), member for member and in order.  Round 1's interface file had
mom6cu_dyn_split_rk2_cs 40 bytes short (six members missing after `unsupported`): a Fortran caller would have handed the library shifted
pointers.  Nothing compiled the file, so nothing noticed; this test parses it instead (tests/abi_parse.py)."""
import ctypes as C
import os

import pytest

import abi_parse as A
from mom6_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "mom6cu.h")
F90 = os.path.join(ROOT, "fortran", "mom6cu_interface.F90")

SIZE = {"int": (4, 4), "double": (8, 8), "ptr": (8, 8), "i64": (8, 8), "size_t": (8, 8)}


def _layout(structs, name):
    """(size, alignment) of a struct under the x86-64 / aarch64 LP64 C ABI (the rule ISO_C_BINDING's bind(C) types follow)."""
    off, amax = 0, 1
    for kind, _, cnt in structs[name]:
        n = int(cnt) if cnt.isdigit() else {"MOM6CU_HOR_VISC_NARRAYS": 30}.get(cnt, 1)
        if kind.startswith("struct:"):
            sz, al = _layout(structs, kind.split(":")[1])
        else:
            sz, al = SIZE[kind]
        off = (off + al - 1) // al * al + sz * n
        amax = max(amax, al)
    return (off + amax - 1) // amax * amax, amax


def test_fortran_types_mirror_the_header_member_for_member():
    c, f = A.c_structs(HDR), A.fortran_bindc_types(F90)
    assert len(c) >= 37
    assert not sorted(set(c) - set(f)), f"structs without a bind(C) type: {sorted(set(c) - set(f))}"
    assert not sorted(set(f) - set(c)), f"bind(C) types without a C struct: {sorted(set(f) - set(c))}"
    for n in c:
        cm = [(k, m.lower(), cnt) for k, m, cnt in c[n]]
        fm = [(k, m.lower(), cnt) for k, m, cnt in f[n]]
        assert cm == fm, (n, next((i, x, y) for i, (x, y) in enumerate(zip(cm + [None] * 99, fm + [None] * 99)) if x != y))


def test_every_c_entry_has_a_fortran_interface():
    cf, fi = A.c_functions(HDR), A.fortran_bindc_interfaces(F90)
    assert len(cf) >= 66
    assert not [x for x in cf if x not in fi]
    assert not [x for x in fi if x not in cf]
    assert all(k == v for k, v in fi.items())          # the Fortran name is the C name: no silent renames


def test_library_sizeof_matches_header_fortran_and_ctypes():
    lib = _lib.load()
    lib.mom6cu_sizeof.restype = C.c_longlong
    lib.mom6cu_sizeof.argtypes = [C.c_char_p]
    c, f = A.c_structs(HDR), A.fortran_bindc_types(F90)
    assert lib.mom6cu_sizeof(b"no_such_struct") == -1
    for n in c:
        got = lib.mom6cu_sizeof(n.encode())
        assert got > 0, n
        assert got == _layout(c, n)[0], (n, got, _layout(c, n))
        assert got == _layout(f, n)[0], (n, "fortran", got, _layout(f, n))
    # the ctypes mirrors (every Structure of _lib.py documents the struct it mirrors in its docstring's first word)
    mirrors = {}
    for k, v in vars(_lib).items():
        if isinstance(v, type) and issubclass(v, C.Structure) and v is not C.Structure and (v.__doc__ or "").startswith("mom6cu_"):
            mirrors[(v.__doc__ or "").split(":")[0].split()[0].rstrip(".,")] = v
    assert len(mirrors) >= 30, sorted(mirrors)
    for n, cls in mirrors.items():
        if n in c:
            assert C.sizeof(cls) == lib.mom6cu_sizeof(n.encode()), (n, C.sizeof(cls), lib.mom6cu_sizeof(n.encode()))
            assert [m.lower() for _, m, _ in c[n]] == [x[0].lower() for x in cls._fields_], n

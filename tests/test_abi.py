"""The C-ABI shared library (mom6_b200/libmom6cu.so) loads without a GPU and exports every entry point include/mom6cu.h
declares; the Fortran interface module binds only names the library exports; without a device the compute entries fail
loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from mom6_b200 import _lib
from mom6_b200.api import make_domain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mom6cu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mom6cu_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    names = _declared()
    assert len(names) > 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/mom6cu.h but not exported: {missing}"


def test_fortran_interface_binds_exported_names():
    lib = _lib.load()
    src = open(os.path.join(ROOT, "fortran", "mom6cu_interface.F90")).read()
    bound = sorted(set(re.findall(r'bind\(C,\s*name="(mom6cu_[a-z0-9_]+)"\)', src, flags=re.I)))
    assert len(bound) > 30
    missing = [n for n in bound if not hasattr(lib, n)]
    assert not missing, f"bound in fortran/mom6cu_interface.F90 but not exported: {missing}"
    undeclared = [n for n in bound if n not in _declared()]
    assert not undeclared


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    lib = _lib.load()
    assert lib.mom6cu_build_arch() == 100
    h = C.c_void_p()
    dom = make_domain(8, 8, nk=2)
    assert lib.mom6cu_create(C.byref(h), C.byref(dom), 0) == 1  # MOM6CU_ERR_NO_DEVICE
    assert not h.value


def _build_hpp_host(tmp_path):
    import subprocess
    exe = str(tmp_path / "hpp_host")
    libdir = os.path.join(ROOT, "mom6_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "harness", "hpp_host.cpp"),
                           "-L", libdir, "-lmom6cu", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_host_mirror_builds_and_fails_loudly_without_a_device(tmp_path):
    """include/mom6cu.hpp (the compiled-language host side: the reference's procedure names over the C ABI, MOM_error(FATAL) as an
    exception) compiles warning-free against the header and links against the library; without a GPU it refuses to run."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    r = subprocess.run([_build_hpp_host(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stdout, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_cpp_host_mirror_runs_the_reference_mu_vectors(tmp_path):
    import subprocess
    r = subprocess.run([_build_hpp_host(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "FATAL as expected" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_every_declared_entry_has_python_argtypes_and_a_cpp_or_fortran_binding():
    """The ctypes mirror declares the argument types of every entry of include/mom6cu.h (a missing declaration would let ctypes pass 32-bit
    ints where the C side expects pointers), and every compute entry is reachable from a compiled-language binding."""
    lib = _lib.load()
    names = _declared()
    untyped = [n for n in names if getattr(lib, n).argtypes is None]
    assert not untyped, f"no argtypes in mom6_b200/_lib.py for: {untyped}"
    hpp = open(os.path.join(ROOT, "include", "mom6cu.hpp")).read()
    f90 = open(os.path.join(ROOT, "fortran", "mom6cu_interface.F90")).read()
    unbound = [n for n in names if n not in hpp and n not in f90]
    assert not unbound, f"neither include/mom6cu.hpp nor fortran/mom6cu_interface.F90 binds: {unbound}"

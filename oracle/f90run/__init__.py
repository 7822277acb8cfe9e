"""oracle.f90run -- run the reference's own Fortran sources through a small Fortran-subset translator.  TEST INFRASTRUCTURE ONLY.

    ref = load(["src/core/MOM_continuity_PPM.F90", ...])       # paths relative to the reference tree
    ref["mom_continuity_ppm"]["continuity_ppm"](u, v, hin, h, uh, vh, dt, G, GV, US, CS, OBC, pbv, ...)

The sources are read where they lie under REFERENCE_ROOT (/root/reference, absent on the GPU box: callers skip); the generated
Python lives in memory only.  F90RUN_KEEP=1 also writes it to oracle/_ref/f90py/ (git-ignored, never committed: it is derived from
the reference's text) so that tracebacks show the generated lines while debugging."""
import hashlib
import os

from . import rt, stubs
from .codegen import Program
from .rt import FArray, NS
from .translate import Cpp, parse_file

REFERENCE_ROOT = os.environ.get("MOM6_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
CACHE = os.path.join(os.path.dirname(_HERE), "_ref", "f90py")
MEMORY_H_DIR = "config_src/memory/dynamic_symmetric"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "core"))


def _selfhash():
    h = hashlib.sha256()
    for f in ("translate.py", "codegen.py", "rt.py"):
        with open(os.path.join(_HERE, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def load(paths, extra_stubs=None, verbose=False, expose=()):
    """Translate and link the given reference source files.  -> {module name: namespace dict}.  expose: names of procedures whose
    local variables are kept at their final RETURN, as ns["_SAVE"]["<name>.__locals__"] (for results the reference only prints)."""
    prog = Program()
    prog.expose = set(expose)
    srcs = []
    for p in paths:
        full = os.path.join(REFERENCE_ROOT, p)
        cpp = Cpp([os.path.join(REFERENCE_ROOT, MEMORY_H_DIR), os.path.join(REFERENCE_ROOT, "src/framework")])
        mods = parse_file(full, cpp)
        prog.add(mods)
        srcs.append((full, mods))
    keep = os.environ.get("F90RUN_KEEP", "0") == "1"
    if keep:
        os.makedirs(CACHE, exist_ok=True)
    spaces = {}
    for full, mods in srcs:
        for m in mods:
            code = prog.gen_module(m)
            out = "<f90run:%s>" % m.name
            if keep:
                out = os.path.join(CACHE, m.name + ".py")
                with open(out, "w") as f:
                    f.write(code)
            ns = {"__name__": "f90ref." + m.name}
            exec(compile(code, out, "exec"), ns)
            spaces[m.name] = ns
    # link: every name a module imports with "use" is copied from the translated module or from the stubs
    stub_ns = dict(stubs.NAMES)
    if extra_stubs:
        stub_ns.update(extra_stubs)
    for full, mods in srcs * 4:   # repeated: modules re-export what they import (MOM_continuity; testing via Recon1d_type)
        for m in mods:
            ns = spaces[m.name]
            uses = list(m.uses)
            for P in m.procs.values():
                uses += P.uses
            for uname, only in uses:
                src = spaces.get(uname)
                if src is not None:
                    um = prog.modules[uname]
                    if only is None:
                        names = [(n, n) for n in list(um.procs) + list(um.vars) + list(um.generics)]
                    else:
                        names = list(only)
                    names += [("_new_" + t, "_new_" + t) for t in um.types]
                    names += [("_new_" + l, "_new_" + r) for l, r in names if ("_new_" + rt.mangle(r)) in src]   # re-exported types
                    for local, remote in names:
                        key = rt.mangle(remote)
                        if key in src:
                            ns.setdefault(rt.mangle(local), src[key])
                        elif rt.mangle(remote) in stub_ns:
                            ns.setdefault(rt.mangle(local), stub_ns[rt.mangle(remote)])
                else:
                    if only is None:
                        continue
                    for local, remote in only:
                        if rt.mangle(remote) in stub_ns:
                            ns.setdefault(rt.mangle(local), stub_ns[rt.mangle(remote)])
            for k, v in stub_ns.items():  # unqualified "use" of an untranslated module: fall back to the stubs by name
                ns.setdefault(k, v)
    return spaces


def new(spaces, module, tname, **members):
    """an instance of the reference's derived type `tname` (defaults from its declaration), then members set from keywords"""
    o = rt.new_type(spaces[module], tname)
    for k, v in members.items():
        setattr(o, rt.mangle(k), v)
    return o

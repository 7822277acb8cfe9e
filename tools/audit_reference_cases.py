#!/usr/bin/env python
"""Which options of the reference-run cases (tests/refcases.py) are exercised?

For every case and every option it sets, the case is re-run through the oracle with that one option left at its default; an option whose
removal changes no output bit is reported.  Such an option pins nothing in that case (it may be a default spelled out, a member that
needs a companion option to act -- PV_ADV_SCHEME under a Coriolis scheme that does not use it -- or a limit that does not bind on the
seeded inputs).  The second and third sweeps of tests/refcases.py were written from this report; what it still prints is listed in
oracle/README.md.  CPU only, about two minutes:

    python tools/audit_reference_cases.py [case-name-prefix ...]
"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refcases  # noqa: E402
from oracle import pyoracle as oracle  # noqa: E402

# geometry and bookkeeping arguments, not options of the path
SKIP = {"land_blocks", "cyclic_x", "cyclic_y", "first_direction", "nsteps", "diags", "ntr", "uhbt_noise", "dt"}


def run(name, kw):
    c = refcases.CASES[name]
    refcases.CASES["audit/tmp"] = dict(stage=c["stage"], shape=c["shape"], outputs=c["outputs"], kw=kw)
    try:
        return refcases.run_oracle(oracle, "audit/tmp", refcases.build("audit/tmp"))
    except Exception as e:   # the combination without the option is refused
        return str(e)[:80]
    finally:
        del refcases.CASES["audit/tmp"]


def main():
    only = sys.argv[1:]
    for name in sorted(refcases.CASES):
        c = refcases.CASES[name]
        if name in refcases.SLOW or c["stage"] in ("diag", "bt_helpers") or (only and not any(name.startswith(p) for p in only)):
            continue
        base = run(name, copy.deepcopy(c["kw"]))
        for k, v in c["kw"].items():
            if k in SKIP:
                continue
            for k2 in (list(v) if isinstance(v, dict) else [None]):
                kw = copy.deepcopy(c["kw"])
                if k2 is None:
                    del kw[k]
                else:
                    del kw[k][k2]
                o = run(name, kw)
                if not isinstance(o, str) and sorted(o) == sorted(base) and all(np.array_equal(o[x], base[x]) for x in o):
                    print(f"{name}: no effect: {k if k2 is None else k + '.' + k2} = {v if k2 is None else v[k2]}", flush=True)


if __name__ == "__main__":
    main()

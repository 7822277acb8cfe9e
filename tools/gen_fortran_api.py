"""Writes tests/golden/fortran_public_api.json: the public lists and dummy-argument lists of the reference modules the Fortran shims shadow
(fortran/install_shims.py SHIMS), parsed from a MOM6 checkout.  usage: python tools/gen_fortran_api.py [/root/reference]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "fortran"))
import abi_parse as A  # noqa: E402
import install_shims  # noqa: E402

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = {}
for rel, spec in install_shims.SHIMS.items():
    mod, pub, procs = A.fortran_public_api(os.path.join(ref, rel))
    out[rel] = dict(module=mod, public=sorted(p.lower() for p in pub),
                    procs={p.lower(): [a.lower() for a in procs[p]] for p, _ in spec["hooks"]})
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "fortran_public_api.json"), "w"), indent=1, sort_keys=True)
print({k: len(v["public"]) for k, v in out.items()})

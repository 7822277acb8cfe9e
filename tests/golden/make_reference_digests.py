"""Writes tests/golden/reference_f90_digests.json: for every case of tests/refcases.py, the sha256 of the outputs of the
REFERENCE'S OWN routine (its Fortran source executed by oracle/f90run), over the computational domain, -0.0 folded onto +0.0.

    python tests/golden/make_reference_digests.py          (needs /root/reference; about two minutes)

The digests travel where the reference tree does not (the GPU box): tests/test_reference_golden.py checks the oracle (CPU) and
the CUDA path (GPU) against them."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refcases  # noqa: E402


def main():
    out = {}
    for name in sorted(refcases.CASES):
        ref = refcases.run_reference(name, refcases.build(name))
        out[name] = dict(digest=refcases.digest(ref), outputs=sorted(ref))
        print(name, out[name]["digest"][:16], len(ref), "arrays")
    with open(os.path.join(HERE, "reference_f90_digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

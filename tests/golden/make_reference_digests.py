"""Writes tests/golden/reference_f90_digests.json: for every case of tests/refcases.py, the sha256 of the outputs of the
REFERENCE'S OWN routine (its Fortran source executed by oracle/f90run), over the computational domain, -0.0 folded onto +0.0.

    python tests/golden/make_reference_digests.py          (needs /root/reference; every case, the 75-layer ones take minutes)
    python tests/golden/make_reference_digests.py diag/    (only the cases whose names start so; the rest of the file is kept)

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
    path = os.path.join(HERE, "reference_f90_digests.json")
    only = sys.argv[1:]   # prefixes of case names: re-run those and keep the other entries of the existing file
    if only:
        out = {k: v for k, v in json.load(open(path)).items() if k in refcases.CASES}
    for name in sorted(refcases.CASES):
        if only and not any(name.startswith(p) for p in only):
            continue
        ref = refcases.run_reference(name, refcases.build(name))
        out[name] = dict(digest=refcases.digest(ref), outputs=sorted(ref))
        print(name, out[name]["digest"][:16], len(ref), "arrays")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

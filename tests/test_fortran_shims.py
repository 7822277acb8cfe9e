"""The Fortran side of the drop-in boundary, checked without a Fortran compiler (none exists in the image).

fortran/install_shims.py builds shadow copies of the reference's own modules (every public name, type and procedure kept; a hook at the
top of each replaced procedure; the binding procedures of fortran/bodies/*.inc appended).  This test runs it against the reference and
verifies what a compiler would:
  * the public list of every shadow module == the reference's, plus the documented additions (round 1's hand-written MOM_continuity_PPM
    exported 3 of the 17 names MOM_continuity.F90:10-16 imports);
  * the dummy-argument list of every hooked procedure is untouched, the hook passes dummies that exist, in the binding's own order;
  * every  x%member  the bindings touch exists in the type of x -- resolved through the dummies' declared types against the derived types
    of the reference tree (MOM_dyn_split_RK2_CS, barotropic_CS, BT_cont_type, vertvisc_type, mech_forcing, ...) and of
    fortran/mom6cu_interface.F90 (the bind(C) mirrors of include/mom6cu.h, themselves checked by tests/test_abi_layout.py);
  * every mom6cu_* procedure a binding calls has an interface.
tests/golden/fortran_public_api.json holds the reference's public lists and dummy lists (tools/gen_fortran_api.py) so that the first two
checks also run where the reference checkout is absent."""
import glob
import json
import os
import re
import sys

import pytest

import abi_parse as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "fortran"))
import install_shims  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "fortran_public_api.json")
have_ref = os.path.isdir(os.path.join(REF, "src", "core"))


def _api_of(path):
    mod, pub, procs = A.fortran_public_api(path)
    return dict(module=mod, public=sorted(p.lower() for p in pub), procs={k.lower(): [a.lower() for a in v] for k, v in procs.items()})


def test_golden_api_is_current():
    gold = json.load(open(GOLD))
    assert set(gold) == set(install_shims.SHIMS)
    if not have_ref:
        pytest.skip("reference checkout absent: the committed fixture is used as it is")
    for rel in install_shims.SHIMS:
        api = _api_of(os.path.join(REF, rel))
        assert gold[rel]["public"] == api["public"], rel
        for proc, _ in install_shims.SHIMS[rel]["hooks"]:
            assert gold[rel]["procs"][proc.lower()] == api["procs"][proc.lower()], (rel, proc)


def test_hook_arguments_match_the_reference_dummies_and_the_bindings():
    gold = json.load(open(GOLD))
    for rel, spec in install_shims.SHIMS.items():
        body = open(os.path.join(ROOT, "fortran", "bodies", os.path.basename(rel).replace(".F90", ".inc"))).read()
        procs = {p[0].lower(): p for p in A.fortran_procedures(body)}
        for proc, args in spec["hooks"]:
            passed = [a.strip().lower() for a in args.split(",")]
            ref_dummies = gold[rel]["procs"][proc.lower()]
            assert [a for a in ref_dummies if a in passed] == passed, (proc, "hook passes names that are not dummies, or out of order")
            bind = procs[(proc + "_mom6cu").lower()]
            assert [a.lower() for a in bind[1]] == passed, (proc, "binding's dummy list differs from what the hook passes")
        for name in spec["public"]:
            assert name.lower() in procs, name


@pytest.mark.skipif(not have_ref, reason="needs the reference checkout")
def test_shadow_modules_keep_every_public_name_and_dummy_list(tmp_path):
    written = install_shims.install(REF, str(tmp_path))
    assert len(written) == len(install_shims.SHIMS) + 1
    for rel, spec in install_shims.SHIMS.items():
        ref, sh = _api_of(os.path.join(REF, rel)), _api_of(os.path.join(str(tmp_path), os.path.basename(rel)))
        assert ref["module"] == sh["module"]
        assert sorted(set(ref["public"]) | {p.lower() for p in spec["public"]}) == sh["public"], rel
        for k, v in ref["procs"].items():
            assert sh["procs"].get(k) == v, (rel, k)               # every procedure of the reference is still there, dummies untouched
        text = open(os.path.join(str(tmp_path), os.path.basename(rel))).read()
        for proc, _ in spec["hooks"]:
            # the hook sits inside the procedure, after its declarations and before its first executable statement
            m = re.search(rf"subroutine\s+{proc}\s*\(.*?\n(.*?)if \(mom6cu_enabled\(\)\) then.*?call {proc}_mom6cu\(", text, flags=re.S | re.I)
            assert m, (rel, proc)
            between = m.group(1)
            assert not re.search(r"^\s*end\s+subroutine", between, flags=re.M | re.I), (rel, proc, "hook landed in a later procedure")
            assert not re.search(r"^\s*(call|do|if)\b", between, flags=re.M | re.I), (rel, proc, "an executable statement precedes the hook")
        assert text.count("use mom6cu_interface") == 1
        # nothing of the reference was removed: the shadow copy minus the inserted lines is the reference file
        ref_lines = open(os.path.join(REF, rel)).read().split("\n")
        it = iter(text.split("\n"))
        assert all(any(r == s for s in it) for r in ref_lines), rel


@pytest.mark.skipif(not have_ref, reason="needs the reference checkout")
def test_every_member_the_bindings_touch_exists():
    srcs = glob.glob(os.path.join(REF, "src", "**", "*.F90"), recursive=True) + [os.path.join(ROOT, "fortran", "mom6cu_interface.F90")]
    db = A.fortran_types(srcs)
    assert "mom_dyn_split_rk2_cs" in db and "barotropic_cs" in db and "mom6cu_step_dyn_args" in db and "bt_cont_type" in db
    ifaces = A.fortran_bindc_interfaces(os.path.join(ROOT, "fortran", "mom6cu_interface.F90"))
    helper = {"mom6cu_check", "mom6cu_enabled", "mom6cu_ctx"}
    checked = 0
    for inc in glob.glob(os.path.join(ROOT, "fortran", "bodies", "*.inc")):
        for name, dummies, types, body in A.fortran_procedures(open(inc).read()):
            code = "\n".join(ln.split("!")[0] for ln in body.split("\n"))
            code = re.sub(r'"[^"]*"', '""', code)
            for chain in set(re.findall(r"\b(\w+(?:%\w+)+)", code)):
                parts = chain.lower().split("%")
                t = types.get(parts[0])
                assert t is not None, (inc, name, chain, "base is not a declared derived-type variable")
                for mem in parts[1:]:
                    assert t in db, (inc, name, chain, f"type {t} not found")
                    assert mem in db[t], (inc, name, chain, f"{t} has no member {mem}")
                    t = db[t][mem]
                    checked += 1
            for f in set(re.findall(r"\b(mom6cu_\w+)\s*\(", code)):
                assert f in ifaces or f in helper or f.lower() in db or f.endswith("_mom6cu"), (inc, name, f)
    assert checked > 150

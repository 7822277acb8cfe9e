#!/usr/bin/env python
"""Builds the shadow source directory a MOM6 maintainer puts ahead of src/ in the source list (the build-time directory swap MOM6 uses
for config_src/infra and config_src/external, ac/configure.ac:227-264):

    python fortran/install_shims.py <MOM6 checkout> <output dir>

For every module of SHIMS it writes a copy of the reference's own source file with three insertions and nothing removed, so the module keeps
every public name, type and procedure the rest of MOM6 uses (round 1 shipped a hand-written MOM_continuity_PPM with 3 of its 17 public names):
  1. `use mom6cu_interface` / `use, intrinsic :: iso_c_binding` after the module statement;
  2. a hook at the first executable statement of each listed procedure:  if (mom6cu_enabled()) then ; call <proc>_mom6cu(<same dummies>) ;
     return ; endif -- the device path is taken once the driver has created the context (mom6cu_create), the reference body otherwise;
  3. the binding procedures of fortran/bodies/<module>.inc before `end module` (they are module procedures because the control structures
     are private to their modules), and their names added to the public list where another module calls them.
No reference source is stored in this repository: the script needs the checkout.  tests/test_fortran_shims.py runs it against the
reference and checks the public lists, the dummy-argument lists, the hooks and every CS%member the bodies touch."""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

# module file (relative to the checkout) -> hooks [(procedure, arguments passed on)] and extra public names
SHIMS = {
    "src/core/MOM_dynamics_split_RK2.F90": dict(
        hooks=[("step_MOM_dyn_split_RK2", "u_inst, v_inst, h, tv, visc, dt, forces, p_surf_begin, p_surf_end, uh, vh, uhtr, vhtr, eta_av, G, GV, US, CS, calc_dtbt")],
        public=["dyn_split_RK2_aux_fill_mom6cu"], uses=["use MOM_barotropic, only : barotropic_fill_mom6cu, barotropic_update_from_mom6cu",
                         "use MOM_continuity_PPM, only : continuity_PPM_send_cs_mom6cu",
                         "use MOM_CoriolisAdv, only : CoriolisAdv_send_cs_mom6cu", "use MOM_hor_visc, only : hor_visc_send_cs_mom6cu",
                         "use MOM_PressureForce, only : PressureForce_send_cs_mom6cu",
                         "use MOM_vert_friction, only : vertvisc_send_cs_mom6cu"]),
    "src/core/MOM_barotropic.F90": dict(
        hooks=[("btstep", "U_in, V_in, eta_in, dt, bc_accel_u, bc_accel_v, forces, pbce, eta_PF_in, U_Cor, V_Cor, accel_layer_u, accel_layer_v, eta_out, "
                          "uhbtav, vhbtav, G, GV, US, CS, visc_rem_u, visc_rem_v, SpV_avg, ADp, OBC, BT_cont, eta_PF_start, taux_bot, tauy_bot, uh0, vh0, "
                          "u_uh0, v_vh0, etaav")],
        public=["barotropic_fill_mom6cu", "barotropic_update_from_mom6cu"], uses=[]),
    "src/core/MOM_continuity_PPM.F90": dict(
        hooks=[("continuity_PPM", "u, v, hin, h, uh, vh, dt, G, GV, US, CS, OBC, pbv, uhbt, vhbt, visc_rem_u, visc_rem_v, u_cor, v_cor, BT_cont, du_cor, dv_cor")],
        public=["continuity_PPM_send_cs_mom6cu"], uses=[]),
    "src/core/MOM_CoriolisAdv.F90": dict(
        hooks=[("CorAdCalc", "u, v, h, uh, vh, CAu, CAv, OBC, AD, G, GV, US, CS, pbv, Waves")],
        public=["CoriolisAdv_send_cs_mom6cu"], uses=[]),
    "src/core/MOM_PressureForce_FV.F90": dict(
        hooks=[("PressureForce_FV_Bouss", "h, tv, PFu, PFv, G, GV, US, CS, ALE_CSp, ADp, p_atm, pbce, eta")],
        public=["PressureForce_FV_send_cs_mom6cu"],
        uses=["use MOM_EOS, only : EOS_query_mom6cu", "use MOM_ALE, only : ALE_answer_date_mom6cu"]),
    "src/core/MOM_PressureForce.F90": dict(
        hooks=[], public=["PressureForce_send_cs_mom6cu"],
        uses=["use MOM_PressureForce_FV, only : PressureForce_FV_send_cs_mom6cu"]),
    "src/parameterizations/vertical/MOM_vert_friction.F90": dict(
        hooks=[("vertvisc_coef", "u, v, h, dz, forces, visc, tv, dt, G, GV, US, CS, OBC, VarMix"),
               ("vertvisc", "u, v, h, forces, visc, dt, OBC, ADp, CDp, G, GV, US, CS, taux_bot, tauy_bot, fpmix, Waves"),
               ("vertvisc_remnant", "visc, visc_rem_u, visc_rem_v, dt, G, GV, US, CS")],
        public=["vertvisc_send_cs_mom6cu"], uses=[]),
    "src/parameterizations/lateral/MOM_hor_visc.F90": dict(
        hooks=[("horizontal_viscosity", "u, v, h, uh, vh, diffu, diffv, MEKE, VarMix, G, GV, US, CS, tv, dt, OBC, BT, TD, ADp, hu_cont, hv_cont, STOCH")],
        public=["hor_visc_send_cs_mom6cu"], uses=[]),
    "src/tracer/MOM_tracer_advect.F90": dict(
        hooks=[("advect_tracer", "h_end, uhtr, vhtr, OBC, dt, G, GV, US, CS, Reg, x_first_in, vol_prev, max_iter_in, update_vol_prev, uhr_out, vhr_out")],
        public=[], uses=[]),
    "src/tracer/MOM_tracer_hor_diff.F90": dict(
        hooks=[("tracer_hordiff", "h, dt, MEKE, VarMix, visc, G, GV, US, CS, Reg, tv, do_online_flag, read_khdt_x, read_khdt_y")],
        public=[], uses=[]),
    "src/parameterizations/lateral/MOM_thickness_diffuse.F90": dict(
        hooks=[("thickness_diffuse", "h, uhtr, vhtr, tv, dt, G, GV, US, MEKE, VarMix, CDp, CS, STOCH")],
        public=[], uses=["use MOM_EOS, only : EOS_query_mom6cu"]),
    "src/parameterizations/lateral/MOM_mixed_layer_restrat.F90": dict(
        hooks=[("mixedlayer_restrat", "h, uhtr, vhtr, tv, forces, dt, MLD, h_MLD, bflux, VarMix, G, GV, US, CS")],
        public=[], uses=["use MOM_EOS, only : EOS_query_mom6cu"]),
    # accessors only: these modules keep the members the bindings above need private
    "src/equation_of_state/MOM_EOS.F90": dict(hooks=[], public=["EOS_query_mom6cu"], uses=[]),
    "src/ALE/MOM_ALE.F90": dict(hooks=[], public=["ALE_answer_date_mom6cu", "ALE_fill_mom6cu", "ALE_set_old_grid_weight_mom6cu"],
                                uses=["use MOM_regridding, only : regridding_fill_mom6cu", "use MOM_remapping, only : remapping_fill_mom6cu"]),
    "src/ALE/MOM_regridding.F90": dict(hooks=[], public=["regridding_fill_mom6cu"], uses=[]),
    "src/ALE/MOM_remapping.F90": dict(hooks=[], public=["remapping_fill_mom6cu"], uses=[]),
    "src/core/MOM.F90": dict(
        hooks=[("ALE_regridding_and_remapping", "CS, G, GV, US, u, v, h, tv, dtdia, Time_end_thermo")], public=[],
        uses=["use MOM_ALE, only : ALE_fill_mom6cu, ALE_set_old_grid_weight_mom6cu",
              "use MOM_dynamics_split_RK2, only : dyn_split_RK2_aux_fill_mom6cu"]),
}

DECL = re.compile(r"^\s*(type\s*\(|class\s*\(|real\b|integer\b|logical\b|character\b|complex\b|double\s+precision\b|use\b|implicit\b|"
                  r"procedure\b|external\b|intrinsic\b|parameter\b|dimension\b|save\b|data\b|common\b|namelist\b|equivalence\b)", re.I)


def first_executable(lines, start):
    """Index of the first executable statement after the procedure statement that begins at lines[start]."""
    i = start
    while lines[i].split("!")[0].rstrip().endswith("&"):      # the (continued) procedure statement itself
        i += 1
    i += 1
    cont = False
    while i < len(lines):
        code = lines[i].split("!")[0].rstrip() if not lines[i].lstrip().startswith("!") else ""
        t = code.strip()
        if not t or t.startswith("#"):        # blank, comment-only and preprocessor lines neither start nor end a statement
            pass
        elif cont:
            cont = t.endswith("&")
        elif DECL.match(t) or "::" in t:
            cont = t.endswith("&")
        else:
            return i
        i += 1
    raise ValueError("no executable statement found")


def install_one(src_path, body_path, spec):
    lines = open(src_path).read().split("\n")
    mod_i = next(i for i, ln in enumerate(lines) if re.match(r"\s*module\s+\w+\s*$", ln, flags=re.I))
    modname = re.match(r"\s*module\s+(\w+)", lines[mod_i], flags=re.I).group(1)
    # 3. bodies before `end module`
    end_i = max(i for i, ln in enumerate(lines) if re.match(rf"\s*end\s+module\s+{modname}\b", ln, flags=re.I))
    body = open(body_path).read().rstrip("\n").split("\n")
    lines[end_i:end_i] = ["", "! ---- mom6cu: device-path bindings (fortran/bodies/%s) ----" % os.path.basename(body_path)] + body + [""]
    # 2. hooks (from the bottom up so that indices stay valid)
    found = []
    for proc, args in spec["hooks"]:
        cands = [i for i, ln in enumerate(lines) if re.match(rf"\s*(?:recursive\s+)?subroutine\s+{proc}\s*\(", ln, flags=re.I)]
        if len(cands) != 1:
            raise ValueError(f"{src_path}: expected one definition of {proc}, found {len(cands)}")
        found.append((cands[0], proc, args))
    for start, proc, args in sorted(found, reverse=True):
        at = first_executable(lines, start)
        hook = [f"  if (mom6cu_enabled()) then   ! mom6cu: the device path (fortran/bodies/{os.path.basename(body_path)})",
                f"    call {proc}_mom6cu({args})", "    return", "  endif", ""]
        # wrap the call at 120 columns with continuation lines
        wrapped = []
        for h in hook:
            while len(h) > 118:
                cut = h.rfind(",", 0, 116)
                wrapped.append(h[:cut + 1] + " &")
                h = " " * 8 + h[cut + 1:].lstrip()
            wrapped.append(h)
        lines[at:at] = wrapped
    # public names other modules call
    if spec["public"]:
        cont_i = next(i for i, ln in enumerate(lines) if re.match(r"\s*contains\s*$", ln, flags=re.I))
        pub_i = max(i for i in range(cont_i) if re.match(r"\s*public\b", lines[i], flags=re.I))
        while lines[pub_i].split("!")[0].rstrip().endswith("&"):
            pub_i += 1
        lines[pub_i + 1:pub_i + 1] = ["public " + ", ".join(spec["public"]) + "   ! mom6cu"]
    # 1. use statements right after the module statement
    lines[mod_i + 1:mod_i + 1] = ["use, intrinsic :: iso_c_binding   ! mom6cu", "use mom6cu_interface             ! mom6cu"] + [u + "   ! mom6cu" for u in spec["uses"]]
    return "\n".join(lines)


def install(mom6_root, out_dir):
    os.makedirs(out_dir, exist_ok=True)
    written = []
    for rel, spec in SHIMS.items():
        src = os.path.join(mom6_root, rel)
        body = os.path.join(HERE, "bodies", os.path.basename(rel).replace(".F90", ".inc"))
        text = install_one(src, body, spec)
        dst = os.path.join(out_dir, os.path.basename(rel))
        open(dst, "w").write(text)
        written.append(dst)
    dst = os.path.join(out_dir, "mom6cu_interface.F90")
    open(dst, "w").write(open(os.path.join(HERE, "mom6cu_interface.F90")).read())
    written.append(dst)
    return written


if __name__ == "__main__":
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    for f in install(sys.argv[1], sys.argv[2]):
        print("wrote", f)

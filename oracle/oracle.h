/* TEST INFRASTRUCTURE ONLY.
 * CPU oracle: a plain C++ restatement of the reference Fortran (mom-ocean/MOM6 @ b18bca24)
 * for the split-explicit dycore hot path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (mom6_b200/) never links, imports or calls it.
 *
 * Parity status: the reference cannot be compiled in the build container (no Fortran
 * compiler, MPI, netCDF or FMS), so every routine here follows the cited Fortran lines
 * loop-for-loop with identical parenthesisation and is compiled -O2 -ffp-contract=off.
 *   - ALE remapping: PINNED against the reference's known-answer vectors
 *     (src/ALE/MOM_remapping.F90:2072+), see tests/test_oracle_remap_kat.py.
 *   - the dycore stages, the whole step, tracer advection, the ALE pass and the three callers: PINNED BY A REFERENCE RUN --
 *     oracle/f90run executes the reference's own Fortran source and tests/test_reference_f90.py compares bit for bit;
 *   - write_energy, create_depth_list and the bit-count checksums: PINNED BY A REFERENCE RUN as well (tests/refcases.py "diag/...").
 *     The ocean.stats record comes out of the reference's own WRITE statement (its format string and output list), edited by
 *     f90run's restatement of Fortran format-directed output -- the Fortran standard's rules, not libgfortran itself.
 */
#ifndef MOM6_ORACLE_H
#define MOM6_ORACLE_H
#include "../include/mom6cu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Optional hook standing in for do_group_pass(CS%pass_eta_ubt, CS%BT_Domain)
 * (MOM_barotropic.F90:2512).  NULL = single tile: cyclic wrap where the domain
 * says cyclic, no-op at closed edges (what mpp_update_domains does there). */
typedef void (*oracle_halo_fn)(void* user, double* eta, double* ubt, double* vbt);

/* btstep_timeloop, MOM_barotropic.F90:2175-2832 */
int oracle_btstep_timeloop(const mom6cu_domain* dom, const mom6cu_bt_timeloop_args* a,
                           oracle_halo_fn halo, void* user, int nthreads);

/* single-tile halo fill used by the default hook and by tests:
 * stagger 0=h,1=u,2=v,3=q; fills halo points of a wide (wide=1) or G-sized array */
void oracle_fill_halo_2d(const mom6cu_domain* dom, double* f, int stagger, int wide);

/* continuity_PPM, MOM_continuity_PPM.F90:86-194 */
int oracle_continuity(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV,
                      const mom6cu_continuity_cs* CS, const mom6cu_continuity_args* a, int nthreads);

/* CorAdCalc, MOM_CoriolisAdv.F90:125-965 (US may be NULL = unscaled) */
int oracle_coradcalc(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV,
                     const mom6cu_unit_scale* US, const mom6cu_coriolisadv_cs* CS,
                     const mom6cu_coradcalc_args* a, int nthreads);

/* horizontal_viscosity, MOM_hor_visc.F90:266-2317 (frozen option set) */
int oracle_horizontal_viscosity(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV,
                                const mom6cu_hor_visc_cs* CS, const mom6cu_hor_visc_args* a, int nthreads);

/* btstep, MOM_barotropic.F90:455-2172 (frozen option set, single tile); btcalc :4360; bt_mass_source :5243 */
int oracle_btstep(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV,
                  const mom6cu_barotropic_cs* CS, const mom6cu_btstep_args* a, int nthreads);
int oracle_btcalc(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV,
                  const mom6cu_btcalc_args* a, int nthreads);
int oracle_bt_mass_source(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV, const double* h,
                          const double* eta, int set_cor, double* eta_cor);

int oracle_set_dtbt(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV, const mom6cu_unit_scale* US,
                    const mom6cu_set_dtbt_args* a, double* dtbt, double* dtbt_max);

/* PressureForce_FV_Bouss, MOM_PressureForce_FV.F90:947-2017 (frozen option set) */
int oracle_pressure_force(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV,
                          const mom6cu_pressureforce_cs* CS, const mom6cu_pressureforce_args* a, int nthreads);

/* ALE remapping (MOM_remapping.F90, PLM/PPM_functions.F90, regrid_edge_values.F90, MOM_ALE.F90): see remap.cpp.
 * PINNED by the reference's known-answer vectors (tests/test_oracle_remap_kat.py). */
int oracle_remapping_core_h(const mom6cu_remapping_cs* CS, int n0, const double* h0, const double* u0, int n1, const double* h1,
                            double* u1, double* net_err);
int oracle_remap_intersect(int n0, const double* h0, int n1, const double* h1, double* h_sub, double* h0_eff, int* isrc_start,
                           int* isrc_end, int* isrc_max, int* itgt_start, int* itgt_end, int* isub_src);
int oracle_remap_reconstruct(int which, int N, const double* h, const double* u, double h_neglect, double* E, double* coefs);
int oracle_remap_plm_sub(int om4, int n0, const double* h0, const double* u0, int n1, const double* h1, double h_neglect, double* u_sub,
                         double* u1);
int oracle_ale_remap_scalar(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_remapping_cs* CS, const double* h_old,
                            const double* h_new, double* field, double conc_underflow, int nthreads);
int oracle_ale_remap_set_h_vel(const mom6cu_domain* dom, const mom6cu_grid* G, const double* h_new, double* h_u, double* h_v);
int oracle_ale_remap_velocities(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_remapping_cs* CS, const double* h_old_u,
                                const double* h_old_v, const double* h_new_u, const double* h_new_v, double* u, double* v, int nthreads);

int oracle_remap_dyn_split_rk2_aux_vars(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_remapping_cs* remapCS,
                                        const mom6cu_dyn_split_rk2_cs* CS, const double* h_old_u, const double* h_old_v,
                                        const double* h_new_u, const double* h_new_v, int nthreads);

/* advect_tracer (MOM_tracer_advect.F90:53-1152): see advect.cpp.  Pinned by a reference run.  *iterations returns the number of passes made. */
int oracle_advect_tracer(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV, const mom6cu_tracer_advect_cs* CS,
                         const mom6cu_advect_tracer_args* a, int* iterations);

/* ALE_regrid, Z* (MOM_ALE.F90:518, MOM_regridding.F90:846-1857, coord_zlike.F90:63): see regrid.cpp.  Pinned by a reference run. */
int oracle_ale_regrid(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV, const mom6cu_unit_scale* US,
                      const mom6cu_regridding_cs* CS, const double* h, double* h_new, double* dzRegrid);

/* vertvisc_coef / vertvisc / vertvisc_remnant (MOM_vert_friction.F90:557-2924): see vertvisc.cpp.  Pinned by a reference run.
 * The CS%a_u, a_v (nk+1 levels), h_u, h_v arrays the reference keeps in vertvisc_CS are explicit arguments here. */
int oracle_vertvisc_coef(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV, const mom6cu_unit_scale* US,
                         const mom6cu_vertvisc_cs* CS, const mom6cu_vertvisc_coef_args* a, double* a_u, double* a_v, double* h_u, double* h_v);
int oracle_vertvisc(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV, const mom6cu_vertvisc_cs* CS,
                    const mom6cu_vertvisc_args* a, const double* a_u, const double* a_v, const double* h_u, const double* h_v);
long long oracle_vertvisc_ntrunc(int reset); /* CS%ntrunc counted by oracle_vertvisc (vertvisc_limit_vel :2926) since the last reset */
int oracle_vertvisc_remnant(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vertvisc_cs* CS, const double* Ray_u,
                            const double* Ray_v, double* visc_rem_u, double* visc_rem_v, double dt, const double* a_u, const double* a_v,
                            const double* h_u, const double* h_v);

/* step_MOM_dyn_split_RK2 (MOM_dynamics_split_RK2.F90:294-1205): see step.cpp.  Pinned by a reference run. */
int oracle_step_dyn_split_rk2(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV, const mom6cu_unit_scale* US,
                              const mom6cu_continuity_cs* cont_cs, const mom6cu_coriolisadv_cs* corad_cs, const mom6cu_hor_visc_cs* hv_cs,
                              const mom6cu_pressureforce_cs* pgf_cs, const mom6cu_vertvisc_cs* vv_cs, mom6cu_dyn_split_rk2_cs* CS,
                              const mom6cu_step_dyn_args* a, int nthreads);

/* Extended-fixed-point sums, bit-count checksums (efp.cpp) and write_energy (sum_output.cpp): the reference's
 * answer-reproducibility metric.  The sums are PINNED by the reference's unit test test_reproducing_sum.F90
 * (tests/test_oracle_efp.py). */
int oracle_reproducing_sum(const mom6cu_domain* dom, const double* array, int stagger, int nk, int isr, int ier, int jsr, int jer,
                           double unscale, int reproducing, int overflow_check, double* sum, double* sums, mom6cu_efp* EFP_sum,
                           mom6cu_efp* EFP_lay_sums);
void oracle_efp_plus(const mom6cu_efp* a, const mom6cu_efp* b, mom6cu_efp* out, int* overflow);
void oracle_efp_minus(const mom6cu_efp* a, const mom6cu_efp* b, mom6cu_efp* out, int* overflow);
double oracle_efp_to_real(mom6cu_efp* a);
int oracle_real_to_efp(double v, mom6cu_efp* out);
double oracle_efp_real_diff(const mom6cu_efp* a, const mom6cu_efp* b);
int oracle_chksum(const mom6cu_domain* d, const double* array, int stagger, int nk, int haloshift, int symmetric, int omit_corners,
                  double scale, int* bc, int* kind, double* stats);
int oracle_create_depth_list(const mom6cu_domain* dom, const mom6cu_grid* G, double Z_ref, double min_depth_inc, int* listsize,
                             double* depth, double* area, double* vol_below);
int oracle_write_energy(const mom6cu_domain* dom, const mom6cu_grid* G, const mom6cu_vgrid* GV, mom6cu_sum_output_cs* CS,
                        const double* u, const double* v, const double* h, const double* T, const double* S, mom6cu_energy_out* out);
int oracle_ocean_stats_line(const mom6cu_sum_output_cs* CS, const mom6cu_energy_out* e, int n, double reday, char* buf, size_t len);

/* interpolate_column (MOM_remapping.F90:1247; PINNED by the vectors at :2648-2682), ALE_remap_interface_vals / _vertex_vals
 * (MOM_ALE.F90:1303, :1342) and the order of operations of ALE_regridding_and_remapping (MOM.F90:1751-1926): ale_chain.cpp. */
void oracle_interpolate_column(int nsrc, const double* h_src, const double* u_src, int ndest, const double* h_dest, double* u_dest,
                               int mask_edges);
int oracle_ale_remap_interface_vals(const mom6cu_domain* d, const mom6cu_grid* G, const double* h_old, const double* h_new, double* int_val);
int oracle_ale_remap_vertex_vals(const mom6cu_domain* d, const mom6cu_grid* G, const double* h_old, const double* h_new, double* vert_val);
int oracle_ale_regridding_and_remapping(const mom6cu_domain* d, const mom6cu_grid* G, const mom6cu_vgrid* GV, const mom6cu_unit_scale* US,
                                        mom6cu_ale_cs* CS, const mom6cu_dyn_split_rk2_cs* dynCS, const mom6cu_ale_args* a, int nthreads);

/* mixedlayer_restrat -> mixedlayer_restrat_OM4 (MOM_mixed_layer_restrat.F90:149-714) and mu (:717; PINNED by the unit test
 * vectors at :2022-2041): mle.cpp. */
double oracle_mle_mu(double sigma, double dh);
int oracle_mixedlayer_restrat(const mom6cu_domain* d, const mom6cu_grid* G, const mom6cu_vgrid* GV, mom6cu_mle_cs* CS, double* h,
                              double* uhtr, double* vhtr, const double* T, const double* S, const double* ustar, double dt,
                              const double* h_MLD, const double* Rd_dx_h);

/* tracer_hordiff, the along-surface path (src/tracer/MOM_tracer_hor_diff.F90:119-640): hordiff.cpp.  num_itts (may be NULL) returns
 * the number of iterations made. */
int oracle_tracer_hordiff(const mom6cu_domain* d, const mom6cu_grid* G, const mom6cu_vgrid* GV, const mom6cu_tracer_hor_diff_cs* CS,
                          const mom6cu_tracer_hordiff_args* a, int* num_itts);

/* thickness_diffuse -> thickness_diffuse_full (MOM_thickness_diffuse.F90:134-1670), find_eta (MOM_interface_heights.F90:48), vert_fill_TS
 * (MOM_isopycnal_slopes.F90:612), calculate_density_derivs (MOM_EOS_Wright.F90:178, MOM_EOS_linear.F90): thickdiff.cpp. */
int oracle_thickness_diffuse(const mom6cu_domain* d, const mom6cu_grid* G, const mom6cu_vgrid* GV, const mom6cu_unit_scale* US,
                             const mom6cu_thickness_diffuse_cs* CS, const mom6cu_thickness_diffuse_args* a);

/* ALE_PLM_edge_values / one field of TS_PPM_edge_values (MOM_ALE.F90:1518-1660) for one column: scheme 1 = PLM, 2 = PPM. */
int oracle_ale_edge_values(int scheme, int nk, const double* h, const double* Q, int bdry_extrap, double h_neglect, double* Q_t, double* Q_b);

/* Equation-of-state elements (oracle/eos.hpp: EOS_WRIGHT, EOS_LINEAR and the unit-rescaling wrapper of MOM_EOS.F90:308-354), PINNED by the
 * reference's check values (EOS_unit_tests, MOM_EOS.F90:2075-2131), and the copies mle.cpp / thickdiff.cpp carry. */
double oracle_eos_eval(int which, int form, const double* lin4, const double* scales, double T, double S, double p, double rho_ref);
double oracle_mle_eos_density(int form, const double* lin4, double T, double S, double p);
void oracle_thickdiff_eos_derivs(int form, const double* lin4, double T, double S, double p, double* drho_dT, double* drho_dS);

#ifdef __cplusplus
}
#endif
#endif

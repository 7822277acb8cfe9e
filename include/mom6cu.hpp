// C++ host-side mirror of the reference's module procedures over the C ABI (include/mom6cu.h).
//
// The reference's host language is Fortran (fortran/mom6cu_interface.F90 holds the ISO_C_BINDING binding); this image has no
// Fortran compiler, so the compiled-language host side that can be built and exercised here is this header: one member
// function per reference subroutine, under the reference's own name, with the reference's error convention -- MOM_error(FATAL)
// (src/framework/MOM_error_handler.F90) becomes a mom6cu::Fatal exception carrying the message, WARNINGs are counted.  There is
// no CPU fallback: constructing a Context without a CUDA device throws.  Header-only; link with -lmom6cu.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "mom6cu.h"

namespace mom6cu {

struct Fatal : std::runtime_error {
  int code;
  Fatal(int rc, const std::string& where, const std::string& msg) : std::runtime_error(where + ": " + msg), code(rc) {}
};

enum Stagger { H_POINT = 0, U_POINT = 1, V_POINT = 2, Q_POINT = 3 };

class Context;

// A device-resident field (mom6cu_plane_*): usable wherever an entry takes an array.
class Plane {
 public:
  double* ptr() const { return p_; }
  operator double*() const { return p_; }
  void upload(const double* host);
  void download(double* host) const;

 private:
  friend class Context;
  Plane(Context* c, double* p, int st, int wide, int nk) : c_(c), p_(p), st_(st), wide_(wide), nk_(nk) {}
  Context* c_;
  double* p_;
  int st_, wide_, nk_;
};

class Context {
 public:
  // MOM_domains / hor_index_init: one context per PE (= one tile = one GPU)
  Context(const mom6cu_domain& dom, int device = 0) {
    const int rc = mom6cu_create(&h_, &dom, device);
    if (rc != 0 || !h_)
      throw Fatal(rc, "mom6cu_create", rc == MOM6CU_ERR_NO_DEVICE ? "no CUDA device: the hot path has no CPU fallback" : "could not create the device context");
  }
  ~Context() { if (h_) mom6cu_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;

  mom6cu_ctx* handle() const { return h_; }
  int warnings() const { return warnings_; }
  long long launch_count() const { return mom6cu_launch_count(h_); }
  double last_kernel_ms() const { return mom6cu_last_kernel_ms(h_); }
  void sync() { check(mom6cu_sync(h_), "mom6cu_sync"); }

  Plane plane(const char* name, Stagger st, int nk = 1, bool wide = false) {
    double* p = mom6cu_plane_alloc(h_, name, nk);
    if (!p) throw Fatal(MOM6CU_ERR_CUDA, "mom6cu_plane_alloc", last_error());
    return Plane(this, p, (int)st, wide ? 1 : 0, nk);
  }

  // ---- initialisation (the *_init routines of the reference hand over their resolved control structures)
  void set_grid(const mom6cu_grid& G) { check(mom6cu_set_grid(h_, &G), "set_grid"); }                              // MOM_grid.F90:20
  void set_verticalGrid(const mom6cu_vgrid& GV) { check(mom6cu_set_vgrid(h_, &GV), "set_verticalGrid"); }            // MOM_verticalGrid.F90
  void set_unit_scale(const mom6cu_unit_scale& US) { check(mom6cu_set_unit_scale(h_, &US), "set_unit_scale"); }      // MOM_unit_scaling.F90
  void continuity_PPM_init(const mom6cu_continuity_cs& CS) { check(mom6cu_set_cs_continuity(h_, &CS), "continuity_PPM_init"); }    // MOM_continuity_PPM.F90:2674
  void CoriolisAdv_init(const mom6cu_coriolisadv_cs& CS) { check(mom6cu_set_cs_coriolisadv(h_, &CS), "CoriolisAdv_init"); }        // MOM_CoriolisAdv.F90:1054
  void hor_visc_init(const mom6cu_hor_visc_cs& CS) { check(mom6cu_set_cs_hor_visc(h_, &CS), "hor_visc_init"); }                    // MOM_hor_visc.F90:2322
  void PressureForce_init(const mom6cu_pressureforce_cs& CS) { check(mom6cu_set_cs_pressureforce(h_, &CS), "PressureForce_init"); }  // MOM_PressureForce.F90:85
  void vertvisc_init(const mom6cu_vertvisc_cs& CS) { check(mom6cu_set_cs_vertvisc(h_, &CS), "vertvisc_init"); }                    // MOM_vert_friction.F90:2929

  // ---- the dycore stages, under the reference's names
  void step_MOM_dyn_split_RK2(mom6cu_dyn_split_rk2_cs& CS, const mom6cu_step_dyn_args& a) {                       // MOM_dynamics_split_RK2.F90:294
    check(mom6cu_step_dyn_split_rk2(h_, &CS, &a), "step_MOM_dyn_split_RK2");
  }
  void continuity(const mom6cu_continuity_args& a) { check(mom6cu_continuity(h_, &a), "continuity_PPM"); }            // MOM_continuity_PPM.F90:86
  void CorAdCalc(const mom6cu_coradcalc_args& a) { check(mom6cu_coradcalc(h_, &a), "CorAdCalc"); }                    // MOM_CoriolisAdv.F90:125
  void PressureForce(const mom6cu_pressureforce_args& a) { check(mom6cu_pressure_force(h_, &a), "PressureForce"); }   // MOM_PressureForce.F90:40
  void horizontal_viscosity(const mom6cu_hor_visc_args& a) { check(mom6cu_horizontal_viscosity(h_, &a), "horizontal_viscosity"); }  // MOM_hor_visc.F90:266
  void btstep(const mom6cu_barotropic_cs& CS, const mom6cu_btstep_args& a) { check(mom6cu_btstep(h_, &CS, &a), "btstep"); }        // MOM_barotropic.F90:455
  void btcalc(const mom6cu_btcalc_args& a) { check(mom6cu_btcalc(h_, &a), "btcalc"); }                                // MOM_barotropic.F90:4360
  void bt_mass_source(const double* h, const double* eta, bool set_cor, double* eta_cor) {                           // MOM_barotropic.F90:5243
    check(mom6cu_bt_mass_source(h_, h, eta, set_cor ? 1 : 0, eta_cor), "bt_mass_source");
  }
  void set_dtbt(const mom6cu_set_dtbt_args& a, double& dtbt, double& dtbt_max) { check(mom6cu_set_dtbt(h_, &a, &dtbt, &dtbt_max), "set_dtbt"); }  // :3509
  void btstep_timeloop(const mom6cu_bt_timeloop_args& a) { check(mom6cu_btstep_timeloop(h_, &a), "btstep_timeloop"); }  // MOM_barotropic.F90:2175
  void vertvisc_coef(const mom6cu_vertvisc_coef_args& a) { check(mom6cu_vertvisc_coef(h_, &a), "vertvisc_coef"); }    // MOM_vert_friction.F90:1357
  void vertvisc(const mom6cu_vertvisc_args& a) { check(mom6cu_vertvisc(h_, &a), "vertvisc"); }                        // MOM_vert_friction.F90:557
  void vertvisc_remnant(const double* Ray_u, const double* Ray_v, double* visc_rem_u, double* visc_rem_v, double dt) {  // :1229
    check(mom6cu_vertvisc_remnant(h_, Ray_u, Ray_v, visc_rem_u, visc_rem_v, dt), "vertvisc_remnant");
  }

  // ---- tracer advection, ALE, the callers between the dycore and the tracer step
  int advect_tracer(const mom6cu_tracer_advect_cs& CS, const mom6cu_advect_tracer_args& a) {                         // MOM_tracer_advect.F90:53
    check(mom6cu_advect_tracer(h_, &CS, &a), "advect_tracer");
    return mom6cu_last_iterations(h_);
  }
  void ALE_regrid(const mom6cu_regridding_cs& CS, const double* h, double* h_new, double* dzRegrid) {                // MOM_ALE.F90:518
    check(mom6cu_ale_regrid(h_, &CS, h, h_new, dzRegrid), "ALE_regrid");
  }
  void ALE_remap_tracers(const mom6cu_remapping_cs& CS, const double* h_old, const double* h_new, const std::vector<double*>& tr,
                         const double* conc_underflow = nullptr) {                                                  // MOM_ALE.F90:760
    check(mom6cu_ale_remap_tracers(h_, &CS, h_old, h_new, (int)tr.size(), const_cast<double**>(tr.data()), conc_underflow), "ALE_remap_tracers");
  }
  void ALE_remap_set_h_vel(const double* h_new, double* h_u, double* h_v) { check(mom6cu_ale_remap_set_h_vel(h_, h_new, h_u, h_v), "ALE_remap_set_h_vel"); }  // :882
  void ALE_remap_velocities(const mom6cu_remapping_cs& CS, const double* h_old_u, const double* h_old_v, const double* h_new_u,
                            const double* h_new_v, double* u, double* v) {                                           // MOM_ALE.F90:1089
    check(mom6cu_ale_remap_velocities(h_, &CS, h_old_u, h_old_v, h_new_u, h_new_v, u, v), "ALE_remap_velocities");
  }
  void ALE_regridding_and_remapping(mom6cu_ale_cs& CS, const mom6cu_dyn_split_rk2_cs* dynCS, const mom6cu_ale_args& a) {  // MOM.F90:1751
    check(mom6cu_ale_regridding_and_remapping(h_, &CS, dynCS, &a), "ALE_regridding_and_remapping");
  }
  void remapping_core_h(const mom6cu_remapping_cs& CS, int ncol, int n0, const double* h0, const double* u0, int n1, const double* h1,
                        double* u1) {                                                                                // MOM_remapping.F90:234
    check(mom6cu_remapping_core_h(h_, &CS, ncol, n0, h0, u0, n1, h1, u1), "remapping_core_h");
  }
  void mixedlayer_restrat(mom6cu_mle_cs& CS, double* h, double* uhtr, double* vhtr, const double* T, const double* S, const double* ustar,
                          double dt, const double* h_MLD, const double* Rd_dx_h = nullptr) {                         // MOM_mixed_layer_restrat.F90:149
    check(mom6cu_mixedlayer_restrat(h_, &CS, h, uhtr, vhtr, T, S, ustar, dt, h_MLD, Rd_dx_h), "mixedlayer_restrat");
  }
  void thickness_diffuse(const mom6cu_thickness_diffuse_cs& CS, const mom6cu_thickness_diffuse_args& a) {           // MOM_thickness_diffuse.F90:134
    check(mom6cu_thickness_diffuse(h_, &CS, &a), "thickness_diffuse");
  }
  int tracer_hordiff(const mom6cu_tracer_hor_diff_cs& CS, const mom6cu_tracer_hordiff_args& a) {                     // MOM_tracer_hor_diff.F90:119
    check(mom6cu_tracer_hordiff(h_, &CS, &a), "tracer_hordiff");
    return mom6cu_last_iterations(h_);
  }
  std::vector<double> mu(const std::vector<double>& sigma, const std::vector<double>& dh) {                          // MOM_mixed_layer_restrat.F90:717
    std::vector<double> out(sigma.size());
    check(mom6cu_mle_mu(h_, (int)sigma.size(), sigma.data(), dh.data(), out.data()), "mu");
    return out;
  }

  void do_group_pass(const std::vector<double*>& fields, const std::vector<int>& stagger, int nk) {                   // MOM_domains.F90 pass_var / do_group_pass
    check(mom6cu_do_group_pass(h_, (int)fields.size(), fields.data(), stagger.data(), nk), "do_group_pass");
  }

  void ALE_remap_interface_vals(const double* h_old, const double* h_new, double* int_val) {                        // MOM_ALE.F90:1303
    check(mom6cu_ale_remap_interface_vals(h_, h_old, h_new, int_val), "ALE_remap_interface_vals");
  }
  void ALE_remap_vertex_vals(const double* h_old, const double* h_new, double* vert_val) {                          // MOM_ALE.F90:1342
    check(mom6cu_ale_remap_vertex_vals(h_, h_old, h_new, vert_val), "ALE_remap_vertex_vals");
  }
  void interpolate_column(int ncol, int nsrc, const double* h_src, const double* u_src, int ndest, const double* h_dest, double* u_dest,
                          bool mask_edges) {                                                                       // MOM_remapping.F90:1247
    check(mom6cu_interpolate_column(h_, ncol, nsrc, h_src, u_src, ndest, h_dest, u_dest, mask_edges ? 1 : 0), "interpolate_column");
  }
  void remap_dyn_split_RK2_aux_vars(const mom6cu_remapping_cs& remapCS, const mom6cu_dyn_split_rk2_cs& CS, const double* h_old_u,
                                    const double* h_old_v, const double* h_new_u, const double* h_new_v) {          // MOM_dynamics_split_RK2.F90:1302
    check(mom6cu_remap_dyn_split_rk2_aux_vars(h_, &remapCS, &CS, h_old_u, h_old_v, h_new_u, h_new_v), "remap_dyn_split_RK2_aux_vars");
  }
  void vertvisc_get_coef(double* a_u, double* a_v, double* h_u, double* h_v) { check(mom6cu_vertvisc_get_coef(h_, a_u, a_v, h_u, h_v), "vertvisc_get_coef"); }
  void btstep_timeloop_resident(const mom6cu_bt_timeloop_args& a, int reps, bool download) {                         // MOM_barotropic.F90:2175, repeated on resident fields
    check(mom6cu_btstep_timeloop_resident(h_, &a, reps, download ? 1 : 0), "btstep_timeloop");
  }
  double total_kernel_ms() const { return mom6cu_total_kernel_ms(h_); }
  std::vector<double> last_step_stage_ms() const {
    std::vector<double> ms(8, 0.0);
    const int n = mom6cu_last_step_stage_ms(h_, ms.data(), (int)ms.size());
    ms.resize(n < 0 ? 0 : (n > 8 ? 8 : n));
    return ms;
  }
  void comm_destroy() { check(mom6cu_comm_destroy(h_), "comm_destroy"); }
  static int build_arch() { return mom6cu_build_arch(); }
  // one neighbour message of a group pass (host-only planning; MOM_domain_infra.F90:171-216): the peer rank or -1
  static int halo_plan(const mom6cu_domain& dom, Stagger st, bool wide, int halo, int dir, int send_box[4], int recv_box[4]) {
    return mom6cu_halo_plan(&dom, (int)st, wide ? 1 : 0, halo, dir, send_box, recv_box);
  }

  // ---- the answer-reproducibility metric
  void write_energy(mom6cu_sum_output_cs& CS, const double* u, const double* v, const double* h, const double* T, const double* S,
                    mom6cu_energy_out& out) {                                                                        // MOM_sum_output.F90:321
    check(mom6cu_write_energy(h_, &CS, u, v, h, T, S, &out), "write_energy");
  }
  static std::string ocean_stats_line(const mom6cu_sum_output_cs& CS, const mom6cu_energy_out& e, int n, double reday) {  // :874-902
    char buf[512];
    if (mom6cu_ocean_stats_line(&CS, &e, n, reday, buf, sizeof buf) != 0) throw Fatal(MOM6CU_ERR_BAD_ARG, "ocean_stats_line", "buffer too small");
    return std::string(buf);
  }

  std::string last_error() const {
    char buf[1024] = {0};
    mom6cu_last_error(h_, buf, sizeof buf);
    return std::string(buf);
  }

  // rc = 0 ok; > 0 FATAL (MOM_error(FATAL, msg)); < 0 minus the number of WARNINGs
  void check(int rc, const char* where) {
    if (rc > 0) throw Fatal(rc, where, last_error());
    if (rc < 0) warnings_ += -rc;
  }

 private:
  mom6cu_ctx* h_ = nullptr;
  int warnings_ = 0;
};

// EFP_type and its operators (src/framework/MOM_coms.F90:76-78, :548-684)
struct EFP {
  mom6cu_efp v;
  EFP() : v() {}
  explicit EFP(double x) : v() { const int rc = mom6cu_real_to_efp(x, &v); if (rc) throw Fatal(rc, "real_to_EFP", rc == 2 ? "NaN in real_to_EFP" : "Overflow in real_to_EFP conversion"); }
  double to_real() const { mom6cu_efp t = v; return mom6cu_efp_to_real(&t); }
  EFP operator+(const EFP& o) const { EFP r; int over = 0; mom6cu_efp_plus(&v, &o.v, &r.v, &over); if (over) throw Fatal(2, "EFP_plus", "Overflow in EFP_plus."); return r; }
  EFP operator-(const EFP& o) const { EFP r; int over = 0; mom6cu_efp_minus(&v, &o.v, &r.v, &over); if (over) throw Fatal(2, "EFP_minus", "Overflow in EFP_minus."); return r; }
  double real_diff(const EFP& o) const { return mom6cu_efp_real_diff(&v, &o.v); }
};

inline void Plane::upload(const double* host) { c_->check(mom6cu_plane_upload(c_->handle(), p_, host, st_, wide_, nk_), "plane_upload"); }
inline void Plane::download(double* host) const { c_->check(mom6cu_plane_download(c_->handle(), p_, host, st_, wide_, nk_), "plane_download"); }

}  // namespace mom6cu

// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of /root/reference/src/diagnostics/MOM_sum_output.F90: write_energy :321-1020 (Boussinesq; tracer stocks and
// min/max locations not included), create_depth_list :1203-1326 (single PE), the lH initialisation of depth_list_setup
// :1194-1196, and the ocean.stats line :874-902.
// PARITY: PINNED BY A REFERENCE RUN -- write_energy (three successive calls on a changing state, with and without temperature,
// with and without the APE calculation, with rescaled units) and create_depth_list of MOM_sum_output.F90 itself, executed by
// oracle/f90run, agree bit for bit with this file (tests/refcases.py "diag/write_energy*"), and so does the record appended to
// ocean.stats: the reference's WRITE statement (:880-905) with its own format string and output list, edited by f90run's
// restatement of Fortran format-directed output (the standard's rules; libgfortran itself is not executed).
#include "oracle.h"
#include "ogrid.hpp"
#include "efp.hpp"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace orc;

// create_depth_list :1203-1326.  depth/area/vol_below must hold niglobal*njglobal + 2 entries; *listsize returns DL%listsize.
extern "C" int oracle_create_depth_list(const mom6cu_domain* dom, const mom6cu_grid* Gh, double Z_ref, double min_depth_inc,
                                        int* listsize, double* depth, double* area_out, double* vol_below) {
  OGrid G(dom, Gh);
  const int niglobal = G.iec - G.isc + 1, njglobal = G.jec - G.jsc + 1;  // one tile: the global domain
  const int mls = niglobal * njglobal;
  std::vector<double> Dlist(mls + 2, 0.0), AreaList(mls + 2, 0.0);  // 1-based
  std::vector<int> indx2(mls + 2, 0);
  for (int j = G.jsc; j <= G.jec; ++j) for (int i = G.isc; i <= G.iec; ++i) {
    const int j_global = j - G.jsc + 1, i_global = i - G.isc + 1;
    const int list_pos = (j_global - 1) * niglobal + i_global;
    Dlist[list_pos] = G.bathyT(i, j) + Z_ref;
    AreaList[list_pos] = G.mask2dT(i, j) * G.areaT(i, j);
  }
  for (int j = 1; j <= mls + 1; ++j) indx2[j] = j;
  int k = mls / 2 + 1, ir = mls;
  for (;;) {  // heap sort :1243-1266
    int indxt; double Dnow;
    if (k > 1) { k = k - 1; indxt = indx2[k]; Dnow = Dlist[indxt]; }
    else {
      indxt = indx2[ir]; Dnow = Dlist[indxt];
      indx2[ir] = indx2[1];
      ir = ir - 1;
      if (ir == 1) { indx2[1] = indxt; break; }
    }
    int i = k, j = k * 2;
    for (;;) {
      if (j > ir) break;
      if (j < ir && Dlist[indx2[j]] < Dlist[indx2[j + 1]]) j = j + 1;
      if (Dnow < Dlist[indx2[j]]) { indx2[i] = indx2[j]; i = j; j = j + i; }
      else j = ir + 1;
    }
    indx2[i] = indxt;
  }
  double D_list_prev = Dlist[indx2[mls]];
  int list_size = 2;
  for (k = mls - 1; k >= 1; --k)
    if (Dlist[indx2[k]] < D_list_prev - min_depth_inc) { list_size = list_size + 1; D_list_prev = Dlist[indx2[k]]; }
  const int DLsize = list_size + 1;
  double vol = 0.0, area = 0.0;
  double Dprev = Dlist[indx2[mls]];
  D_list_prev = Dprev;
  int kl = 0;
  for (k = mls; k >= 1; --k) {
    const int i = indx2[k];
    vol = vol + area * (Dprev - Dlist[i]);
    area = area + AreaList[i];
    bool add_to_list = false;
    if ((kl == 0) || (k == 1)) add_to_list = true;
    else if (Dlist[indx2[k - 1]] < D_list_prev - min_depth_inc) { add_to_list = true; D_list_prev = Dlist[indx2[k - 1]]; }
    if (add_to_list) { kl = kl + 1; depth[kl - 1] = Dlist[i]; area_out[kl - 1] = area; vol_below[kl - 1] = vol; }
    Dprev = Dlist[i];
  }
  while (kl + 1 < DLsize) {
    kl = kl + 1;
    vol_below[kl - 1] = vol_below[kl - 2] * 1.000001;
    area_out[kl - 1] = area_out[kl - 2];
    depth[kl - 1] = depth[kl - 2];
  }
  vol_below[DLsize - 1] = vol_below[DLsize - 2] * 1000.0;
  area_out[DLsize - 1] = area_out[DLsize - 2];
  depth[DLsize - 1] = depth[DLsize - 2];
  *listsize = DLsize;
  return 0;
}

// write_energy :321-1020.  Returns 0 or a FATAL: 3 unsupported (non-Boussinesq), 2x the sums' FATALs, 41 NaN energy.
extern "C" int oracle_write_energy(const mom6cu_domain* dom, const mom6cu_grid* Gh, const mom6cu_vgrid* GV, mom6cu_sum_output_cs* CS,
                                   const double* u_, const double* v_, const double* h_, const double* T_, const double* S_,
                                   mom6cu_energy_out* out) {
  if (!GV->Boussinesq) return 3;
  OGrid G(dom, Gh);
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, nz = G.ke;
  const int Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  const int isr = is - (G.isd - 1), ier = ie - (G.isd - 1), jsr = js - (G.jsd - 1), jer = je - (G.jsd - 1);
  V3 u = G.U3(u_), v = G.V3_(v_), h = G.H3(h_);
  const double RZL4_T2_to_J = CS->RZL2_to_kg * (CS->L_T_to_m_s * CS->L_T_to_m_s);  // US%RZL2_to_kg*US%L_T_to_m_s**2
  const double kg_to_RZL2 = CS->kg_m3_to_R * CS->m_to_Z * (CS->m_to_L * CS->m_to_L);
  const double J_to_QRZL2 = CS->J_kg_to_Q * kg_to_RZL2;
  int rc;
  A2 areaTm = G.aH();
  for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) areaTm(i, j) = G.mask2dT(i, j) * G.areaT(i, j);
  A3 tmp1(G.isd, G.ied, G.jsd, G.jed, nz);
  for (int k = 1; k <= nz; ++k) for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i)
    tmp1(i, j, k) = h(i, j, k) * (GV->H_to_RZ * areaTm(i, j));
  std::vector<double> mass_lay(nz), vol_lay(nz), KE(nz), PE(nz + 1, 0.0), Z_0APE(nz + 1, 0.0);
  mom6cu_efp mass_EFP, salt_EFP = {}, heat_EFP = {};
  double mass_tot;
  if ((rc = oracle_reproducing_sum(dom, tmp1.p, 0, nz, isr, ier, jsr, jer, CS->RZL2_to_kg, 1, 1, &mass_tot, mass_lay.data(), &mass_EFP, nullptr)))
    return rc;
  for (int k = 0; k < nz; ++k) vol_lay[k] = (1.0 / GV->Rho0) * mass_lay[k];

  if (CS->previous_calls == 0) {  // :578-584
    CS->mass_prev_EFP = mass_EFP;
    real_to_efp(0.0, &CS->fresh_water_in_EFP);
    if (CS->use_temperature) { real_to_efp(0.0, &CS->net_salt_in_EFP); real_to_efp(0.0, &CS->net_heat_in_EFP); }
  }

  double PE_tot = 0.0;
  if (CS->do_APE_calc) {  // :642-711
    const double* DLv = CS->DL_vol_below - 1; const double* DLd = CS->DL_depth - 1; const double* DLa = CS->DL_area - 1;  // 1-based
    int* lH = CS->lH - 1;
    int lbelow = 1, li = 0; double volbelow = 0.0;
    for (int k = nz; k >= 1; --k) {
      volbelow = volbelow + vol_lay[k - 1];
      if ((volbelow >= DLv[lH[k]]) && (volbelow < DLv[lH[k] + 1])) li = lH[k];
      else {
        int labove = CS->DL_listsize;
        li = (labove + lbelow) / 2;
        while (li > lbelow) {
          if (volbelow < DLv[li]) labove = li; else lbelow = li;
          li = (labove + lbelow) / 2;
        }
        lH[k] = li;
      }
      lbelow = li;
      Z_0APE[k - 1] = DLd[li] - (volbelow - DLv[li]) / DLa[li];
    }
    Z_0APE[nz] = DLd[2];
    A3 PE_pt(G.isd, G.ied, G.jsd, G.jed, nz + 1);
    for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
      double hbelow = 0.0;
      for (int k = nz; k >= 1; --k) {
        hbelow = hbelow + h(i, j, k) * GV->H_to_Z;
        const double hint = Z_0APE[k - 1] + (hbelow - (G.bathyT(i, j) + CS->Z_ref));
        double hbot = Z_0APE[k - 1] - (G.bathyT(i, j) + CS->Z_ref);
        hbot = (hbot + std::fabs(hbot)) * 0.5;
        PE_pt(i, j, k) = (0.5 * areaTm(i, j)) * (GV->Rho0 * CS->g_prime[k - 1]) * (hint * hint - hbot * hbot);
      }
    }
    if ((rc = oracle_reproducing_sum(dom, PE_pt.p, 0, nz + 1, isr, ier, jsr, jer, RZL4_T2_to_J, 1, 1, &PE_tot, PE.data(), nullptr, nullptr)))
      return rc;
  }

  // kinetic energy :713-720
  tmp1.fill(0.0);
  for (int k = 1; k <= nz; ++k) for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i)
    tmp1(i, j, k) = (0.25 * GV->H_to_RZ * (areaTm(i, j) * h(i, j, k))) *
                    (((u(i - 1, j, k) * u(i - 1, j, k)) + (u(i, j, k) * u(i, j, k))) + ((v(i, j - 1, k) * v(i, j - 1, k)) + (v(i, j, k) * v(i, j, k))));
  double KE_tot;
  if ((rc = oracle_reproducing_sum(dom, tmp1.p, 0, nz, isr, ier, jsr, jer, RZL4_T2_to_J, 1, 1, &KE_tot, KE.data(), nullptr, nullptr))) return rc;

  EfpFlags F;
  const long long prec_error = 0x7fffffffffffffffLL;
  if (CS->use_temperature) {  // :723-744
    V3 T = G.H3(T_), S = G.H3(S_);
    A2 Temp_int = G.aH(), Salt_int = G.aH();
    for (int k = 1; k <= nz; ++k) for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
      Salt_int(i, j) = Salt_int(i, j) + S(i, j, k) * (h(i, j, k) * (GV->H_to_RZ * areaTm(i, j)));
      Temp_int(i, j) = Temp_int(i, j) + (CS->C_p * T(i, j, k)) * (h(i, j, k) * (GV->H_to_RZ * areaTm(i, j)));
    }
    double dummy;
    if ((rc = oracle_reproducing_sum(dom, Salt_int.p, 0, 1, isr, ier, jsr, jer, CS->RZL2_to_kg * CS->S_to_ppt, 1, 1, &dummy, nullptr, &salt_EFP, nullptr)))
      return rc;
    if ((rc = oracle_reproducing_sum(dom, Temp_int.p, 0, 1, isr, ier, jsr, jer, CS->RZL2_to_kg * CS->Q_to_J_kg, 1, 1, &dummy, nullptr, &heat_EFP, nullptr)))
      return rc;
    // EFP_sum_across_PEs(EFP_list, 5) :870-916 on one PE: the sum is the identity, the overflows are carried
    mom6cu_efp* list[5] = {&salt_EFP, &heat_EFP, &CS->fresh_water_in_EFP, &CS->net_salt_in_EFP, &CS->net_heat_in_EFP};
    for (int q = 0; q < 5; ++q) carry_overflow((long long*)list[q]->v, prec_error, F);
  } else {
    carry_overflow((long long*)CS->fresh_water_in_EFP.v, prec_error, F);
  }
  if (F.overflow_error) return 24;

  // maximum CFL numbers :746-768
  double max_CFL[2] = {0.0, 0.0};
  for (int k = 1; k <= nz; ++k) for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) {
    double CFL_Iarea = G.IareaT(I, j);
    if (u(I, j, k) < 0.0) CFL_Iarea = G.IareaT(I + 1, j);
    const double CFL_trans = std::fabs(u(I, j, k) * CS->dt_in_T) * (G.dy_Cu(I, j) * CFL_Iarea);
    const double CFL_lin = std::fabs(u(I, j, k) * CS->dt_in_T) * G.IdxCu(I, j);
    max_CFL[0] = fmax2(max_CFL[0], CFL_trans);
    max_CFL[1] = fmax2(max_CFL[1], CFL_lin);
  }
  for (int k = 1; k <= nz; ++k) for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) {
    double CFL_Iarea = G.IareaT(i, J);
    if (v(i, J, k) < 0.0) CFL_Iarea = G.IareaT(i, J + 1);
    const double CFL_trans = std::fabs(v(i, J, k) * CS->dt_in_T) * (G.dx_Cv(i, J) * CFL_Iarea);
    const double CFL_lin = std::fabs(v(i, J, k) * CS->dt_in_T) * G.IdyCv(i, J);
    max_CFL[0] = fmax2(max_CFL[0], CFL_trans);
    max_CFL[1] = fmax2(max_CFL[1], CFL_lin);
  }

  double Salt = 0.0, Heat = 0.0, Salt_chg = 0.0, Salt_anom = 0.0, Heat_chg = 0.0, Heat_anom = 0.0;
  mom6cu_efp t1, t2;
  if (CS->use_temperature) {  // :774-790
    Salt = kg_to_RZL2 * efp_to_real(&salt_EFP);
    Heat = J_to_QRZL2 * efp_to_real(&heat_EFP);
    if (CS->previous_calls == 0) { CS->salt_prev_EFP = salt_EFP; CS->heat_prev_EFP = heat_EFP; }
    efp_minus(&salt_EFP, &CS->salt_prev_EFP, &t1, F);
    { mom6cu_efp c = t1; Salt_chg = kg_to_RZL2 * efp_to_real(&c); t1 = c; }
    efp_minus(&t1, &CS->net_salt_in_EFP, &t2, F);
    Salt_anom = kg_to_RZL2 * efp_to_real(&t2);
    efp_minus(&heat_EFP, &CS->heat_prev_EFP, &t1, F);
    { mom6cu_efp c = t1; Heat_chg = J_to_QRZL2 * efp_to_real(&c); t1 = c; }
    efp_minus(&t1, &CS->net_heat_in_EFP, &t2, F);
    Heat_anom = J_to_QRZL2 * efp_to_real(&t2);
  }
  mom6cu_efp mass_chg_EFP, mass_anom_EFP;
  efp_minus(&mass_EFP, &CS->mass_prev_EFP, &mass_chg_EFP, F);
  efp_minus(&mass_chg_EFP, &CS->fresh_water_in_EFP, &mass_anom_EFP, F);
  const double mass_anom = kg_to_RZL2 * efp_to_real(&mass_anom_EFP);
  const double mass_chg = kg_to_RZL2 * efp_to_real(&mass_chg_EFP);
  double salin = 0.0, salin_anom = 0.0, temp = 0.0, temp_anom = 0.0;
  if (CS->use_temperature) {
    salin = Salt / mass_tot;
    salin_anom = Salt_anom / mass_tot;
    temp = Heat / (mass_tot * CS->C_p);
    temp_anom = Heat_anom / (mass_tot * CS->C_p);
  }
  const double toten = KE_tot + PE_tot;
  const double En_mass = toten / mass_tot;

  out->En_mass = En_mass; out->toten = toten; out->KE_tot = KE_tot; out->PE_tot = PE_tot; out->mass_tot = mass_tot;
  out->mass_chg = mass_chg; out->mass_anom = mass_anom; out->max_CFL[0] = max_CFL[0]; out->max_CFL[1] = max_CFL[1];
  out->Salt = Salt; out->Salt_chg = Salt_chg; out->Salt_anom = Salt_anom; out->Heat = Heat; out->Heat_chg = Heat_chg;
  out->Heat_anom = Heat_anom; out->salin = salin; out->salin_anom = salin_anom; out->temp = temp; out->temp_anom = temp_anom;
  out->ntrunc = CS->ntrunc;
  if (out->KE) for (int k = 0; k < nz; ++k) out->KE[k] = KE[k];
  if (out->mass_lay) for (int k = 0; k < nz; ++k) out->mass_lay[k] = mass_lay[k];
  if (out->PE) for (int k = 0; k <= nz; ++k) out->PE[k] = PE[k];
  if (out->Z_0APE) for (int k = 0; k <= nz; ++k) out->Z_0APE[k] = Z_0APE[k];

  if (En_mass != En_mass) return 41;  // "NaNs in total model energy forced model termination."
  CS->ntrunc = 0;  // :1010-1018
  CS->previous_calls = CS->previous_calls + 1;
  CS->mass_prev_EFP = mass_EFP; real_to_efp(0.0, &CS->fresh_water_in_EFP);
  if (CS->use_temperature) {
    CS->salt_prev_EFP = salt_EFP; real_to_efp(0.0, &CS->net_salt_in_EFP);
    CS->heat_prev_EFP = heat_EFP; real_to_efp(0.0, &CS->net_heat_in_EFP);
  }
  return 0;
}

// ---- Fortran edit descriptors, as gfortran writes them
static std::string f_ES(double x, int w, int d) {  // ESw.d
  char b[64];
  if (x != x) { std::string s(w, ' '); s.replace(w - 3, 3, "NaN"); return s; }
  snprintf(b, sizeof b, "%.*E", d, x);
  std::string s(b);
  const size_t e = s.find('E');
  int ex = atoi(s.c_str() + e + 1);
  char eb[16];
  if (ex > -100 && ex < 100) snprintf(eb, sizeof eb, "E%c%02d", ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
  else snprintf(eb, sizeof eb, "%c%03d", ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);  // the E is dropped for 3-digit exponents
  s = s.substr(0, e) + eb;
  if ((int)s.size() > w) return std::string(w, '*');
  return std::string(w - s.size(), ' ') + s;
}
static std::string f_F(double x, int w, int d) {  // Fw.d
  char b[400];
  snprintf(b, sizeof b, "%.*f", d, x);
  std::string s(b);
  if ((int)s.size() > w) {  // gfortran drops the optional leading zero before giving up
    if (s.compare(0, 2, "0.") == 0) s = s.substr(1);
    else if (s.compare(0, 3, "-0.") == 0) s = "-" + s.substr(2);
  }
  if ((int)s.size() > w) return std::string(w, '*');
  return std::string(w - s.size(), ' ') + s;
}
static std::string f_I(long long n, int w) {
  char b[32]; snprintf(b, sizeof b, "%lld", n);
  std::string s(b);
  if ((int)s.size() > w) return std::string(w, '*');
  return std::string(w - s.size(), ' ') + s;
}

// The ocean.stats line :874-902 (not date-stamped).  reday in CS%Timeunit units.
extern "C" int oracle_ocean_stats_line(const mom6cu_sum_output_cs* CS, const mom6cu_energy_out* e, int n, double reday, char* buf,
                                       size_t len) {
  std::string day_str, n_str;
  if (reday < 1.0e8) day_str = f_F(reday, 12, 3);
  else if (reday < 1.0e11) day_str = f_F(reday, 15, 3);
  else day_str = f_ES(reday, 15, 9);
  if (n < 1000000) n_str = f_I(n, 6);
  else if (n < 10000000) n_str = f_I(n, 7);
  else if (n < 100000000) n_str = f_I(n, 8);
  else n_str = f_I(n, 10);
  auto trim = [](std::string s) { while (!s.empty() && s.back() == ' ') s.pop_back(); return s; };
  const double L2 = CS->L_T_to_m_s * CS->L_T_to_m_s;
  const double SL = e->Z_0APE ? -CS->Z_to_m * e->Z_0APE[0] : -CS->Z_to_m * 0.0;
  std::string s = trim(n_str) + "," + trim(day_str) + "," + f_I(e->ntrunc, 6) + ", En " + f_ES(L2 * e->En_mass, 22, 16) + ", CFL " +
                  f_F(e->max_CFL[0], 8, 5) + ", SL " + f_ES(SL, 11, 4);
  if (CS->use_temperature)
    s += ", M " + f_ES(CS->RZL2_to_kg * e->mass_tot, 11, 5) + ", S" + f_F(e->salin, 8, 4) + ", T" + f_F(CS->C_to_degC * e->temp, 8, 4) +
         ", Me " + f_ES(e->mass_anom / e->mass_tot, 9, 2) + ", Se " + f_ES(e->salin_anom, 9, 2) + ", Te " +
         f_ES(CS->C_to_degC * e->temp_anom, 9, 2);
  else
    s += ", Mass " + f_ES(CS->RZL2_to_kg * e->mass_tot, 11, 5) + ", Me " + f_ES(e->mass_anom / e->mass_tot, 9, 2);
  if (s.size() + 1 > len) return 2;
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return 0;
}

// The init symbols the reference's C++ driver calls before its solver loop, re-implemented for the device path:
//   f_materialpoint_initBase(paths...)   /root/reference/src/grid/main.cpp:9-13,80-83 -> materialpoint.f90:68 (CLI, IO, config)
//   f_discretization_init(...)           /root/reference/src/grid/discretization_grid.h:10-13 -> discretization.f90:79-99
//   f_materialpoint_initDamask()         main.cpp:14,86 -> materialpoint.f90:138 (material, homogenization, phase init)
// plus the coordinate accessors of discretization_grid.h:14-19.  With them the reference's main.cpp links against this
// library unchanged ("Level 1" of INTEGRATION.md): initBase remembers the file names, discretization_init the material index
// of every local cell and the cell-centre coordinates (from which the local grid is recovered), initDamask parses
// material.yaml / numerics.yaml with the library's own reader (yaml_min.cpp), builds the dmk_phase tables -- the same set-up
// solver_cpp_b200/material.py does for the Python host; tests/test_gpu_consumer.py checks the two against each other --
// creates the context, initialises the per-point state on the device and binds it for the f_homogenization_* symbols.
//
// Host-side set-up the reference performs in Fortran:
//   elastic Hooke C66(T)                /root/reference/src/phase_mechanical_elastic.f90:24-104
//   phenopowerlaw parameters            /root/reference/src/phase_mechanical_plastic_phenopowerlaw.f90:72-290
//   isotropic parameters                /root/reference/src/phase_mechanical_plastic_isotropic.f90:48-148
//   dislotwin parameters                /root/reference/src/phase_mechanical_plastic_dislotwin.f90:129-452
//   numerics of the integrator          /root/reference/src/phase_mechanical.f90:277-294
//   polynomials in (T - T_ref)          /root/reference/src/polynomials.f90:65-110
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dmk_grid.h"
#include "dmk_common.cuh"
#include "yaml_min.h"

extern "C" {
int dmk_crystal_schmid_slip_ns(const char* lattice, const int* N_sl, int n_fam, const double* non_schmid, int n_coeff, int sense, double* P);
int dmk_crystal_schmid_slip(const char* lattice, const int* N_sl, int n_fam, double* P);
int dmk_crystal_schmid_twin(const char* lattice, const int* N_tw, int n_fam, double* P);
int dmk_crystal_slip_normal(const char* lattice, const int* N_sl, int n_fam, double* n);
int dmk_crystal_interaction(const char* kind, const char* lattice, const int* N_react, int nf_react, const int* N_act, int nf_act,
                            const double* values, int n_values, double* out);
int dmk_crystal_forest_projection(const char* lattice, const int* N_sl, int n_fam, const double* f_edge, double* proj);
int dmk_crystal_C66_twin(const char* lattice, const int* N_tw, int n_fam, const double* C66, double* C66_tw);
int dmk_crystal_twin_nucleation_pairs(const int* N_tw, int n_fam, int* pairs);
void dmk_qu2om(const double q[4], double om[9]);
int dmk_bind_global(dmk_ctx* c);
}

namespace {

const double T_ROOM = 293.15;            // constants.f90:12
const double K_B = 1.380649e-23;         // constants.f90:13

[[noreturn]] void bad(const std::string& what) { throw std::runtime_error(what); }

const YamlNode& need(const YamlNode& m, const std::string& key, const std::string& where) {
  const YamlNode* n = m.get(key);
  if (!n) bad(where + ": missing key '" + key + "'");
  return *n;
}
double opt_d(const YamlNode& m, const std::string& key, double dflt) { const YamlNode* n = m.get(key); return n ? n->as_double() : dflt; }
bool opt_b(const YamlNode& m, const std::string& key, bool dflt) { const YamlNode* n = m.get(key); return n ? n->as_bool() : dflt; }

// value of a (possibly temperature dependent) parameter: key, "key,T", "key,T^2" ... around T_ref (polynomials.f90:65-110)
double poly_at(const YamlNode& m, const std::string& key, double T, const std::string& where) {
  std::vector<double> coef{need(m, key, where).as_double()};
  if (m.has(key + ",T")) coef.push_back(m.get(key + ",T")->as_double());
  for (int o = 2; o <= 4; o++) {
    const std::string k = key + ",T^" + std::to_string(o);
    if (m.has(k)) { coef.resize(o, 0.0); coef.push_back(m.get(k)->as_double()); }
  }
  if (coef.size() == 1) return coef[0];
  const double x = T - need(m, "T_ref", where).as_double();
  double y = coef.back();
  for (int i = (int)coef.size() - 2; i >= 0; i--) y = y * x + coef[i];
  return y;
}

// per-family values repeated for every system of the family
std::vector<double> expand(const std::vector<double>& vals, const std::vector<int>& N, const std::string& what) {
  if (vals.size() != N.size()) bad(what + ": " + std::to_string(vals.size()) + " values for " + std::to_string(N.size()) + " families");
  std::vector<double> out;
  for (size_t f = 0; f < N.size(); f++) out.insert(out.end(), (size_t)N[f], vals[f]);
  return out;
}
std::vector<double> fam(const YamlNode& pl, const std::string& key, const std::vector<int>& N, const double* dflt, const std::string& where) {
  const YamlNode* n = pl.get(key);
  if (!n) {
    if (!dflt) bad(where + ": missing key '" + key + "'");
    return expand(std::vector<double>(N.size(), *dflt), N, key);
  }
  return expand(n->as_doubles(), N, key);
}
int total(const std::vector<int>& N) { int s = 0; for (int v : N) s += v; return s; }
void put(double* dst, const std::vector<double>& v, size_t cap, const std::string& what) {
  if (v.size() > cap) bad("phase table overflow: " + what);
  std::copy(v.begin(), v.end(), dst);
}
void crystal_check(int rc, const std::string& what) { if (rc) bad("crystallography of " + what + ": " + dmk_last_error()); }

double isotropic_mu(const double* C, const std::string& bound) {                       // crystal.f90 isotropic_mu
  if (bound == "isostrain") return (C[0] - C[1] + C[21] * 3.0) / 5.0;
  if (bound == "isostress") return 5.0 / (4.0 / (C[0] - C[1]) + 3.0 / C[21]);
  bad("isotropic_bound " + bound);
}
double isotropic_nu(const double* C, const std::string& bound) {
  double K;
  if (bound == "isostrain") { K = 0.0; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) K += C[6 * i + j]; K /= 9.0; }
  else if (bound == "isostress") {
    // 1 / sum of the upper-left 3x3 block of C^-1; cubic: that block is [[a,b,b],[b,a,b],[b,b,a]]^-1 -> sum = 3 / (a + 2 b)
    K = (C[0] + 2.0 * C[1]) / 3.0;
  } else bad("isotropic_bound " + bound);
  const double mu = isotropic_mu(C, bound);
  return (1.5 * K - mu) / (3.0 * K + mu);
}

// one `phase:` entry of material.yaml -> dmk_phase (material.py: make_phase does the same for the Python host)
void make_phase(const std::string& name, const YamlNode& cfg, double T, dmk_phase& ph) {
  memset(&ph, 0, sizeof ph);
  const std::string where = "phase " + name;
  const std::string lattice = need(cfg, "lattice", where).scalar;
  if (lattice != "cF" && lattice != "cI") bad(where + ": lattice " + lattice + " is not offered on the device path (cF, cI)");
  const YamlNode& mech = need(cfg, "mechanical", where);
  const YamlNode& el = need(mech, "elastic", where);
  if (el.has("type") && el.get("type")->scalar != "Hooke") bad(where + ": elastic type " + el.get("type")->scalar);
  const double C11 = poly_at(el, "C_11", T, where), C12 = poly_at(el, "C_12", T, where), C44 = poly_at(el, "C_44", T, where);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) ph.C66[6 * i + j] = (i == j) ? C11 : C12;      // crystal.f90:1680-1720, cubic
  for (int i = 3; i < 6; i++) ph.C66[6 * i + i] = C44;
  const YamlNode* plp = mech.get("plastic");
  const std::string kind = (plp && plp->is_map()) ? need(*plp, "type", where).scalar : "none";
  std::vector<double> state_init, atol;
  const double one = 1.0, zero = 0.0;
  const char* lat = lattice.c_str();
  if (kind == "none") {
    ph.law = DMK_LAW_NONE;
  } else if (kind == "isotropic") {
    const YamlNode& pl = *plp;
    ph.law = DMK_LAW_ISOTROPIC;
    ph.iso_xi_inf = need(pl, "xi_inf", where).as_double(); ph.iso_dot_gamma_0 = need(pl, "dot_gamma_0", where).as_double();
    ph.iso_n = need(pl, "n", where).as_double(); ph.iso_h_0 = need(pl, "h_0", where).as_double(); ph.iso_h = opt_d(pl, "h", 3.0);
    ph.iso_M = need(pl, "M", where).as_double(); ph.iso_h_ln = opt_d(pl, "h_ln", 0.0);
    ph.iso_c_1 = opt_d(pl, "c_1", 0.0); ph.iso_c_2 = opt_d(pl, "c_2", 0.0); ph.iso_c_3 = opt_d(pl, "c_3", 0.0); ph.iso_c_4 = opt_d(pl, "c_4", 0.0);
    ph.iso_a = need(pl, "a", where).as_double();
    ph.iso_dilatation = opt_b(pl, "dilatation", false) ? 1 : 0;
    state_init = {need(pl, "xi_0", where).as_double()};
    atol = {opt_d(pl, "atol_xi", 1.0)};
  } else if (kind == "phenopowerlaw") {
    const YamlNode& pl = *plp;
    ph.law = DMK_LAW_PHENOPOWERLAW;
    const std::vector<int> N_sl = pl.has("N_sl") ? pl.get("N_sl")->as_ints() : std::vector<int>();
    const std::vector<int> N_tw = pl.has("N_tw") ? pl.get("N_tw")->as_ints() : std::vector<int>();
    const int ns = total(N_sl), nt = total(N_tw);
    if (ns > DMK_MAX_SL || nt > DMK_MAX_TW) bad(where + ": too many slip / twin systems");
    ph.n_sl = ns; ph.n_tw = nt;
    std::vector<double> xi0_sl, xi0_tw;
    if (ns) {
      put(ph.dot_gamma_0_sl, fam(pl, "dot_gamma_0_sl", N_sl, &one, where), DMK_MAX_SL, "dot_gamma_0_sl");
      put(ph.n_sl_exp, fam(pl, "n_sl", N_sl, &one, where), DMK_MAX_SL, "n_sl");
      put(ph.a_sl, fam(pl, "a_sl", N_sl, &one, where), DMK_MAX_SL, "a_sl");
      put(ph.h_0_sl_sl, fam(pl, "h_0_sl-sl", N_sl, &zero, where), DMK_MAX_SL, "h_0_sl-sl");
      put(ph.c_1, fam(pl, "c_1", N_sl, &zero, where), DMK_MAX_SL, "c_1");
      put(ph.c_2, fam(pl, "c_2", N_sl, &one, where), DMK_MAX_SL, "c_2");
      xi0_sl = fam(pl, "xi_0_sl", N_sl, nullptr, where);
      put(ph.xi_inf_sl, fam(pl, "xi_inf_sl", N_sl, nullptr, where), DMK_MAX_SL, "xi_inf_sl");
      put(ph.f_sat_sl_tw, fam(pl, "f_sat_sl-tw", N_sl, &zero, where), DMK_MAX_SL, "f_sat_sl-tw");
      const std::vector<double> h = need(pl, "h_sl-sl", where).as_doubles();
      std::vector<double> M((size_t)ns * ns);
      crystal_check(dmk_crystal_interaction("sl-sl", lat, N_sl.data(), (int)N_sl.size(), N_sl.data(), (int)N_sl.size(), h.data(), (int)h.size(), M.data()), where);
      for (int a = 0; a < ns; a++) for (int b = 0; b < ns; b++) ph.h_sl_sl[a][b] = M[(size_t)a * ns + b];
      std::vector<double> P((size_t)ns * 9);
      crystal_check(dmk_crystal_schmid_slip(lat, N_sl.data(), (int)N_sl.size(), P.data()), where);
      memcpy(ph.P_sl, P.data(), P.size() * sizeof(double));
      if (lattice == "cI") {
        const std::vector<double> a_nS = pl.has("a_nonSchmid_110") ? pl.get("a_nonSchmid_110")->as_doubles() : std::vector<double>();
        const double none = 0.0;
        crystal_check(dmk_crystal_schmid_slip_ns(lat, N_sl.data(), (int)N_sl.size(), a_nS.empty() ? &none : a_nS.data(), (int)a_nS.size(), +1, P.data()), where);
        memcpy(ph.P_nS_pos, P.data(), P.size() * sizeof(double));
        crystal_check(dmk_crystal_schmid_slip_ns(lat, N_sl.data(), (int)N_sl.size(), a_nS.empty() ? &none : a_nS.data(), (int)a_nS.size(), -1, P.data()), where);
        memcpy(ph.P_nS_neg, P.data(), P.size() * sizeof(double));
      } else {
        memcpy(ph.P_nS_pos, P.data(), P.size() * sizeof(double));
        for (int a = 0; a < ns; a++) for (int k = 0; k < 9; k++) ph.P_nS_neg[a][k] = -P[(size_t)a * 9 + k];
      }
    }
    if (nt) {
      put(ph.dot_gamma_0_tw, fam(pl, "dot_gamma_0_tw", N_tw, &one, where), DMK_MAX_TW, "dot_gamma_0_tw");
      put(ph.n_tw_exp, fam(pl, "n_tw", N_tw, &one, where), DMK_MAX_TW, "n_tw");
      put(ph.c_3, fam(pl, "c_3", N_tw, &zero, where), DMK_MAX_TW, "c_3");
      put(ph.c_4, fam(pl, "c_4", N_tw, &zero, where), DMK_MAX_TW, "c_4");
      put(ph.h_0_tw_tw, fam(pl, "h_0_tw-tw", N_tw, &zero, where), DMK_MAX_TW, "h_0_tw-tw");
      xi0_tw = fam(pl, "xi_0_tw", N_tw, nullptr, where);
      for (int a = 0; a < nt; a++) ph.gamma_char_tw[a] = 0.5 * std::sqrt(2.0);                           // crystal.f90:466-473 (cF, cI)
      const std::vector<double> h = need(pl, "h_tw-tw", where).as_doubles();
      std::vector<double> M((size_t)nt * nt);
      crystal_check(dmk_crystal_interaction("tw-tw", lat, N_tw.data(), (int)N_tw.size(), N_tw.data(), (int)N_tw.size(), h.data(), (int)h.size(), M.data()), where);
      for (int a = 0; a < nt; a++) for (int b = 0; b < nt; b++) ph.h_tw_tw[a][b] = M[(size_t)a * nt + b];
      std::vector<double> P((size_t)nt * 9);
      crystal_check(dmk_crystal_schmid_twin(lat, N_tw.data(), (int)N_tw.size(), P.data()), where);
      memcpy(ph.P_tw, P.data(), P.size() * sizeof(double));
    }
    if (ns && nt) {
      put(ph.h_0_tw_sl, fam(pl, "h_0_tw-sl", N_tw, &zero, where), DMK_MAX_TW, "h_0_tw-sl");
      const std::vector<double> hst = need(pl, "h_sl-tw", where).as_doubles(), hts = need(pl, "h_tw-sl", where).as_doubles();
      std::vector<double> M((size_t)ns * nt);
      crystal_check(dmk_crystal_interaction("sl-tw", lat, N_sl.data(), (int)N_sl.size(), N_tw.data(), (int)N_tw.size(), hst.data(), (int)hst.size(), M.data()), where);
      for (int a = 0; a < ns; a++) for (int b = 0; b < nt; b++) ph.h_sl_tw[a][b] = M[(size_t)a * nt + b];
      crystal_check(dmk_crystal_interaction("tw-sl", lat, N_tw.data(), (int)N_tw.size(), N_sl.data(), (int)N_sl.size(), hts.data(), (int)hts.size(), M.data()), where);
      for (int a = 0; a < nt; a++) for (int b = 0; b < ns; b++) ph.h_tw_sl[a][b] = M[(size_t)a * ns + b];
    }
    state_init = xi0_sl; state_init.insert(state_init.end(), xi0_tw.begin(), xi0_tw.end()); state_init.resize(2 * (size_t)(ns + nt), 0.0);
    atol.assign((size_t)(ns + nt), opt_d(pl, "atol_xi", 1.0)); atol.resize(2 * (size_t)(ns + nt), opt_d(pl, "atol_gamma", 1.0e-6));
  } else if (kind == "kinehardening") {                // plastic_kinehardening_init, phase_mechanical_plastic_kinehardening.f90:72-252
    const YamlNode& pl = *plp;
    ph.law = DMK_LAW_KINEHARDENING;
    const std::vector<int> N_sl = pl.has("N_sl") ? pl.get("N_sl")->as_ints() : std::vector<int>();
    const int ns = total(N_sl);
    if (ns < 1 || 6 * ns > DMK_MAX_STATE) bad(where + ": kinehardening needs 1.." + std::to_string(DMK_MAX_STATE / 6) + " slip systems on the device path");
    ph.n_sl = ns; ph.n_tw = 0;
    std::vector<double> P((size_t)ns * 9);
    crystal_check(dmk_crystal_schmid_slip(lat, N_sl.data(), (int)N_sl.size(), P.data()), where);
    memcpy(ph.P_sl, P.data(), P.size() * sizeof(double));
    if (lattice == "cI") {
      const std::vector<double> a_nS = pl.has("a_nonSchmid_110") ? pl.get("a_nonSchmid_110")->as_doubles() : std::vector<double>();
      const double none = 0.0;
      crystal_check(dmk_crystal_schmid_slip_ns(lat, N_sl.data(), (int)N_sl.size(), a_nS.empty() ? &none : a_nS.data(), (int)a_nS.size(), +1, P.data()), where);
      memcpy(ph.P_nS_pos, P.data(), P.size() * sizeof(double));
      crystal_check(dmk_crystal_schmid_slip_ns(lat, N_sl.data(), (int)N_sl.size(), a_nS.empty() ? &none : a_nS.data(), (int)a_nS.size(), -1, P.data()), where);
      memcpy(ph.P_nS_neg, P.data(), P.size() * sizeof(double));
    } else {
      memcpy(ph.P_nS_pos, P.data(), P.size() * sizeof(double));
      for (int a = 0; a < ns; a++) for (int k = 0; k < 9; k++) ph.P_nS_neg[a][k] = -P[(size_t)a * 9 + k];
    }
    const std::vector<double> h = need(pl, "h_sl-sl", where).as_doubles();
    std::vector<double> M((size_t)ns * ns);
    crystal_check(dmk_crystal_interaction("sl-sl", lat, N_sl.data(), (int)N_sl.size(), N_sl.data(), (int)N_sl.size(), h.data(), (int)h.size(), M.data()), where);
    for (int a = 0; a < ns; a++) for (int b = 0; b < ns; b++) ph.h_sl_sl[a][b] = M[(size_t)a * ns + b];
    put(ph.dot_gamma_0_sl, fam(pl, "dot_gamma_0", N_sl, nullptr, where), DMK_MAX_SL, "dot_gamma_0");
    put(ph.n_sl_exp, fam(pl, "n", N_sl, nullptr, where), DMK_MAX_SL, "n");
    put(ph.xi_inf_sl, fam(pl, "xi_inf", N_sl, nullptr, where), DMK_MAX_SL, "xi_inf");
    put(ph.kh_chi_inf, fam(pl, "chi_inf", N_sl, nullptr, where), DMK_MAX_SL, "chi_inf");
    put(ph.kh_h_0_xi, fam(pl, "h_0_xi", N_sl, nullptr, where), DMK_MAX_SL, "h_0_xi");
    put(ph.kh_h_0_chi, fam(pl, "h_0_chi", N_sl, nullptr, where), DMK_MAX_SL, "h_0_chi");
    put(ph.kh_h_inf_xi, fam(pl, "h_inf_xi", N_sl, nullptr, where), DMK_MAX_SL, "h_inf_xi");
    put(ph.kh_h_inf_chi, fam(pl, "h_inf_chi", N_sl, nullptr, where), DMK_MAX_SL, "h_inf_chi");
    state_init = fam(pl, "xi_0", N_sl, nullptr, where);
    for (int a = 0; a < ns; a++)
      if (!(ph.dot_gamma_0_sl[a] > 0.0 && ph.n_sl_exp[a] > 0.0 && state_init[a] > 0.0 && ph.xi_inf_sl[a] > 0.0 && ph.kh_chi_inf[a] > 0.0))
        bad(where + ": kinehardening: dot_gamma_0, n, xi_0, xi_inf, chi_inf must be positive");        // :164-169
    state_init.resize(6 * (size_t)ns, 0.0);
    atol.assign(2 * (size_t)ns, opt_d(pl, "atol_xi", 1.0)); atol.resize(3 * (size_t)ns, opt_d(pl, "atol_gamma", 1.0e-6)); atol.resize(6 * (size_t)ns, 0.0);
  } else if (kind == "dislotungsten") {                // plastic_dislotungsten_init, phase_mechanical_plastic_dislotungsten.f90:79-268
    const YamlNode& pl = *plp;
    ph.law = DMK_LAW_DISLOTUNGSTEN;
    const std::vector<int> N_sl = pl.has("N_sl") ? pl.get("N_sl")->as_ints() : std::vector<int>();
    const int ns = total(N_sl);
    if (ns < 1 || ns > DMK_MAX_SL) bad(where + ": dislotungsten needs 1.." + std::to_string(DMK_MAX_SL) + " slip systems");
    ph.n_sl = ns; ph.n_tw = 0;
    const std::string bound = pl.has("isotropic_bound") ? pl.get("isotropic_bound")->scalar : "isostrain";
    ph.dt_mu = isotropic_mu(ph.C66, bound); ph.dt_nu = isotropic_nu(ph.C66, bound);
    ph.dt_kB_T = K_B * T;
    std::vector<double> P((size_t)ns * 9);
    crystal_check(dmk_crystal_schmid_slip(lat, N_sl.data(), (int)N_sl.size(), P.data()), where);
    memcpy(ph.P_sl, P.data(), P.size() * sizeof(double));
    if (lattice == "cI") {
      const std::vector<double> a_nS = pl.has("a_nonSchmid_110") ? pl.get("a_nonSchmid_110")->as_doubles() : std::vector<double>();
      const double none = 0.0;
      crystal_check(dmk_crystal_schmid_slip_ns(lat, N_sl.data(), (int)N_sl.size(), a_nS.empty() ? &none : a_nS.data(), (int)a_nS.size(), +1, P.data()), where);
      memcpy(ph.P_nS_pos, P.data(), P.size() * sizeof(double));
      crystal_check(dmk_crystal_schmid_slip_ns(lat, N_sl.data(), (int)N_sl.size(), a_nS.empty() ? &none : a_nS.data(), (int)a_nS.size(), -1, P.data()), where);
      memcpy(ph.P_nS_neg, P.data(), P.size() * sizeof(double));
    } else {
      memcpy(ph.P_nS_pos, P.data(), P.size() * sizeof(double));
      for (int a = 0; a < ns; a++) for (int k = 0; k < 9; k++) ph.P_nS_neg[a][k] = -P[(size_t)a * 9 + k];
    }
    ph.dt_omit_dipoles = opt_b(pl, "no_dipole_formation", false) ? 1 : 0;
    ph.dt_D = need(pl, "D", where).as_double(); ph.tu_D_0 = need(pl, "D_0", where).as_double(); ph.dt_Q_cl = need(pl, "Q_cl", where).as_double();
    const double half = 0.5;
    const std::vector<double> f_edge = fam(pl, "f_edge", N_sl, &half, where);
    const std::vector<double> rho_mob_0 = fam(pl, "rho_mob_0", N_sl, nullptr, where), rho_dip_0 = fam(pl, "rho_dip_0", N_sl, nullptr, where);
    const std::vector<double> b_sl = fam(pl, "b_sl", N_sl, nullptr, where);
    put(ph.dt_b_sl, b_sl, DMK_MAX_SL, "b_sl"); put(ph.dt_Q_sl, fam(pl, "Q_s", N_sl, nullptr, where), DMK_MAX_SL, "Q_s");
    put(ph.dt_i_sl, fam(pl, "i_sl", N_sl, nullptr, where), DMK_MAX_SL, "i_sl"); put(ph.dt_tau_0, fam(pl, "tau_Peierls", N_sl, nullptr, where), DMK_MAX_SL, "tau_Peierls");
    put(ph.dt_p, fam(pl, "p_sl", N_sl, nullptr, where), DMK_MAX_SL, "p_sl"); put(ph.dt_q, fam(pl, "q_sl", N_sl, nullptr, where), DMK_MAX_SL, "q_sl");
    put(ph.dt_B, fam(pl, "B", N_sl, nullptr, where), DMK_MAX_SL, "B");
    put(ph.tu_h, fam(pl, "h", N_sl, nullptr, where), DMK_MAX_SL, "h"); put(ph.tu_w, fam(pl, "w", N_sl, nullptr, where), DMK_MAX_SL, "w");
    put(ph.tu_omega, fam(pl, "omega", N_sl, nullptr, where), DMK_MAX_SL, "omega");
    const double D_a = need(pl, "D_a", where).as_double(), f_at = need(pl, "f_at", where).as_double();
    for (int a = 0; a < ns; a++) { ph.dt_d_caron[a] = b_sl[a] * D_a; ph.tu_f_at[a] = b_sl[a] * b_sl[a] * b_sl[a] * f_at; }
    {
      const std::vector<double> h = need(pl, "h_sl-sl", where).as_doubles();
      std::vector<double> M((size_t)ns * ns);
      crystal_check(dmk_crystal_interaction("sl-sl", lat, N_sl.data(), (int)N_sl.size(), N_sl.data(), (int)N_sl.size(), h.data(), (int)h.size(), M.data()), where);
      for (int a = 0; a < ns; a++) for (int b = 0; b < ns; b++) ph.h_sl_sl[a][b] = M[(size_t)a * ns + b];
      crystal_check(dmk_crystal_forest_projection(lat, N_sl.data(), (int)N_sl.size(), f_edge.data(), M.data()), where);
      for (int a = 0; a < ns; a++) for (int b = 0; b < ns; b++) ph.dt_forest[a][b] = M[(size_t)a * ns + b];
    }
    bool okp = ph.tu_D_0 >= 0.0 && ph.dt_Q_cl > 0.0 && D_a >= 0.0 && f_at > 0.0;                         // :205-215
    for (int a = 0; a < ns; a++) okp = okp && rho_mob_0[a] >= 0.0 && rho_dip_0[a] >= 0.0 && b_sl[a] > 0.0 && ph.dt_Q_sl[a] > 0.0 && ph.dt_tau_0[a] >= 0.0 && ph.dt_B[a] >= 0.0;
    if (!okp) bad(where + ": dislotungsten: parameter out of range");
    state_init = rho_mob_0; state_init.insert(state_init.end(), rho_dip_0.begin(), rho_dip_0.end()); state_init.resize(3 * (size_t)ns, 0.0);
    atol.assign(2 * (size_t)ns, opt_d(pl, "atol_rho", 1.0)); atol.resize(3 * (size_t)ns, opt_d(pl, "atol_gamma", 1.0e-6));
  } else if (kind == "dislotwin") {
    const YamlNode& pl = *plp;
    ph.law = DMK_LAW_DISLOTWIN;
    const std::vector<int> N_sl = pl.has("N_sl") ? pl.get("N_sl")->as_ints() : std::vector<int>();
    const std::vector<int> N_tw = pl.has("N_tw") ? pl.get("N_tw")->as_ints() : std::vector<int>();
    if (pl.has("N_tr") && total(pl.get("N_tr")->as_ints()) > 0) bad(where + ": dislotwin transformation systems (N_tr) are not offered on the device path");
    const int ns = total(N_sl), nt = total(N_tw);
    if (ns == 0) bad(where + ": dislotwin without slip systems");
    if (ns > DMK_MAX_SL || nt > DMK_MAX_TW) bad(where + ": too many slip / twin systems");
    ph.n_sl = ns; ph.n_tw = nt;
    const std::string bound = pl.has("isotropic_bound") ? pl.get("isotropic_bound")->scalar : "isostrain";
    ph.dt_mu = isotropic_mu(ph.C66, bound); ph.dt_nu = isotropic_nu(ph.C66, bound);
    ph.dt_kB_T = K_B * T;
    std::vector<double> P((size_t)ns * 9), nrm((size_t)ns * 3);
    crystal_check(dmk_crystal_schmid_slip(lat, N_sl.data(), (int)N_sl.size(), P.data()), where);
    memcpy(ph.P_sl, P.data(), P.size() * sizeof(double));
    crystal_check(dmk_crystal_slip_normal(lat, N_sl.data(), (int)N_sl.size(), nrm.data()), where);
    memcpy(ph.dt_n0_sl, nrm.data(), nrm.size() * sizeof(double));
    ph.dt_extended = opt_b(pl, "extend_dislocations", false) ? 1 : 0;
    ph.dt_omit_dipoles = opt_b(pl, "omit_dipoles", false) ? 1 : 0;
    ph.dt_Q_cl = need(pl, "Q_cl", where).as_double();
    const double half = 0.5;
    const std::vector<double> f_edge = fam(pl, "f_edge", N_sl, &half, where);
    const std::vector<double> rho_mob_0 = fam(pl, "rho_mob_0", N_sl, nullptr, where), rho_dip_0 = fam(pl, "rho_dip_0", N_sl, nullptr, where);
    const std::vector<double> b_sl = fam(pl, "b_sl", N_sl, nullptr, where);
    put(ph.dt_v_0, fam(pl, "v_0", N_sl, nullptr, where), DMK_MAX_SL, "v_0"); put(ph.dt_b_sl, b_sl, DMK_MAX_SL, "b_sl");
    put(ph.dt_Q_sl, fam(pl, "Q_sl", N_sl, nullptr, where), DMK_MAX_SL, "Q_sl"); put(ph.dt_i_sl, fam(pl, "i_sl", N_sl, nullptr, where), DMK_MAX_SL, "i_sl");
    put(ph.dt_p, fam(pl, "p_sl", N_sl, nullptr, where), DMK_MAX_SL, "p_sl"); put(ph.dt_q, fam(pl, "q_sl", N_sl, nullptr, where), DMK_MAX_SL, "q_sl");
    put(ph.dt_tau_0, fam(pl, "tau_0", N_sl, nullptr, where), DMK_MAX_SL, "tau_0"); put(ph.dt_B, fam(pl, "B", N_sl, &zero, where), DMK_MAX_SL, "B");
    const double D_a = need(pl, "D_a", where).as_double();
    for (int a = 0; a < ns; a++) ph.dt_d_caron[a] = b_sl[a] * D_a;
    {
      const std::vector<double> h = need(pl, "h_sl-sl", where).as_doubles();
      std::vector<double> M((size_t)ns * ns);
      crystal_check(dmk_crystal_interaction("sl-sl", lat, N_sl.data(), (int)N_sl.size(), N_sl.data(), (int)N_sl.size(), h.data(), (int)h.size(), M.data()), where);
      for (int a = 0; a < ns; a++) for (int b = 0; b < ns; b++) ph.h_sl_sl[a][b] = M[(size_t)a * ns + b];
      crystal_check(dmk_crystal_forest_projection(lat, N_sl.data(), (int)N_sl.size(), f_edge.data(), M.data()), where);
      for (int a = 0; a < ns; a++) for (int b = 0; b < ns; b++) ph.dt_forest[a][b] = M[(size_t)a * ns + b];
    }
    const bool fcc_nucleation = lattice == "cF" && N_sl[0] == 12;
    ph.dt_omega = opt_d(pl, "omega", 1000.0) * (lattice == "cF" ? 12.0 : 8.0);
    if (nt) {
      if (!fcc_nucleation || N_tw.size() != 1) bad(where + ": twinning in dislotwin needs cF with the 12 {111} slip systems");
      std::vector<double> Pt((size_t)nt * 9);
      crystal_check(dmk_crystal_schmid_twin(lat, N_tw.data(), (int)N_tw.size(), Pt.data()), where);
      memcpy(ph.P_tw, Pt.data(), Pt.size() * sizeof(double));
      for (int a = 0; a < nt; a++) ph.gamma_char_tw[a] = 0.5 * std::sqrt(2.0);
      ph.dt_L_tw = need(pl, "L_tw", where).as_double(); ph.dt_i_tw = need(pl, "i_tw", where).as_double();
      put(ph.dt_b_tw, fam(pl, "b_tw", N_tw, nullptr, where), DMK_MAX_TW, "b_tw"); put(ph.dt_t_tw, fam(pl, "t_tw", N_tw, nullptr, where), DMK_MAX_TW, "t_tw");
      put(ph.dt_r, fam(pl, "p_tw", N_tw, nullptr, where), DMK_MAX_TW, "p_tw");
      const std::vector<double> htt = need(pl, "h_tw-tw", where).as_doubles(), hst = need(pl, "h_sl-tw", where).as_doubles();
      std::vector<double> M((size_t)std::max(ns, nt) * nt);
      crystal_check(dmk_crystal_interaction("tw-tw", lat, N_tw.data(), (int)N_tw.size(), N_tw.data(), (int)N_tw.size(), htt.data(), (int)htt.size(), M.data()), where);
      for (int a = 0; a < nt; a++) for (int b = 0; b < nt; b++) ph.h_tw_tw[a][b] = M[(size_t)a * nt + b];
      crystal_check(dmk_crystal_interaction("sl-tw", lat, N_sl.data(), (int)N_sl.size(), N_tw.data(), (int)N_tw.size(), hst.data(), (int)hst.size(), M.data()), where);
      for (int a = 0; a < ns; a++) for (int b = 0; b < nt; b++) ph.h_sl_tw[a][b] = M[(size_t)a * nt + b];
      crystal_check(dmk_crystal_twin_nucleation_pairs(N_tw.data(), (int)N_tw.size(), &ph.dt_pair[0][0]), where);
      std::vector<double> Ct((size_t)nt * 36);
      crystal_check(dmk_crystal_C66_twin(lat, N_tw.data(), (int)N_tw.size(), ph.C66, Ct.data()), where);
      memcpy(ph.dt_C66_tw, Ct.data(), Ct.size() * sizeof(double));
      ph.dt_x_c = need(pl, "x_c", where).as_double(); ph.dt_V_cs = need(pl, "V_cs", where).as_double();
    }
    if (nt || ph.dt_extended) ph.dt_Gamma_sf = poly_at(pl, "Gamma_sf", T, where);
    ph.dt_gamma_0_sb = opt_d(pl, "gamma_0_sb", 0.0);
    ph.dt_tau_sb = ph.dt_E_sb = ph.dt_p_sb = ph.dt_q_sb = 1.0;
    if (ph.dt_gamma_0_sb > 0.0) {
      ph.dt_tau_sb = need(pl, "tau_sb", where).as_double(); ph.dt_E_sb = need(pl, "Q_sb", where).as_double();
      ph.dt_p_sb = need(pl, "p_sb", where).as_double(); ph.dt_q_sb = need(pl, "q_sb", where).as_double();
    }
    ph.dt_D = need(pl, "D", where).as_double();
    state_init = rho_mob_0; state_init.insert(state_init.end(), rho_dip_0.begin(), rho_dip_0.end());
    state_init.resize(3 * (size_t)ns + nt, 0.0);
    atol.assign(2 * (size_t)ns, opt_d(pl, "atol_rho", 1.0)); atol.resize(3 * (size_t)ns, opt_d(pl, "atol_gamma", 1.0e-6));
    atol.resize(3 * (size_t)ns + nt, opt_d(pl, "atol_f_tw", 1.0e-6));
  } else {
    bad(where + ": plastic law " + kind + " is not offered on the device path");
  }
  if (state_init.size() > (size_t)DMK_MAX_STATE) bad(where + ": too many state variables");
  ph.n_state = (int)state_init.size();
  std::copy(state_init.begin(), state_init.end(), ph.state_init);
  std::copy(atol.begin(), atol.end(), ph.atol);
}

// numerics.yaml -> dmk_numerics; keys as Config::parse_numerics (config.cpp:383-402) and phase_mechanical.f90:277-294
void default_numerics(dmk_numerics& n) {
  n.r_cutback_min = 1.0e-3; n.r_cutback = 0.25; n.r_increase = 1.5; n.eps_rel_state = 1.0e-6; n.N_iter_state_max = 20;
  n.N_iter_Lp_max = 40; n.r_linesearch_Lp = 0.5; n.eps_rel_Lp = 1.0e-6; n.eps_abs_Lp = 1.0e-8; n.f_update_jacobi_Lp = 1;
  n.integrator_state = 0; n.itmin = 1; n.itmax = 100; n.update_gamma = 0;
  n.eps_abs_div = 1.0e-4; n.eps_rel_div = 5.0e-4; n.eps_abs_P = 1.0e3; n.eps_rel_P = 1.0e-3;
}
void read_numerics(const YamlNode& root, dmk_numerics& n, int* derivative, int* divcorr) {
  if (const YamlNode* solver = root.get("solver")) if (const YamlNode* grid = solver->get("grid")) {
    if (const YamlNode* m = grid->get("mechanical")) {
      // a key the parser does not know must not be dropped silently (the Python host refuses it as well)
      static const char* known[] = {"N_iter_min", "N_iter_max", "eps_abs_div(P)", "eps_rel_div(P)", "eps_abs_P", "eps_rel_P", "update_gamma",
                                    "itmin", "itmax", "eps_abs_div", "eps_rel_div"};                 // + the struct's own names, as the Python host
      if (m->is_map()) for (const auto& kv : m->map) {
        bool ok = false;
        for (const char* k : known) ok = ok || kv.first == k;
        if (!ok) bad("numerics: unknown key '" + kv.first + "' under solver/grid/mechanical");
      }
      if (m->has("N_iter_min")) n.itmin = (int)m->get("N_iter_min")->as_int();
      if (m->has("N_iter_max")) n.itmax = (int)m->get("N_iter_max")->as_int();
      if (m->has("eps_abs_div(P)")) n.eps_abs_div = m->get("eps_abs_div(P)")->as_double();
      if (m->has("eps_rel_div(P)")) n.eps_rel_div = m->get("eps_rel_div(P)")->as_double();
      if (m->has("eps_abs_P")) n.eps_abs_P = m->get("eps_abs_P")->as_double();
      if (m->has("eps_rel_P")) n.eps_rel_P = m->get("eps_rel_P")->as_double();
      if (m->has("update_gamma")) n.update_gamma = m->get("update_gamma")->as_bool() ? 1 : 0;
      if (m->has("itmin")) n.itmin = (int)m->get("itmin")->as_int();
      if (m->has("itmax")) n.itmax = (int)m->get("itmax")->as_int();
      if (m->has("eps_abs_div")) n.eps_abs_div = m->get("eps_abs_div")->as_double();
      if (m->has("eps_rel_div")) n.eps_rel_div = m->get("eps_rel_div")->as_double();
    }
    if (const YamlNode* f = grid->get("FFT")) {
      if (f->has("derivative")) {
        const std::string d = f->get("derivative")->scalar;
        *derivative = d == "continuous" ? DMK_DERIVATIVE_CONTINUOUS : d == "central_difference" ? DMK_DERIVATIVE_CENTRAL_DIFF
                    : d == "FWBW_difference" ? DMK_DERIVATIVE_FWBW_DIFF : -1;
        if (*derivative < 0) bad("numerics: FFT.derivative must be continuous, central_difference or FWBW_difference");
      }
      if (f->has("divergence_correction")) {
        const std::string d = f->get("divergence_correction")->scalar;
        *divcorr = d == "none" ? DMK_DIVCORR_NONE : d == "size" ? DMK_DIVCORR_SIZE : (d == "size+grid" || d == "grid+size") ? DMK_DIVCORR_SIZE_GRID : -1;
        if (*divcorr < 0) bad("numerics: FFT.divergence_correction must be none, size or size+grid");
      }
    }
  }
  if (const YamlNode* phase = root.get("phase")) if (const YamlNode* m = phase->get("mechanical")) {
    // phase_mechanical.f90:277-294; keys the integrator does not know are refused (the Python host does the same); the Li keys of
    // `eigen` are accepted and unused: no law with Li != 0 is offered on the device path
    auto refuse_unknown = [](const YamlNode* node, std::initializer_list<const char*> known, const std::string& where) {
      if (!node || !node->is_map()) return;
      for (const auto& kv : node->map) {
        bool ok = false;
        for (const char* k : known) ok = ok || kv.first == k;
        if (!ok) bad("numerics: unknown key '" + kv.first + "' under " + where);
      }
    };
    refuse_unknown(m, {"r_cutback_min", "r_cutback", "r_increase", "eps_rel_state", "N_iter_state_max", "plastic", "eigen"}, "phase/mechanical");
    refuse_unknown(m->get("plastic"), {"N_iter_Lp_max", "r_linesearch_Lp", "eps_rel_Lp", "eps_abs_Lp", "f_update_jacobi_Lp", "integrator_state"}, "phase/mechanical/plastic");
    refuse_unknown(m->get("eigen"), {"N_iter_Li_max", "r_linesearch_Li", "eps_rel_Li", "eps_abs_Li", "f_update_jacobi_Li"}, "phase/mechanical/eigen");
    if (m->has("r_cutback_min")) n.r_cutback_min = m->get("r_cutback_min")->as_double();
    if (m->has("r_cutback")) n.r_cutback = m->get("r_cutback")->as_double();
    if (m->has("r_increase")) n.r_increase = m->get("r_increase")->as_double();
    if (m->has("eps_rel_state")) n.eps_rel_state = m->get("eps_rel_state")->as_double();
    if (m->has("N_iter_state_max")) n.N_iter_state_max = (int)m->get("N_iter_state_max")->as_int();
    if (const YamlNode* p = m->get("plastic")) {
      if (p->has("N_iter_Lp_max")) n.N_iter_Lp_max = (int)p->get("N_iter_Lp_max")->as_int();
      if (p->has("r_linesearch_Lp")) n.r_linesearch_Lp = p->get("r_linesearch_Lp")->as_double();
      if (p->has("eps_rel_Lp")) n.eps_rel_Lp = p->get("eps_rel_Lp")->as_double();
      if (p->has("eps_abs_Lp")) n.eps_abs_Lp = p->get("eps_abs_Lp")->as_double();
      if (p->has("f_update_jacobi_Lp")) n.f_update_jacobi_Lp = (int)p->get("f_update_jacobi_Lp")->as_int();
      if (p->has("integrator_state")) {
        static const char* names[5] = {"FPI", "Euler", "AdaptiveEuler", "RK4", "RKCK45"};
        const std::string v = p->get("integrator_state")->scalar;
        int id = -1;
        for (int k = 0; k < 5; k++) if (v == names[k]) id = k;
        if (id < 0) bad("numerics: unknown state integrator " + v);
        n.integrator_state = id;
      }
    }
  }
}

// what the init calls leave for each other
struct InitState {
  std::string material_path, load_path, grid_path, numerics_path;
  std::vector<int> material_at;                 // 0-based material index of every local cell, x fastest
  std::vector<double> ip_coords, node_coords;   // 3 x n, as handed in
  int shared_nodes_begin = 0;
  dmk_ctx* ctx = nullptr;
} g_init;

bool file_exists(const std::string& p) { std::ifstream f(p); return f.good(); }

// local grid from the cell-centre coordinates (x fastest, then y, then z; a regular grid): cells and size of the slab
void grid_from_ip_coords(const std::vector<double>& x, long long n, int cells[3], double size[3]) {
  auto coord = [&](long long i, int d) { return x[(size_t)(3 * i + d)]; };
  long long nx = n, ny = 1;
  for (long long i = 1; i < n; i++) if (coord(i, 1) != coord(0, 1) || coord(i, 2) != coord(0, 2)) { nx = i; break; }
  const long long rows = n / nx;
  ny = rows;
  for (long long r = 1; r < rows; r++) if (coord(r * nx, 2) != coord(0, 2)) { ny = r; break; }
  const long long nz = rows / ny;
  if (nx * ny * nz != n) bad("f_discretization_init: the cell-centre coordinates do not describe a regular grid (x fastest)");
  cells[0] = (int)nx; cells[1] = (int)ny; cells[2] = (int)nz;
  const double dx = nx > 1 ? coord(1, 0) - coord(0, 0) : 2.0 * coord(0, 0);
  const double dy = ny > 1 ? coord(nx, 1) - coord(0, 1) : 2.0 * coord(0, 1);
  const double dz = nz > 1 ? coord(nx * ny, 2) - coord(0, 2) : (dx > 0 ? dx : 1.0);
  size[0] = dx * nx; size[1] = dy * ny; size[2] = dz * nz;
  for (int d = 0; d < 3; d++) if (!(size[d] > 0.0)) bad("f_discretization_init: non-positive grid spacing");
}

}  // namespace

// material.yaml (parsed) -> phase tables and, per material entry, its constituents: phase index, orientation matrix (row-major
// 3x3, Fp0 = O) and volume fraction.  Entry (m, k) of the per-material tables sits at m * max_constituents + k.  Checks of
// material.f90:145,161: the number of constituents is the homogenization's N_constituents, their fractions sum to 1; mechanical
// homogenisation types offered: pass (one constituent), isostrain.
extern "C" int dmk_material_constituents_from_yaml(const char* material_yaml_path, double T, dmk_phase* phases, int max_phases, int* n_phases,
                                                   int max_constituents, int* n_constituents_of_material, int* phase_of_material,
                                                   double* O_of_material, double* v_of_material, int max_materials, int* n_materials) {
  try {
    if (!material_yaml_path || !phases || !n_phases || !n_constituents_of_material || !phase_of_material || !O_of_material || !v_of_material || !n_materials || max_constituents < 1) bad("dmk_material_constituents_from_yaml: null argument");
    auto root = yaml_parse_file(material_yaml_path);
    const YamlNode& ph = need(*root, "phase", "material.yaml");
    if (!ph.is_map() || ph.map.empty()) bad("material.yaml: `phase` must be a non-empty mapping");
    if ((int)ph.map.size() > max_phases) bad("material.yaml: too many phases");
    for (size_t i = 0; i < ph.map.size(); i++) make_phase(ph.map[i].first, *ph.map[i].second, T, phases[i]);
    *n_phases = (int)ph.map.size();
    const YamlNode& mats = need(*root, "material", "material.yaml");
    if (!mats.is_seq()) bad("material.yaml: `material` must be a sequence");
    if ((int)mats.seq.size() > max_materials) bad("material.yaml: too many materials for the caller's buffers");
    const YamlNode* homs = root->get("homogenization");
    for (size_t m = 0; m < mats.seq.size(); m++) {
      const YamlNode& cons = need(*mats.seq[m], "constituents", "material entry");
      if (!cons.is_seq() || cons.seq.empty()) bad("material entry without constituents");
      const int nc = (int)cons.seq.size();
      if (nc > max_constituents) bad("material entry " + std::to_string(m) + ": " + std::to_string(nc) + " constituents, the caller's buffers hold " + std::to_string(max_constituents));
      const YamlNode* hname = mats.seq[m]->get("homogenization");
      const YamlNode* hom = (homs && homs->is_map() && hname) ? homs->get(hname->scalar) : nullptr;
      if (hom) {
        const YamlNode* ncn = hom->get("N_constituents");
        if (ncn && ncn->as_int() != nc) bad("material entry " + std::to_string(m) + ": " + std::to_string(nc) + " constituents, its homogenization asks for " + ncn->scalar);
        const YamlNode* mech = hom->get("mechanical");
        const YamlNode* type = mech ? mech->get("type") : nullptr;
        const std::string t = type ? type->scalar : std::string("pass");
        if ((t != "pass" && t != "isostrain") || (t == "pass" && nc != 1)) bad("mechanical homogenisation " + t + " with " + std::to_string(nc) + " constituents is not offered (pass, isostrain)");
      } else if (nc != 1) bad("material entry " + std::to_string(m) + ": several constituents without a homogenization entry");
      n_constituents_of_material[m] = nc;
      double sv = 0.0;
      for (int k = 0; k < nc; k++) {
        const size_t at = m * (size_t)max_constituents + (size_t)k;
        const std::string name = need(*cons.seq[(size_t)k], "phase", "constituent").scalar;
        int idx = -1;
        for (size_t i = 0; i < ph.map.size(); i++) if (ph.map[i].first == name) idx = (int)i;
        if (idx < 0) bad("material entry refers to unknown phase " + name);
        phase_of_material[at] = idx;
        const std::vector<double> q = need(*cons.seq[(size_t)k], "O", "constituent").as_doubles();
        if (q.size() != 4) bad("constituent orientation O needs 4 quaternion components");
        dmk_qu2om(q.data(), O_of_material + 9 * at);
        const YamlNode* vn = cons.seq[(size_t)k]->get("v");
        v_of_material[at] = vn ? vn->as_double() : 1.0;
        sv += v_of_material[at];
      }
      if (fabs(sv - 1.0) > 1.0e-9) bad("material entry " + std::to_string(m) + ": volume fractions of the constituents do not sum to 1");
    }
    *n_materials = (int)mats.seq.size();
    return DMK_OK;
  } catch (const std::exception& e) {
    dmk_set_error("%s", e.what());
    return DMK_ERR_ARG;
  }
}

// the single-constituent form: per-material phase index / orientation matrix
extern "C" int dmk_material_from_yaml(const char* material_yaml_path, double T, dmk_phase* phases, int max_phases, int* n_phases,
                                      int* phase_of_material, double* O_of_material, int max_materials, int* n_materials) {
  if (max_materials < 0) { dmk_set_error("dmk_material_from_yaml: bad argument"); return DMK_ERR_ARG; }
  std::vector<int> nc((size_t)max_materials + 1);
  std::vector<double> v((size_t)max_materials + 1);
  // one slot per material: an entry with more constituents is refused by the general reader
  int rc = dmk_material_constituents_from_yaml(material_yaml_path, T, phases, max_phases, n_phases, 1, nc.data(), phase_of_material,
                                               O_of_material, v.data(), max_materials, n_materials);
  return rc;
}

// the grid-level keys of Config::parse_numerics (config.cpp:379-384, 449-457; defaults config.h:60-75): staggered-iteration and
// cut-back limits of the increment loop, and the polarisation scheme's parameters
extern "C" int dmk_grid_numerics_from_yaml(const char* numerics_yaml_path, dmk_grid_numerics* out) {
  try {
    if (!out) bad("dmk_grid_numerics_from_yaml: null argument");
    out->N_staggered_iter_max = 10; out->N_cutback_max = 3;
    out->alpha = 1.0; out->beta = 1.0; out->eps_abs_curl = 1.0e-10; out->eps_rel_curl = 5.0e-4;
    if (!numerics_yaml_path || !*numerics_yaml_path) return DMK_OK;
    auto root = yaml_parse_file(numerics_yaml_path);
    const YamlNode* solver = root->get("solver");
    const YamlNode* grid = solver ? solver->get("grid") : nullptr;
    if (!grid) return DMK_OK;
    if (grid->has("N_staggered_iter_max")) out->N_staggered_iter_max = (int)grid->get("N_staggered_iter_max")->as_int();
    if (grid->has("N_cutback_max")) out->N_cutback_max = (int)grid->get("N_cutback_max")->as_int();
    if (const YamlNode* f = grid->get("FFT")) {
      if (f->has("alpha")) out->alpha = f->get("alpha")->as_double();
      if (f->has("beta")) out->beta = f->get("beta")->as_double();
      if (f->has("eps_abs_curl(F)")) out->eps_abs_curl = f->get("eps_abs_curl(F)")->as_double();
      if (f->has("eps_rel_curl(F)")) out->eps_rel_curl = f->get("eps_rel_curl(F)")->as_double();
    }
    return DMK_OK;
  } catch (const std::exception& e) {
    dmk_set_error("%s", e.what());
    return DMK_ERR_ARG;
  }
}

extern "C" int dmk_numerics_from_yaml(const char* numerics_yaml_path, dmk_numerics* num, int* derivative, int* divergence_correction) {
  try {
    default_numerics(*num);
    int der = DMK_DERIVATIVE_CONTINUOUS, div = DMK_DIVCORR_SIZE_GRID;
    if (numerics_yaml_path && *numerics_yaml_path) read_numerics(*yaml_parse_file(numerics_yaml_path), *num, &der, &div);
    // the range checks of upstream DAMASK (grid_mechanical_spectral_basic_init, IO_error 301; phase_mechanical.f90:296-309), in the
    // Python host's order; Config::parse_numerics does not check
    std::string wrong;
    auto flag = [&](bool is_wrong, const char* name) { if (is_wrong) wrong += std::string(wrong.empty() ? "" : ", ") + name; };
    const dmk_numerics& n = *num;
    flag(!(n.eps_abs_div > 0.0), "eps_abs_div"); flag(!(n.eps_rel_div > 0.0), "eps_rel_div"); flag(!(n.eps_abs_P > 0.0), "eps_abs_P");
    flag(!(n.eps_rel_P > 0.0), "eps_rel_P"); flag(!(n.r_cutback_min > 0.0), "r_cutback_min"); flag(!(n.r_cutback > 0.0), "r_cutback");
    flag(!(n.r_increase > 0.0), "r_increase"); flag(!(n.eps_rel_state > 0.0), "eps_rel_state"); flag(!(n.r_linesearch_Lp > 0.0), "r_linesearch_Lp");
    flag(!(n.eps_rel_Lp > 0.0), "eps_rel_Lp"); flag(!(n.eps_abs_Lp > 0.0), "eps_abs_Lp");
    flag(n.N_iter_state_max < 1, "N_iter_state_max"); flag(n.N_iter_Lp_max < 1, "N_iter_Lp_max"); flag(n.f_update_jacobi_Lp < 1, "f_update_jacobi_Lp");
    flag(n.itmax <= 1, "N_iter_max"); flag(n.itmin > n.itmax || n.itmin < 1, "N_iter_min");
    if (!wrong.empty()) bad("numerics: value out of range: " + wrong);
    if (derivative) *derivative = der;
    if (divergence_correction) *divergence_correction = div;
    return DMK_OK;
  } catch (const std::exception& e) {
    dmk_set_error("%s", e.what());
    return DMK_ERR_ARG;
  }
}

// ------------------------------------------------------------------------------------------------ the reference's symbols
// Errors cannot be returned through these signatures (the Fortran originals abort): they are left in dmk_last_error(), the
// context stays unbound, and dmk_bound_context() returns NULL.
extern "C" void f_materialpoint_initBase(const char* material_path, int material_path_len, const char* load_path, int load_path_len,
                                         const char* grid_path, int grid_path_len, const char* numerics_path, int numerics_path_len) {
  g_init.material_path.assign(material_path ? material_path : "", material_path ? (size_t)material_path_len : 0);
  g_init.load_path.assign(load_path ? load_path : "", load_path ? (size_t)load_path_len : 0);
  g_init.grid_path.assign(grid_path ? grid_path : "", grid_path ? (size_t)grid_path_len : 0);
  g_init.numerics_path.assign(numerics_path ? numerics_path : "", numerics_path ? (size_t)numerics_path_len : 0);
}

extern "C" void f_discretization_init(int* material_at, int* n_materialpoints, double* IPcoords0, int* n_ips, double* NodeCoords0,
                                      int* n_nodes, int* shared_nodes_begin) {
  const int n = *n_materialpoints;
  g_init.material_at.resize((size_t)n);
  for (int i = 0; i < n; i++) g_init.material_at[(size_t)i] = material_at[i] - 1;     // 1-based after the VTI read (VTI.f90:474)
  g_init.ip_coords.assign(IPcoords0, IPcoords0 + 3 * (size_t)*n_ips);
  g_init.node_coords.assign(NodeCoords0, NodeCoords0 + 3 * (size_t)*n_nodes);
  g_init.shared_nodes_begin = shared_nodes_begin ? *shared_nodes_begin : 0;
}
extern "C" void f_discretization_setIPcoords(double* IPcoords0) { std::copy(IPcoords0, IPcoords0 + g_init.ip_coords.size(), g_init.ip_coords.begin()); }
extern "C" void f_discretization_setNodeCoords(double* NodeCoords) { std::copy(NodeCoords, NodeCoords + g_init.node_coords.size(), g_init.node_coords.begin()); }
extern "C" void f_discretization_fetch_ip_node_coord_pointers(double** ip_coords_ptr, double** node_coords_ptr) {
  if (ip_coords_ptr) *ip_coords_ptr = g_init.ip_coords.data();
  if (node_coords_ptr) *node_coords_ptr = g_init.node_coords.data();
}
extern "C" void f_CLI_get_geomFilename(char* filename) { if (filename) strcpy(filename, g_init.grid_path.c_str()); }

extern "C" dmk_ctx* dmk_bound_context(void) { return g_init.ctx; }

extern "C" void f_materialpoint_initDamask(void) {
  try {
    if (g_init.ctx) { dmk_bind_global(nullptr); dmk_destroy(g_init.ctx); g_init.ctx = nullptr; }
    const long long n = (long long)g_init.material_at.size();
    if (n < 1 || (long long)g_init.ip_coords.size() != 3 * n) bad("f_materialpoint_initDamask: f_discretization_init has not been called (or n_ips != n_materialpoints)");
    if (g_init.material_path.empty()) bad("f_materialpoint_initDamask: f_materialpoint_initBase has not been called");
    std::vector<dmk_phase> phases(DMK_MAX_PHASES);
    int n_phases = 0, n_mat = 0;
    int max_mat = 0;
    for (int v : g_init.material_at) max_mat = std::max(max_mat, v + 1);
    // material.yaml may hold more entries than the geometry uses: size the tables by the file
    auto root = yaml_parse_file(g_init.material_path);
    const YamlNode* mats = root->get("material");
    const int file_mats = (mats && mats->is_seq()) ? (int)mats->seq.size() : 0;
    if (max_mat > file_mats) bad("geometry refers to material " + std::to_string(max_mat) + ", material.yaml has " + std::to_string(file_mats) + " entries");
    // constituents per material entry: the widest entry sizes the tables
    int max_const = 1;
    for (int m = 0; m < file_mats; m++) { const YamlNode* cs = mats->seq[(size_t)m]->get("constituents"); if (cs && cs->is_seq()) max_const = std::max(max_const, (int)cs->seq.size()); }
    std::vector<int> nc_of_mat((size_t)file_mats), ph_of_mat((size_t)file_mats * max_const);
    std::vector<double> O_of_mat(9 * (size_t)file_mats * max_const), v_of_mat((size_t)file_mats * max_const);
    if (dmk_material_constituents_from_yaml(g_init.material_path.c_str(), T_ROOM, phases.data(), DMK_MAX_PHASES, &n_phases, max_const, nc_of_mat.data(),
                                            ph_of_mat.data(), O_of_mat.data(), v_of_mat.data(), file_mats, &n_mat)) return;
    dmk_numerics num;
    int derivative = 0, divcorr = 0;
    const std::string np = file_exists(g_init.numerics_path) ? g_init.numerics_path : std::string();
    if (dmk_numerics_from_yaml(np.c_str(), &num, &derivative, &divcorr)) return;
    dmk_grid_desc g;
    memset(&g, 0, sizeof g);
    grid_from_ip_coords(g_init.ip_coords, n, g.cells, g.size);
    g.world_size = 1; g.rank = 0;                   // the driver's own MPI decomposition stays with the driver: this is ITS slab
    g.derivative = derivative; g.divergence_correction = divcorr;
    // every material in use must bring the same number of constituents (point k * n + cell is constituent k of the cell)
    int n_const = 0;
    for (long long i = 0; i < n; i++) {
      const int m = g_init.material_at[(size_t)i];
      if (m < 0) bad("negative material index");
      if (n_const == 0) n_const = nc_of_mat[(size_t)m];
      else if (nc_of_mat[(size_t)m] != n_const) bad("materials with different numbers of constituents in one grid are not offered");
    }
    std::vector<int> phase_of_point((size_t)n * n_const);
    std::vector<double> O_of_point(9 * (size_t)n * n_const), v_of_point((size_t)n * n_const);
    for (int k = 0; k < n_const; k++)
      for (long long i = 0; i < n; i++) {
        const size_t at = (size_t)g_init.material_at[(size_t)i] * max_const + (size_t)k, pt = (size_t)k * n + (size_t)i;
        phase_of_point[pt] = ph_of_mat[at];
        v_of_point[pt] = v_of_mat[at];
        memcpy(&O_of_point[9 * pt], &O_of_mat[9 * at], 9 * sizeof(double));
      }
    dmk_ctx* c = nullptr;
    if (dmk_create(&c, &g)) return;
    if (dmk_material_init_constituents(c, phases.data(), n_phases, n_const, phase_of_point.data(), O_of_point.data(), v_of_point.data(), &num) ||
        dmk_bind_global(c)) { dmk_destroy(c); return; }
    g_init.ctx = c;
  } catch (const std::exception& e) {
    dmk_set_error("%s", e.what());
  }
}

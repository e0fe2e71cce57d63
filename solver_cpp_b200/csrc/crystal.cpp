// Host-side crystallography for the phase tables uploaded at dmk_material_init: slip / twin geometry, Schmid and
// non-Schmid projections, interaction matrices, forest projections, twinned stiffness, orientation matrices.
// Runs once at set-up (SURVEY: "HOT-init"); cF is complete, cI offers geometry, Schmid/non-Schmid, twins.
//
// Restates the lattice data and rules of /root/reference/src/crystal.f90 (systems :41-95 cF, :118-199 cI; coordinate
// systems :1897-1961; Schmid matrices :1353-1478; interaction types :602-648, :920-945, :1061-1100, :1274-1290;
// buildInteraction :1853-1890; forest projection :1764-1806; twinned C66 :500-531) and of
// /root/reference/src/rotations.f90:423-447 (qu2om, P = -1).
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/dmk_grid.h"

void dmk_set_error(const char* fmt, ...);

namespace {

struct Lattice {
  std::vector<int> n_slip_family, n_twin_family;
  std::vector<std::vector<int>> slip, twin;          // rows: direction(3) | plane normal(3)
  std::vector<std::vector<int>> slip_slip, twin_twin, slip_twin, twin_slip;   // [reacting][acting] interaction type, 1-based
};

std::vector<std::vector<int>> rows6(const std::vector<int>& flat) {
  std::vector<std::vector<int>> r(flat.size() / 6, std::vector<int>(6));
  for (size_t i = 0; i < flat.size(); i++) r[i / 6][i % 6] = flat[i];
  return r;
}

// ---- cF: <110>{111} x 12 then <110>{110} x 6; twins <112>{111} x 12
const Lattice& lattice_cF() {
  static Lattice L;
  if (!L.slip.empty()) return L;
  L.n_slip_family = {12, 6};
  L.n_twin_family = {12};
  // {111} planes in Thompson order d, c, b, a; three <110> directions per plane
  const int planes[4][3] = {{1, 1, 1}, {-1, -1, 1}, {1, -1, -1}, {-1, 1, -1}};
  const int dirs[12][3] = {{0, 1, -1}, {-1, 0, 1}, {1, -1, 0}, {0, -1, -1}, {1, 0, 1}, {-1, 1, 0},
                           {0, -1, 1}, {-1, 0, -1}, {1, 1, 0}, {0, 1, 1}, {1, 0, -1}, {-1, -1, 0}};
  for (int s = 0; s < 12; s++) L.slip.push_back({dirs[s][0], dirs[s][1], dirs[s][2], planes[s / 3][0], planes[s / 3][1], planes[s / 3][2]});
  const int nonoct[6][6] = {{1, 1, 0, 1, -1, 0}, {1, -1, 0, 1, 1, 0}, {1, 0, 1, 1, 0, -1}, {1, 0, -1, 1, 0, 1}, {0, 1, 1, 0, 1, -1}, {0, 1, -1, 0, 1, 1}};
  for (auto& r : nonoct) L.slip.push_back(std::vector<int>(r, r + 6));
  const int tdirs[12][3] = {{-2, 1, 1}, {1, -2, 1}, {1, 1, -2}, {2, -1, 1}, {-1, 2, 1}, {-1, -1, -2},
                            {-2, -1, -1}, {1, 2, -1}, {1, -1, 2}, {2, 1, -1}, {-1, -2, -1}, {-1, 1, 2}};
  for (int s = 0; s < 12; s++) L.twin.push_back({tdirs[s][0], tdirs[s][1], tdirs[s][2], planes[s / 3][0], planes[s / 3][1], planes[s / 3][2]});
  // slip-slip interaction types (Madec & Kubin 2017 labelling as used by the reference), row = reacting, column = acting
  static const int ss[18][18] = {
      {1, 2, 2, 4, 7, 5, 3, 5, 5, 4, 6, 7, 10, 11, 10, 11, 12, 13},  {2, 1, 2, 7, 4, 5, 6, 4, 7, 5, 3, 5, 10, 11, 12, 13, 10, 11},
      {2, 2, 1, 5, 5, 3, 6, 7, 4, 7, 6, 4, 12, 13, 10, 11, 10, 11},  {4, 7, 6, 1, 2, 2, 4, 6, 7, 3, 5, 5, 10, 11, 11, 10, 13, 12},
      {7, 4, 6, 2, 1, 2, 5, 3, 5, 6, 4, 7, 10, 11, 13, 12, 11, 10},  {5, 5, 3, 2, 2, 1, 7, 6, 4, 6, 7, 4, 12, 13, 11, 10, 11, 10},
      {3, 5, 5, 4, 6, 7, 1, 2, 2, 4, 7, 6, 11, 10, 11, 10, 12, 13},  {6, 4, 7, 5, 3, 5, 2, 1, 2, 7, 4, 6, 11, 10, 13, 12, 10, 11},
      {6, 7, 4, 7, 6, 4, 2, 2, 1, 5, 5, 3, 13, 12, 11, 10, 10, 11},  {4, 6, 7, 3, 5, 5, 4, 7, 6, 1, 2, 2, 11, 10, 10, 11, 13, 12},
      {5, 3, 5, 6, 4, 7, 7, 4, 6, 2, 1, 2, 11, 10, 12, 13, 11, 10},  {7, 6, 4, 6, 7, 4, 5, 5, 3, 2, 2, 1, 13, 12, 10, 11, 11, 10},
      {10, 10, 12, 10, 10, 12, 11, 11, 13, 11, 11, 13, 1, 8, 9, 9, 9, 9}, {11, 11, 13, 11, 11, 13, 10, 10, 12, 10, 10, 12, 8, 1, 9, 9, 9, 9},
      {10, 12, 10, 11, 13, 11, 11, 13, 11, 10, 12, 10, 9, 9, 1, 8, 9, 9}, {11, 13, 11, 10, 12, 10, 10, 12, 10, 11, 13, 11, 9, 9, 8, 1, 9, 9},
      {12, 10, 10, 13, 11, 11, 12, 10, 10, 13, 11, 11, 9, 9, 9, 9, 1, 8}, {13, 11, 11, 12, 10, 10, 13, 11, 11, 12, 10, 10, 9, 9, 9, 9, 8, 1}};
  for (auto& r : ss) L.slip_slip.push_back(std::vector<int>(r, r + 18));
  // twin-twin: coplanar 1, else 2
  for (int i = 0; i < 12; i++) { std::vector<int> r(12); for (int j = 0; j < 12; j++) r[j] = (i / 3 == j / 3) ? 1 : 2; L.twin_twin.push_back(r); }
  // slip-twin: 1 coplanar, 2 screw trace of the slip direction in the twin habit plane, 3 other; non-octahedral slip: 4
  for (int i = 0; i < 18; i++) {
    std::vector<int> r(12);
    for (int j = 0; j < 12; j++) {
      if (i >= 12) { r[j] = 4; continue; }
      if (i / 3 == j / 3) { r[j] = 1; continue; }
      const int* d = dirs[i]; const int* n = planes[j / 3];
      r[j] = (d[0] * n[0] + d[1] * n[1] + d[2] * n[2] == 0) ? 2 : 3;
    }
    L.slip_twin.push_back(r);
  }
  for (int i = 0; i < 12; i++) L.twin_slip.push_back(std::vector<int>(18, 1));
  return L;
}

// ---- cI: <111>{110} x 12, <111>{112} x 12, <111>{123} x 24; twins <111>{112} x 12
const Lattice& lattice_cI() {
  static Lattice L;
  if (!L.slip.empty()) return L;
  L.n_slip_family = {12, 12, 24};
  L.n_twin_family = {12};
  L.slip = rows6({
      1, -1, 1, 0, 1, 1,   -1, -1, 1, 0, -1, -1,   1, 1, 1, 0, 1, -1,    -1, 1, 1, 0, -1, 1,
      -1, 1, 1, -1, 0, -1, -1, -1, 1, 1, 0, 1,     1, 1, 1, -1, 0, 1,    1, -1, 1, 1, 0, -1,
      -1, 1, 1, 1, 1, 0,   1, -1, 1, -1, -1, 0,    1, 1, 1, 1, -1, 0,    -1, -1, 1, -1, 1, 0,
      -1, 1, 1, 2, 1, 1,   1, 1, 1, -2, 1, 1,      1, 1, -1, 2, -1, 1,   1, -1, 1, 2, 1, -1,
      1, -1, 1, 1, 2, 1,   1, 1, -1, -1, 2, 1,     1, 1, 1, 1, -2, 1,    -1, 1, 1, 1, 2, -1,
      1, 1, -1, 1, 1, 2,   1, -1, 1, -1, 1, 2,     -1, 1, 1, 1, -1, 2,   1, 1, 1, 1, 1, -2,
      1, 1, -1, 1, 2, 3,   1, -1, 1, -1, 2, 3,     -1, 1, 1, 1, -2, 3,   1, 1, 1, 1, 2, -3,
      1, -1, 1, 1, 3, 2,   1, 1, -1, -1, 3, 2,     1, 1, 1, 1, -3, 2,    -1, 1, 1, 1, 3, -2,
      1, 1, -1, 2, 1, 3,   1, -1, 1, -2, 1, 3,     -1, 1, 1, 2, -1, 3,   1, 1, 1, 2, 1, -3,
      1, -1, 1, 2, 3, 1,   1, 1, -1, -2, 3, 1,     1, 1, 1, 2, -3, 1,    -1, 1, 1, 2, 3, -1,
      -1, 1, 1, 3, 1, 2,   1, 1, 1, -3, 1, 2,      1, 1, -1, 3, -1, 2,   1, -1, 1, 3, 1, -2,
      -1, 1, 1, 3, 2, 1,   1, 1, 1, -3, 2, 1,      1, 1, -1, 3, -2, 1,   1, -1, 1, 3, 2, -1});
  for (int i = 12; i < 24; i++) L.twin.push_back(L.slip[i]);
  // slip-slip interaction types, <111>{110} family only (first 12 of the 48 systems; crystal.f90:648-701); 0 = not offered
  static const int ss[12][12] = {
      {1, 3, 6, 6, 7, 5, 4, 2, 4, 2, 7, 5},
      {3, 1, 6, 6, 4, 2, 7, 5, 7, 5, 4, 2},
      {6, 6, 1, 3, 5, 7, 2, 4, 5, 7, 2, 4},
      {6, 6, 3, 1, 2, 4, 5, 7, 2, 4, 5, 7},
      {7, 5, 4, 2, 1, 3, 6, 6, 2, 4, 7, 5},
      {4, 2, 7, 5, 3, 1, 6, 6, 5, 7, 4, 2},
      {5, 7, 2, 4, 6, 6, 1, 3, 7, 5, 2, 4},
      {2, 4, 5, 7, 6, 6, 3, 1, 4, 2, 5, 7},
      {5, 7, 4, 2, 2, 4, 7, 5, 1, 3, 6, 6},
      {2, 4, 7, 5, 5, 7, 4, 2, 3, 1, 6, 6},
      {7, 5, 2, 4, 7, 5, 2, 4, 6, 6, 1, 3},
      {4, 2, 5, 7, 4, 2, 5, 7, 6, 6, 3, 1}};
  for (int i = 0; i < 48; i++) { std::vector<int> r(48, 0); if (i < 12) for (int j = 0; j < 12; j++) r[j] = ss[i][j]; L.slip_slip.push_back(r); }
  static const int tt[12][12] = {
      {1, 3, 3, 3, 3, 3, 3, 2, 3, 3, 2, 3}, {3, 1, 3, 3, 3, 3, 2, 3, 3, 3, 3, 2}, {3, 3, 1, 3, 3, 2, 3, 3, 2, 3, 3, 3},
      {3, 3, 3, 1, 2, 3, 3, 3, 3, 2, 3, 3}, {3, 3, 3, 2, 1, 3, 3, 3, 3, 2, 3, 3}, {3, 3, 2, 3, 3, 1, 3, 3, 2, 3, 3, 3},
      {3, 2, 3, 3, 3, 3, 1, 3, 3, 3, 3, 2}, {2, 3, 3, 3, 3, 3, 3, 1, 3, 3, 2, 3}, {3, 3, 2, 3, 3, 2, 3, 3, 1, 3, 3, 3},
      {3, 3, 3, 2, 2, 3, 3, 3, 3, 1, 3, 3}, {2, 3, 3, 3, 3, 3, 3, 2, 3, 3, 1, 3}, {3, 2, 3, 3, 3, 3, 2, 3, 3, 3, 3, 1}};
  for (auto& r : tt) L.twin_twin.push_back(std::vector<int>(r, r + 12));
  for (int i = 0; i < 12; i++) L.twin_slip.push_back(std::vector<int>(48, 1));
  return L;
}

const Lattice* get_lattice(const char* name) {
  if (name && !strcmp(name, "cF")) return &lattice_cF();
  if (name && !strcmp(name, "cI")) return &lattice_cI();
  dmk_set_error("lattice '%s' is not offered (cF, cI)", name ? name : "(null)");
  return nullptr;
}

// indices of the active systems inside the table of potential systems (first N of each family)
bool active_rows(const int* N, int nf, const std::vector<int>& potential, std::vector<int>* idx) {
  idx->clear();
  if (nf > (int)potential.size()) { dmk_set_error("more families (%d) than the lattice has (%zu)", nf, potential.size()); return false; }
  int start = 0;
  for (int f = 0; f < nf; f++) {
    if (N[f] < 0 || N[f] > potential[f]) { dmk_set_error("family %d: %d active systems, %d possible", f, N[f], potential[f]); return false; }
    for (int i = 0; i < N[f]; i++) idx->push_back(start + i);
    start += potential[f];
  }
  return true;
}

struct Frame { double d[3], n[3], t[3]; };
void unit(const int* v, double* u) {
  double l = std::sqrt((double)(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
  for (int i = 0; i < 3; i++) u[i] = v[i] / l;
}
void cross(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
bool frames(const int* N, int nf, const std::vector<int>& potential, const std::vector<std::vector<int>>& sys, std::vector<Frame>* out) {
  std::vector<int> idx;
  if (!active_rows(N, nf, potential, &idx)) return false;
  out->clear();
  for (int r : idx) { Frame f; unit(&sys[r][0], f.d); unit(&sys[r][3], f.n); cross(f.d, f.n, f.t); out->push_back(f); }
  return true;
}
void outer(const double* a, const double* b, double s, double* M) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M[3 * i + j] += s * a[i] * b[j]; }

void voigt_to_3333(const double* C66, double* C) {
  static const int V[6][2] = {{0, 0}, {1, 1}, {2, 2}, {1, 2}, {0, 2}, {0, 1}};
  for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) {
    int i = V[a][0], j = V[a][1], k = V[b][0], l = V[b][1]; double v = C66[6 * a + b];
    C[((i * 3 + j) * 3 + k) * 3 + l] = v; C[((j * 3 + i) * 3 + k) * 3 + l] = v; C[((i * 3 + j) * 3 + l) * 3 + k] = v; C[((j * 3 + i) * 3 + l) * 3 + k] = v;
  }
}
}  // namespace

extern "C" {

int dmk_crystal_schmid_slip_ns(const char* lattice, const int* N_sl, int n_fam, const double* non_schmid, int n_coeff, int sense, double* P) {
  const Lattice* L = get_lattice(lattice); if (!L) return 3;
  std::vector<Frame> fr;
  if (!frames(N_sl, n_fam, L->n_slip_family, L->slip, &fr)) return 2;
  std::vector<int> fam; for (int f = 0; f < n_fam; f++) for (int i = 0; i < N_sl[f]; i++) fam.push_back(f);
  for (size_t s = 0; s < fr.size(); s++) {
    double d[3] = {fr[s].d[0] * sense, fr[s].d[1] * sense, fr[s].d[2] * sense};
    const double* n = fr[s].n;
    double* M = P + 9 * s;
    for (int k = 0; k < 9; k++) M[k] = 0.0;
    outer(d, n, 1.0, M);
    if (non_schmid && n_coeff > 0 && !strcmp(lattice, "cI") && fam[s] == 0) {
      double c[6] = {0, 0, 0, 0, 0, 0};
      for (int k = 0; k < n_coeff && k < 6; k++) c[k] = non_schmid[k];
      // n' = normal rotated by 60 degrees about the (signed) slip direction, in the sense the reference obtains from
      // fromAxisAngle(P=1) + passive rotate: the active rotation about -d
      const double ang = 60.0 * M_PI / 180.0, cs = std::cos(ang), sn = std::sin(ang);
      double ax[3] = {-d[0], -d[1], -d[2]}, axn[3], np[3];
      cross(ax, n, axn);
      double dot = ax[0] * n[0] + ax[1] * n[1] + ax[2] * n[2];
      for (int k = 0; k < 3; k++) np[k] = n[k] * cs + axn[k] * sn + ax[k] * dot * (1.0 - cs);
      double nxd[3], npxd[3];
      cross(n, d, nxd); cross(np, d, npxd);
      outer(d, np, c[0], M); outer(nxd, n, c[1], M); outer(npxd, np, c[2], M);
      outer(n, n, c[3], M); outer(nxd, nxd, c[4], M); outer(d, d, c[5], M);
    }
  }
  return 0;
}

int dmk_crystal_schmid_slip(const char* lattice, const int* N_sl, int n_fam, double* P) {
  return dmk_crystal_schmid_slip_ns(lattice, N_sl, n_fam, nullptr, 0, 1, P);
}

int dmk_crystal_schmid_twin(const char* lattice, const int* N_tw, int n_fam, double* P) {
  const Lattice* L = get_lattice(lattice); if (!L) return 3;
  std::vector<Frame> fr;
  if (!frames(N_tw, n_fam, L->n_twin_family, L->twin, &fr)) return 2;
  for (size_t s = 0; s < fr.size(); s++) { double* M = P + 9 * s; for (int k = 0; k < 9; k++) M[k] = 0.0; outer(fr[s].d, fr[s].n, 1.0, M); }
  return 0;
}

int dmk_crystal_slip_normal(const char* lattice, const int* N_sl, int n_fam, double* n) {
  const Lattice* L = get_lattice(lattice); if (!L) return 3;
  std::vector<Frame> fr;
  if (!frames(N_sl, n_fam, L->n_slip_family, L->slip, &fr)) return 2;
  for (size_t s = 0; s < fr.size(); s++) for (int k = 0; k < 3; k++) n[3 * s + k] = fr[s].n[k];
  return 0;
}

int dmk_crystal_interaction(const char* kind, const char* lattice, const int* N_react, int nf_react, const int* N_act, int nf_act,
                            const double* values, int n_values, double* h) {
  const Lattice* L = get_lattice(lattice); if (!L) return 3;
  const std::vector<std::vector<int>>* tab = nullptr; const std::vector<int>*pr = nullptr, *pa = nullptr;
  std::string k(kind ? kind : "");
  if (k == "sl-sl") { tab = &L->slip_slip; pr = &L->n_slip_family; pa = &L->n_slip_family; }
  else if (k == "tw-tw") { tab = &L->twin_twin; pr = &L->n_twin_family; pa = &L->n_twin_family; }
  else if (k == "sl-tw") { tab = &L->slip_twin; pr = &L->n_slip_family; pa = &L->n_twin_family; }
  else if (k == "tw-sl") { tab = &L->twin_slip; pr = &L->n_twin_family; pa = &L->n_slip_family; }
  if (!tab) { dmk_set_error("unknown interaction kind '%s'", k.c_str()); return 2; }
  if (tab->empty()) { dmk_set_error("%s interaction table of lattice %s is not offered", k.c_str(), lattice); return 3; }
  std::vector<int> ir, ia;
  if (!active_rows(N_react, nf_react, *pr, &ir) || !active_rows(N_act, nf_act, *pa, &ia)) return 2;
  for (size_t r = 0; r < ir.size(); r++)
    for (size_t a = 0; a < ia.size(); a++) {
      const int type = (*tab)[ir[r]][ia[a]];
      if (type < 1) { dmk_set_error("%s interaction types of these system families of lattice %s are not offered", k.c_str(), lattice); return 3; }
      if (type > n_values) { dmk_set_error("%s interaction needs at least %d values, %d given", k.c_str(), type, n_values); return 2; }
      h[r * ia.size() + a] = values[type - 1];
    }
  return 0;
}

// forest projection (dislotwin.f90:229-232): f_edge(j) |n_i . t_j| + (1 - f_edge(j)) |n_i . d_j|
int dmk_crystal_forest_projection(const char* lattice, const int* N_sl, int n_fam, const double* f_edge, double* proj) {
  const Lattice* L = get_lattice(lattice); if (!L) return 3;
  std::vector<Frame> fr;
  if (!frames(N_sl, n_fam, L->n_slip_family, L->slip, &fr)) return 2;
  const size_t n = fr.size();
  for (size_t i = 0; i < n; i++)
    for (size_t j = 0; j < n; j++) {
      double e = 0, s = 0;
      for (int k = 0; k < 3; k++) { e += fr[i].n[k] * fr[j].t[k]; s += fr[i].n[k] * fr[j].d[k]; }
      proj[i * n + j] = f_edge[j] * std::fabs(e) + (1.0 - f_edge[j]) * std::fabs(s);
    }
  return 0;
}

// stiffness of each twin variant: C rotated by 180 degrees about the twin-plane normal (crystal.f90:526-529)
int dmk_crystal_C66_twin(const char* lattice, const int* N_tw, int n_fam, const double* C66, double* C66_tw) {
  const Lattice* L = get_lattice(lattice); if (!L) return 3;
  std::vector<Frame> fr;
  if (!frames(N_tw, n_fam, L->n_twin_family, L->twin, &fr)) return 2;
  static const int V[6][2] = {{0, 0}, {1, 1}, {2, 2}, {1, 2}, {0, 2}, {0, 1}};
  double C[81];
  voigt_to_3333(C66, C);
  for (size_t s = 0; s < fr.size(); s++) {
    double R[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[3 * i + j] = 2.0 * fr[s].n[i] * fr[s].n[j] - (i == j ? 1.0 : 0.0);
    for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) {
      const int i = V[a][0], j = V[a][1], k = V[b][0], l = V[b][1];
      double acc = 0.0;
      for (int m = 0; m < 3; m++) for (int n = 0; n < 3; n++) for (int o = 0; o < 3; o++) for (int p = 0; p < 3; p++)
        acc += R[3 * i + m] * R[3 * j + n] * R[3 * k + o] * R[3 * l + p] * C[((m * 3 + n) * 3 + o) * 3 + p];
      C66_tw[36 * s + 6 * a + b] = acc;
    }
  }
  return 0;
}

int dmk_crystal_twin_nucleation_pairs(const int* N_tw, int n_fam, int* pairs) {
  const Lattice& L = lattice_cF();
  std::vector<int> idx;
  if (!active_rows(N_tw, n_fam, L.n_twin_family, &idx)) return 2;
  // twin i on plane p = i / 3 is nucleated by the two slip systems of that plane other than i % 3 (crystal.f90:82-95)
  for (size_t s = 0; s < idx.size(); s++) {
    const int p = idx[s] / 3, k = idx[s] % 3;
    int w = 0;
    for (int q = 0; q < 3; q++) if (q != k) pairs[2 * s + w++] = 3 * p + q;
  }
  return 0;
}

// unit quaternion (w,x,y,z) -> rotation matrix, DAMASK convention P = -1, normalised to det = 1
void dmk_qu2om(const double q[4], double om[9]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double qq = w * w - (x * x + y * y + z * z);
  double m[9];
  m[0] = qq + 2.0 * x * x; m[4] = qq + 2.0 * y * y; m[8] = qq + 2.0 * z * z;
  m[1] = 2.0 * (x * y - w * z); m[5] = 2.0 * (y * z - w * x); m[6] = 2.0 * (z * x - w * y);
  m[3] = 2.0 * (y * x + w * z); m[7] = 2.0 * (z * y + w * x); m[2] = 2.0 * (x * z + w * y);
  double t[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = m[3 * j + i];        // P = -1
  const double det = t[0] * (t[4] * t[8] - t[5] * t[7]) - t[1] * (t[3] * t[8] - t[5] * t[6]) + t[2] * (t[3] * t[7] - t[4] * t[6]);
  const double s = std::cbrt(det);
  for (int k = 0; k < 9; k++) om[k] = t[k] / s;
}

}  // extern "C"

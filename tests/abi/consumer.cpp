// A compiled C++ consumer of include/dmk_grid.h, written the way the reference's driver uses the boundary
// (/root/reference/src/grid/main.cpp:80-86, discretization_grid.h:101-112, spectral.h:16-23,43-67, mech_solver_basic.cpp:286-289):
//   level 1  f_materialpoint_initBase -> f_discretization_init -> f_materialpoint_initDamask, then the Fortran-named
//            f_homogenization_* symbols on the callee-owned host arrays;
//   level 2  the mechanical_spectral_basic surface on the same context, with the residual evaluated through a function of
//            PETSc's DMDA signature  PetscErrorCode (*)(DMDALocalInfo*, double***, double***, void*)  that only
//            dereferences &F[0][0][0] / &r[0][0][0], exactly as MechSolverBasic::form_residual does.
// Prints its results as `key value...` lines; tests/test_gpu_consumer.py compares them with the Python host on the same input.
// Usage: consumer material.yaml nx ny nz n_materials [numerics.yaml]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "dmk_grid.h"

// ---- stand-ins for the two PETSc types in the residual's signature
typedef int PetscErrorCode;
struct DMDALocalInfo { int dim, dof, sw; int mx, my, mz; int xs, ys, zs, xm, ym, zm; };

static dmk_ctx* g_ctx = nullptr;

static PetscErrorCode form_residual(DMDALocalInfo* info, double*** F, double*** r, void* user) {
  (void)info; (void)user;
  return dmk_mech_form_residual_host(g_ctx, &F[0][0][0], &r[0][0][0]);
}

// double*** view [z][y][9 * x] over a contiguous field, as DMDAVecGetArray hands it out
struct Field3 {
  std::vector<double> data; std::vector<double*> rows; std::vector<double**> planes;
  Field3(int nx, int ny, int nz) : data((size_t)9 * nx * ny * nz), rows((size_t)ny * nz), planes((size_t)nz) {
    for (int k = 0; k < nz; k++) { for (int j = 0; j < ny; j++) rows[(size_t)k * ny + j] = data.data() + ((size_t)k * ny + j) * 9 * nx; planes[(size_t)k] = rows.data() + (size_t)k * ny; }
  }
  double*** ptr() { return planes.data(); }
};

static void print9(const char* key, const double* v) { printf("%s", key); for (int a = 0; a < 9; a++) printf(" %.17g", v[a]); printf("\n"); }

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: consumer material.yaml nx ny nz n_materials [numerics.yaml]\n"); return 2; }
  const std::string material = argv[1], numerics = argc > 6 ? argv[6] : "numerics.yaml", load = "tensionX.yaml", grid = "grid.vti";
  const int nx = atoi(argv[2]), ny = atoi(argv[3]), nz = atoi(argv[4]), n_mat = atoi(argv[5]);
  int n = nx * ny * nz;
  const double size[3] = {1.0e-5 * nx / ny, 1.0e-5, 1.0e-5 * nz / ny};
  // ---- level 1: the init trio
  f_materialpoint_initBase(material.c_str(), (int)material.size(), load.c_str(), (int)load.size(), grid.c_str(), (int)grid.size(),
                           numerics.c_str(), (int)numerics.size());
  std::vector<int> material_at((size_t)n);
  std::vector<double> ip((size_t)3 * n), nodes((size_t)3 * (nx + 1) * (ny + 1) * (nz + 1));
  for (int k = 0, i = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int x = 0; x < nx; x++, i++) {
    material_at[(size_t)i] = (x / 2 + 3 * (j / 2) + 5 * (k / 2)) % n_mat + 1;          // 1-based, as after the VTI read
    ip[(size_t)3 * i] = size[0] / nx * (x + 0.5); ip[(size_t)3 * i + 1] = size[1] / ny * (j + 0.5); ip[(size_t)3 * i + 2] = size[2] / nz * (k + 0.5);
  }
  for (int k = 0, i = 0; k <= nz; k++) for (int j = 0; j <= ny; j++) for (int x = 0; x <= nx; x++, i++) {
    nodes[(size_t)3 * i] = size[0] / nx * x; nodes[(size_t)3 * i + 1] = size[1] / ny * j; nodes[(size_t)3 * i + 2] = size[2] / nz * k;
  }
  int n_nodes = (nx + 1) * (ny + 1) * (nz + 1), shared = (nx + 1) * (ny + 1) * nz;
  f_discretization_init(material_at.data(), &n, ip.data(), &n, nodes.data(), &n_nodes, &shared);
  f_materialpoint_initDamask();
  g_ctx = dmk_bound_context();
  if (!g_ctx) { fprintf(stderr, "initDamask failed: %s\n", dmk_last_error()); return 1; }
  double *F0 = nullptr, *F = nullptr, *P = nullptr, *dPdF = nullptr; void* ill = nullptr;
  f_homogenization_fetch_tensor_pointers(&F0, &F, &P, &dPdF, &ill);
  if (!F0 || !F || !P || !dPdF || !ill) { fprintf(stderr, "fetch_tensor_pointers failed: %s\n", dmk_last_error()); return 1; }
  for (int i = 0; i < n; i++) {
    const double s = 1.0e-3 * std::sin(0.7 * i), c = 5.0e-4 * std::cos(0.3 * i);
    const double Fi[9] = {1.0 + 1.0e-3 + s, c, 0.0, -c, 1.0 - 4.0e-4, 0.5 * s, 0.0, 0.0, 1.0 - 4.0e-4 + c};   // natural [i][j]
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) F[(size_t)9 * i + a + 3 * b] = Fi[3 * a + b];         // Fortran order
  }
  double dt = 0.5; int one = 1, ip_range[2] = {1, 1}, el_range[2] = {1, n};
  f_homogenization_mechanical_response(&dt, &one, &n);
  f_homogenization_mechanical_response2(&dt, ip_range, el_range);
  double Pav[9] = {0}, Csum = 0.0;
  for (int i = 0; i < n; i++) for (int a = 0; a < 9; a++) Pav[a] += P[(size_t)9 * i + a] / n;
  for (size_t k = 0; k < (size_t)81 * n; k++) Csum += dPdF[k];
  printf("level1_ill %d\n", *(int*)ill);
  print9("level1_P_av", Pav);
  printf("level1_dPdF_sum %.17g\n", Csum);
  // ---- level 2: the mechanical_spectral_basic surface, residual through the DMDA-style signature
  if (dmk_mech_init(g_ctx, nullptr)) { fprintf(stderr, "mech_init: %s\n", dmk_last_error()); return 1; }
  const double dotF[9] = {1.0e-3, 0, 0, 0, 0, 0, 0, 0, 0}, Pbc[9] = {0};
  const int mask_F[9] = {0, 0, 0, 0, 1, 0, 0, 0, 1}, mask_P[9] = {1, 1, 1, 1, 0, 1, 1, 1, 0};     // 1 = 'x' (tensionX)
  const double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (dmk_mech_forward(g_ctx, 0, 0, dt, 1.0, 5.0, "dot_F", dotF, mask_F, "P", Pbc, mask_P, R)) { fprintf(stderr, "forward: %s\n", dmk_last_error()); return 1; }
  if (dmk_materialpoint_forward(g_ctx) || dmk_mech_prepare(g_ctx)) { fprintf(stderr, "prepare: %s\n", dmk_last_error()); return 1; }
  Field3 Ff(nx, ny, nz), rf(nx, ny, nz);
  if (dmk_mech_get_field(g_ctx, "F", Ff.data.data())) { fprintf(stderr, "get_field: %s\n", dmk_last_error()); return 1; }
  DMDALocalInfo info = {3, 9, 0, nx, ny, nz, 0, 0, 0, nx, ny, nz};
  for (int it = 0; it < 2; it++) {                                  // two fixed-point steps of the basic scheme
    if (form_residual(&info, Ff.ptr(), rf.ptr(), nullptr)) { fprintf(stderr, "form_residual: %s\n", dmk_last_error()); return 1; }
    double v[9], e[1], sum = 0.0, sq = 0.0;
    dmk_mech_get(g_ctx, "P_av", v, 9);
    printf("level2_eval %d\n", it);
    print9("level2_P_av", v);
    dmk_mech_get(g_ctx, "err_div", e, 1); printf("level2_err_div %.17g\n", e[0]);
    dmk_mech_get(g_ctx, "err_BC", e, 1); printf("level2_err_BC %.17g\n", e[0]);
    for (double x : rf.data) { sum += x; sq += x * x; }
    printf("level2_r_sum %.17g\nlevel2_r_sq %.17g\n", sum, sq);
    for (size_t k = 0; k < Ff.data.size(); k++) Ff.data[k] -= rf.data[k];
  }
  // ---- the pure helpers of utilities_fortran.h
  double A[9] = {4, 1, 0, 1, 3, 1, 0, 1, 2}, Ai[9]; int err = 7, three = 3;
  f_math_invert(Ai, &err, A, &three);
  printf("math_invert_err %d\nmath_invert_00 %.17g\n", err, Ai[0]);
  f_materialpoint_forward();
  printf("CONSUMER_OK\n");
  return 0;
}

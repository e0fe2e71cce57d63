// What the reference's main (/root/reference/src/grid/main.cpp:31-101) reduces to above include/dmk_grid.h, in C++ and without any
// Python: geometry (VTI) -> init trio -> mechanical solver -> load case, every step a call into the C-ABI library:
//   DiscretizationGrid::init           f_VTI_readCellsSizeOrigin / f_VTI_readDataset_int, f_discretization_init
//   f_materialpoint_initBase/Damask    material.yaml, numerics.yaml -> context, per-point state on the device
//   Config::parse_load / parse_numerics   dmk_load_from_yaml, dmk_grid_numerics_from_yaml
//   MechSolverBasic::init              dmk_mech_init (or dmk_mech_init_polarisation when the load file asks for that scheme)
//   DamaskGrid::loop_over_loadcase     dmk_loop_over_loadcase above the context's own solver
// Prints one line per increment the load step's f_out selects: `increment n t iterations P_av(9, row-major) F_aim(9)`.
// Usage: native_driver geometry.vti material.yaml load.yaml [numerics.yaml [max_increments]]
// Built by __graft_entry__.build() (a compile and link check of the boundary: every symbol used here must be declared in the
// header and exported by the library).  STATUS: written after this round's GPU budget was spent -- compiled and linked, not yet
// run on a GPU; the calls it makes are the ones tests/abi/consumer.cpp and the Python host exercise on the device, except
// dmk_loop_over_loadcase's default table (three one-line calls, csrc/loadcase.cpp).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "dmk_grid.h"

static int die(const char* what) { fprintf(stderr, "native_driver: %s: %s\n", what, dmk_last_error()); return 1; }

static int report(void* user, const dmk_increment_record* rec) {
  dmk_ctx* ctx = (dmk_ctx*)user;
  double P[9], Fa[9];
  if (dmk_mech_get(ctx, "P_av", P, 9) || dmk_mech_get(ctx, "F_aim", Fa, 9)) return DMK_ERR_STATE;
  printf("increment %d %.17g %d", rec->inc, rec->t, rec->iterations);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) printf(" %.17g", P[i + 3 * j]);          // dmk_mech_get: Fortran order
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) printf(" %.17g", Fa[i + 3 * j]);
  printf("\n");
  return DMK_OK;
}

// the default operations of the loop, spelled out, so that `result` can be added to them
static int op_forward(void* u, int cut_back, int guess, double dt, double dt_old, double t_remaining, const dmk_load_step* st) {
  return dmk_mech_forward((dmk_ctx*)u, cut_back, guess, dt, dt_old, t_remaining, st->deformation_type, st->deformation_values,
                          st->deformation_mask, st->stress_type, st->stress_values, st->stress_mask, st->R);
}
static int op_point_forward(void* u) { return dmk_materialpoint_forward((dmk_ctx*)u); }
static int op_solve(void* u, dmk_solution_state* out) { return dmk_mech_calculate_solution((dmk_ctx*)u, out); }

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: native_driver geometry.vti material.yaml load.yaml [numerics.yaml [max_increments]]\n"); return 2; }
  const std::string geometry = argv[1], material = argv[2], load = argv[3], numerics = argc > 4 ? argv[4] : "";
  const int max_increments = argc > 5 ? atoi(argv[5]) : 0;
  // ---- DiscretizationGrid::init (discretization_grid.cpp:8-99)
  std::ifstream in(geometry, std::ios::binary);
  if (!in) { fprintf(stderr, "native_driver: cannot read %s\n", geometry.c_str()); return 2; }
  std::stringstream buf; buf << in.rdbuf();
  const std::string vti = buf.str();
  int cells[3]; double size[3], origin[3];
  f_VTI_readCellsSizeOrigin(vti.c_str(), (int)vti.size(), cells, size, origin);
  if (cells[0] < 1) return die("f_VTI_readCellsSizeOrigin");
  int n = cells[0] * cells[1] * cells[2];
  std::vector<int> material_at((size_t)n);
  int n_read = n;
  f_VTI_readDataset_int(vti.c_str(), (int)vti.size(), "material", 8, material_at.data(), &n_read);       // 1-based IDs
  if (n_read != n) return die("f_VTI_readDataset_int");
  const int nx = cells[0], ny = cells[1], nz = cells[2];
  std::vector<double> ip((size_t)3 * n), nodes((size_t)3 * (nx + 1) * (ny + 1) * (nz + 1));
  for (int k = 0, i = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int x = 0; x < nx; x++, i++) {
    ip[(size_t)3 * i] = size[0] / nx * (x + 0.5); ip[(size_t)3 * i + 1] = size[1] / ny * (j + 0.5); ip[(size_t)3 * i + 2] = size[2] / nz * (k + 0.5);
  }
  for (int k = 0, i = 0; k <= nz; k++) for (int j = 0; j <= ny; j++) for (int x = 0; x <= nx; x++, i++) {
    nodes[(size_t)3 * i] = size[0] / nx * x; nodes[(size_t)3 * i + 1] = size[1] / ny * j; nodes[(size_t)3 * i + 2] = size[2] / nz * k;
  }
  int n_nodes = (nx + 1) * (ny + 1) * (nz + 1), shared = (nx + 1) * (ny + 1) * nz;
  // ---- main.cpp:80-86
  f_materialpoint_initBase(material.c_str(), (int)material.size(), load.c_str(), (int)load.size(), geometry.c_str(), (int)geometry.size(),
                           numerics.c_str(), (int)numerics.size());
  f_discretization_init(material_at.data(), &n, ip.data(), &n, nodes.data(), &n_nodes, &shared);
  f_materialpoint_initDamask();
  dmk_ctx* ctx = dmk_bound_context();
  if (!ctx) return die("f_materialpoint_initDamask");
  // ---- Config::parse_load / parse_numerics
  std::vector<dmk_load_step> steps(64);
  int n_steps = 0;
  char scheme[32];
  if (dmk_load_from_yaml(load.c_str(), steps.data(), (int)steps.size(), &n_steps, scheme, (int)sizeof scheme)) return die("dmk_load_from_yaml");
  dmk_grid_numerics gn;
  if (dmk_grid_numerics_from_yaml(numerics.c_str(), &gn)) return die("dmk_grid_numerics_from_yaml");
  // ---- DamaskGrid::init (damask_grid.cpp:8-28): the mechanical solver the load file names
  const std::string s = scheme;
  if (s == "spectral_basic") { if (dmk_mech_init(ctx, nullptr)) return die("dmk_mech_init"); }
  else if (s == "spectral_polarisation") { if (dmk_mech_init_polarisation(ctx, gn.alpha, gn.beta, gn.eps_abs_curl, gn.eps_rel_curl)) return die("dmk_mech_init_polarisation"); }
  else { fprintf(stderr, "native_driver: unknown mechanical solver %s\n", scheme); return 2; }
  // ---- DamaskGrid::loop_over_loadcase
  dmk_loop_ops ops = {ctx, op_forward, op_point_forward, op_solve, report};
  int n_records = 0;
  if (dmk_loop_over_loadcase(ctx, steps.data(), n_steps, gn.N_cutback_max, max_increments, &ops, nullptr, 0, &n_records)) return die("dmk_loop_over_loadcase");
  printf("NATIVE_DRIVER_OK %d increments\n", n_records);
  f_materialpoint_initBase(nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0);
  return 0;
}

// The reference's load-case driver above the mechanical solver, in C++ like the original:
//   Config::parse_load                 /root/reference/src/grid/config.cpp:247-358 (load file -> Config::LoadStep, config.h:74-96)
//   DamaskGrid::loop_over_loadcase     /root/reference/src/grid/damask_grid.cpp:82-196
// with the cut-back bookkeeping of upstream DAMASK (DAMASK_grid.f90; SURVEY 0.1 D10: the level drops by one after an increment that
// needed cut-backs, the step fraction doubles on every cut-back).  The loop talks to the solver through a table of four operations;
// the default table drives the context's device-resident solver (dmk_mech_forward, dmk_materialpoint_forward,
// dmk_mech_calculate_solution), a caller-supplied one lets the reference's own MechSolverBasic -- or a scripted stand-in, which is
// how tests/test_loadcase_loop_cpu.py checks the bookkeeping against the Python host's mech.loop_over_loadcase -- sit underneath.
// Pure host code.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dmk_grid.h"
#include "dmk_common.cuh"
#include "yaml_min.h"

extern "C" void dmk_qu2om(const double q[4], double om[9]);

namespace {

[[noreturn]] void fail(const std::string& what) { throw std::runtime_error(what); }

const YamlNode& need_key(const YamlNode& m, const char* key, const char* message) {
  const YamlNode* n = m.get(key);
  if (!n) fail(message);
  return *n;
}

// parse_boundary_condition_matrices (config.cpp:198-245): a 3 x 3 list of numbers or 'x' (component not controlled by this one)
void parse_bc(const YamlNode& rows, const std::string& name, double values[9], int mask[9]) {
  if (!rows.is_seq() || rows.seq.size() != 3) fail("load step: " + name + " must have 3 rows");
  for (int i = 0; i < 3; i++) {
    const YamlNode& row = *rows.seq[(size_t)i];
    if (!row.is_seq() || row.seq.size() != 3) fail("load step: " + name + " must have 3 columns in every row");
    for (int j = 0; j < 3; j++) {
      const YamlNode& v = *row.seq[(size_t)j];
      if (v.is_scalar() && (v.scalar == "x" || v.scalar == "X")) { mask[3 * i + j] = 1; values[3 * i + j] = 0.0; }
      else { mask[3 * i + j] = 0; values[3 * i + j] = v.as_double(); }
    }
  }
}

void copy_name(char* dst, size_t cap, const std::string& s) {
  if (s.size() + 1 > cap) fail("name too long: " + s);
  memset(dst, 0, cap);
  memcpy(dst, s.c_str(), s.size());
}

// ---- the default operations: the context's own solver
int ctx_forward(void* user, int cut_back, int guess, double dt, double dt_old, double t_remaining, const dmk_load_step* st) {
  return dmk_mech_forward((dmk_ctx*)user, cut_back, guess, dt, dt_old, t_remaining, st->deformation_type, st->deformation_values,
                          st->deformation_mask, st->stress_type, st->stress_values, st->stress_mask, st->R);
}
int ctx_point_forward(void* user) { return dmk_materialpoint_forward((dmk_ctx*)user); }
int ctx_solve(void* user, dmk_solution_state* out) { return dmk_mech_calculate_solution((dmk_ctx*)user, out); }

}  // namespace

extern "C" int dmk_load_from_yaml(const char* load_yaml_path, dmk_load_step* steps, int max_steps, int* n_steps, char* mechanical_solver,
                                  int solver_name_capacity) {
  try {
    if (!load_yaml_path || !steps || !n_steps || max_steps < 1) fail("dmk_load_from_yaml: bad argument");
    auto root = yaml_parse_file(load_yaml_path);
    if (!root->is_map()) fail("load file contains no data");
    if (mechanical_solver && solver_name_capacity > 0) {
      // damask_grid.cpp:8-12 reads solver/mechanical; absent: the basic scheme
      const YamlNode* solver = root->get("solver");
      const YamlNode* mech = solver ? solver->get("mechanical") : nullptr;
      copy_name(mechanical_solver, (size_t)solver_name_capacity, mech ? mech->scalar : std::string("spectral_basic"));
    }
    const YamlNode& list = need_key(*root, "loadstep", "Load definition is missing 'loadstep' entry");
    if (!list.is_seq() || list.seq.empty()) fail("'loadstep' must be a non-empty list");
    if ((int)list.seq.size() > max_steps) fail("more load steps than the caller's buffer holds");
    int k = 0;
    for (const auto& item : list.seq) {
      if (!item->is_map()) fail("Item in loadstep list contains no data");
      dmk_load_step& st = steps[k];
      memset(&st, 0, sizeof st);
      const YamlNode& disc = need_key(*item, "discretization", "loadstep entry needs to define 'discretization'");
      st.t = need_key(disc, "t", "'t' is missing in loadstep.discretization").as_double();
      st.N = (int)need_key(disc, "N", "'N' is missing in loadstep.discretization").as_int();
      st.r = disc.has("r") ? disc.get("r")->as_double() : 1.0;
      if (st.N < 1 || !(st.t > 0.0) || !(st.r > 0.0)) fail("load step " + std::to_string(k) + ": t, N and r must be positive");
      const YamlNode& bc = need_key(*item, "boundary_conditions", "loadstep entry needs to define 'boundary_conditions'");
      const YamlNode& mech = need_key(bc, "mechanical", "loadstep.boundary_conditions entry needs to define 'mechanical'");
      for (const char* name : {"F", "dot_F", "L"}) {
        if (const YamlNode* n = mech.get(name)) {
          if (st.deformation_type[0]) fail("Only one deformation specification allowed, found more than one in loadstep at index " + std::to_string(k));
          parse_bc(*n, name, st.deformation_values, st.deformation_mask);
          copy_name(st.deformation_type, sizeof st.deformation_type, name);
        }
      }
      if (!st.deformation_type[0]) fail("Deformation specification required for each loadstep, not specified for loadstep at index " + std::to_string(k));
      for (const char* name : {"P", "dot_P"}) {
        if (const YamlNode* n = mech.get(name)) {
          if (st.stress_type[0]) fail("Only one stress specification allowed, found more than one in loadstep at index " + std::to_string(k));
          parse_bc(*n, name, st.stress_values, st.stress_mask);
          copy_name(st.stress_type, sizeof st.stress_type, name);
        }
      }
      if (!st.stress_type[0]) {
        // upstream DAMASK_grid.f90: no stress entry = no component is stress controlled (mask all 'x'); the deformation must then
        // prescribe all nine (next check).  Config::parse_load (config.cpp:309-311) throws here -- on the reference's own example
        // decks examples/grid/shearXY.yaml and shearZX.yaml
        copy_name(st.stress_type, sizeof st.stress_type, "P");
        for (int c = 0; c < 9; c++) { st.stress_mask[c] = 1; st.stress_values[c] = 0.0; }
      }
      for (int c = 0; c < 9; c++)
        if (st.deformation_mask[c] == st.stress_mask[c]) fail("load step " + std::to_string(k) + ": each component must be controlled by exactly one of deformation and stress");
      st.estimate_rate = item->has("estimate_rate") ? (int)item->get("estimate_rate")->as_bool() : 1;      // config.h:92-95 defaults
      st.f_out = item->has("f_out") ? (int)item->get("f_out")->as_int() : 1;
      st.f_restart = item->has("f_restart") ? (int)item->get("f_restart")->as_int() : 0;
      const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      memcpy(st.R, I3, sizeof I3);
      if (const YamlNode* R = mech.get("R")) {
        const std::vector<double> q = R->as_doubles();
        if (q.size() != 4) fail("R entry in loadstep must have exactly 4 elements");
        dmk_qu2om(q.data(), st.R);
      }
      k++;
    }
    *n_steps = k;
    return DMK_OK;
  } catch (const std::exception& e) {
    dmk_set_error("%s", e.what());
    return DMK_ERR_ARG;
  }
}

extern "C" int dmk_loop_over_loadcase(dmk_ctx* ctx, const dmk_load_step* steps, int n_steps, int max_cut_back, int max_increments,
                                      const dmk_loop_ops* ops, dmk_increment_record* records, int max_records, int* n_records) {
  if (!steps || n_steps < 1 || max_cut_back < 0) { dmk_set_error("dmk_loop_over_loadcase: bad argument"); return DMK_ERR_ARG; }
  dmk_loop_ops own;
  if (!ops) {
    if (!ctx) { dmk_set_error("dmk_loop_over_loadcase: neither a context nor a table of operations"); return DMK_ERR_ARG; }
    own.user = ctx; own.forward = ctx_forward; own.materialpoint_forward = ctx_point_forward; own.calculate_solution = ctx_solve;
    own.result = nullptr;
    ops = &own;
  }
  if (!ops->forward || !ops->materialpoint_forward || !ops->calculate_solution) { dmk_set_error("dmk_loop_over_loadcase: incomplete table of operations"); return DMK_ERR_ARG; }
  if (n_records) *n_records = 0;
  int rc;
  bool cut_back = false;
  const int sub_step_factor = 2;
  int cut_back_level = 0, total_inc = 0, n_rec = 0;
  double t = 0.0, dt = 1.0, dt_prev = 0.0;
  for (int l = 0; l < n_steps; l++) {
    const dmk_load_step& st = steps[l];
    const double t0 = t;
    bool guess = st.estimate_rate && l > 0;
    for (int inc = 0; inc < st.N; inc++) {
      if (max_increments > 0 && total_inc >= max_increments) { if (n_records) *n_records = n_rec; return DMK_OK; }
      total_inc++;
      dt_prev = dt;
      if (fabs(st.r - 1.0) < 1.0e-9) dt = st.t / st.N;                                              // linear time scale
      else dt = st.t * (pow(st.r, inc) - pow(st.r, inc + 1)) / (1.0 - pow(st.r, st.N));             // geometric one
      dt = dt * pow((double)sub_step_factor, -cut_back_level);                                      // what a previous cut-back left
      long step_fraction = 0;
      dmk_solution_state sol;
      memset(&sol, 0, sizeof sol);
      while (step_fraction < lround(pow((double)sub_step_factor, cut_back_level))) {
        const double t_remaining = st.t + t0 - t;
        t += dt;
        step_fraction++;
        if ((rc = ops->forward(ops->user, cut_back ? 1 : 0, guess ? 1 : 0, dt, dt_prev, t_remaining, &st))) return rc;
        if (!cut_back) { if ((rc = ops->materialpoint_forward(ops->user))) return rc; }             // f_materialpoint_forward
        memset(&sol, 0, sizeof sol);
        if ((rc = ops->calculate_solution(ops->user, &sol))) return rc;
        const bool found = sol.converged && sol.stag_converged && !sol.terminally_ill;
        if (found) { dt_prev = dt; cut_back = false; guess = true; }
        else if (cut_back_level < max_cut_back) {
          cut_back = true;
          step_fraction = (step_fraction - 1) * sub_step_factor;
          cut_back_level++;
          t -= dt;
          dt = dt / sub_step_factor;
        } else { dmk_set_error("no solution found for increment %d", total_inc); return DMK_ERR_STATE; }
      }
      cut_back_level = cut_back_level > 0 ? cut_back_level - 1 : 0;
      dmk_increment_record rec;
      rec.inc = total_inc; rec.step = l; rec.inc_in_step = inc + 1; rec.f_out = st.f_out; rec.t = t;
      rec.converged = sol.converged; rec.iterations = sol.required_iterations;
      if (records && n_rec < max_records) records[n_rec] = rec;
      n_rec++;
      if (n_records) *n_records = n_rec < max_records || !records ? n_rec : max_records;
      // damask_grid.cpp:165-168: result output of the increments the step's f_out selects
      if (ops->result && rec.inc_in_step % (st.f_out > 0 ? st.f_out : 1) == 0) { if ((rc = ops->result(ops->user, &rec))) return rc; }
    }
  }
  return DMK_OK;
}

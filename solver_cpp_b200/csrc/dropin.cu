// The Fortran-named symbols the reference's C++ driver binds (/root/reference/src/grid/spectral.h:16-23,
// damask_grid.h:15-18, utilities_fortran.h:8-15), re-implemented on top of the device path.  They act on the context made
// current with dmk_bind_global() and exchange data through pinned HOST mirrors of F0, F, P, dPdF in the reference's
// memory order ([3][3][N] / [3][3][3][3][N], first index fastest), exactly like the Fortran module arrays they replace
// (/root/reference/src/homogenization.f90:491-510).  Semantics: homogenization.f90:219-324, materialpoint.f90:208-234,
// math.f90:428,599,819,849, rotations.f90:331-359.  Errors cannot be returned through these signatures: they set the
// terminally-ill flag and leave the message in dmk_last_error().
#include <cstring>
#include <cmath>
#include <vector>
#include "ctx.h"
#include "constitutive.cuh"
#include "reduce.cuh"

static dmk_ctx* g_ctx = nullptr;
static double* g_dF = nullptr;     // device staging: F records
static double* g_dP = nullptr;     // device staging: P records
static double* g_dT = nullptr;     // device staging: dPdF records (81 per cell)
static long long g_n = 0;

// cell records -> the per-point component planes; under isostrain homogenisation every constituent of a cell (point k * n + cell)
// receives the cell's tensor (homogenization_mechanical_isostrain.f90:36-45)
__global__ void dmk_records_to_soa_kernel(const double* __restrict__ rec, double* __restrict__ soa, long long n, int n_const) {
  const long long stride = n * n_const;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
      const double x = rec[i * 9 + a + 3 * b];
      for (int k = 0; k < n_const; k++) soa[(3 * a + b) * stride + k * n + i] = x;
    }
}

extern "C" int dmk_bind_global(dmk_ctx* c) {
  if (g_dF) { cudaFree(g_dF); cudaFree(g_dP); cudaFree(g_dT); g_dF = g_dP = g_dT = nullptr; }
  g_ctx = c;
  if (!c) return DMK_OK;
  if (!c->pts) { dmk_set_error("dmk_bind_global: material not initialised"); return DMK_ERR_STATE; }
  g_n = c->n_local;
  const size_t n = (size_t)g_n;
  if (!c->m_F) {
    DMK_CUDA_CHECK(cudaMallocHost((void**)&c->m_F0, 9 * n * sizeof(double)));
    DMK_CUDA_CHECK(cudaMallocHost((void**)&c->m_F, 9 * n * sizeof(double)));
    DMK_CUDA_CHECK(cudaMallocHost((void**)&c->m_P, 9 * n * sizeof(double)));
    DMK_CUDA_CHECK(cudaMallocHost((void**)&c->m_dPdF, 81 * n * sizeof(double)));
    DMK_CUDA_CHECK(cudaMallocHost((void**)&c->m_ill, sizeof(int)));
    for (size_t i = 0; i < n; i++) for (int k = 0; k < 9; k++) { c->m_F0[9 * i + k] = c->m_F[9 * i + k] = (k % 4 == 0) ? 1.0 : 0.0; c->m_P[9 * i + k] = 0.0; }
    memset(c->m_dPdF, 0, 81 * n * sizeof(double));
    *c->m_ill = 0;
  }
  DMK_CUDA_CHECK(cudaMalloc((void**)&g_dF, 9 * n * sizeof(double)));
  DMK_CUDA_CHECK(cudaMalloc((void**)&g_dP, 9 * n * sizeof(double)));
  DMK_CUDA_CHECK(cudaMalloc((void**)&g_dT, 81 * n * sizeof(double)));
  return DMK_OK;
}

void dmk_unbind_if_current(dmk_ctx* c) { if (g_ctx == c) dmk_bind_global(nullptr); }

static void fail(const char* what) {
  if (g_ctx && g_ctx->m_ill) *g_ctx->m_ill = 1;
  (void)what;
}

extern "C" void f_homogenization_fetch_tensor_pointers(double** F0, double** F, double** P, double** dPdF, void** terminally_ill) {
  if (!g_ctx || !g_ctx->m_F) { dmk_set_error("f_homogenization_fetch_tensor_pointers: no bound context (dmk_bind_global)"); return; }
  if (F0) *F0 = g_ctx->m_F0; if (F) *F = g_ctx->m_F; if (P) *P = g_ctx->m_P; if (dPdF) *dPdF = g_ctx->m_dPdF;
  if (terminally_ill) *terminally_ill = g_ctx->m_ill;
}

extern "C" void f_homogenization_mechanical_response(double* delta_t, int* cell_start, int* cell_end) {
  dmk_ctx* c = g_ctx;
  if (!c || !c->m_F) { dmk_set_error("f_homogenization_mechanical_response: no bound context"); return; }
  if (*cell_start != 1 || *cell_end != (int)g_n) { dmk_set_error("device path integrates the whole slab: cells must be 1..%lld", g_n); fail("range"); return; }
  const size_t bytes = 9 * (size_t)g_n * sizeof(double);
  // the caller owns F0 and F (mech_solver_basic.cpp:196-206): F0 is re-read every call, like the Fortran module array
  if (cudaMemcpyAsync(g_dF, c->m_F0, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { fail("h2d"); return; }
  dmk_records_to_soa_kernel<<<148 * 4, 256, 0, c->stream>>>(g_dF, c->pts->F0, g_n, c->pts->n_const);
  c->launches++;
  if (cudaMemcpyAsync(g_dF, c->m_F, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { fail("h2d"); return; }
  if (dmk_points_constitutive(c, g_dF, g_dP, *delta_t)) { fail("kernel"); return; }
  int ill = 0;
  if (dmk_terminally_ill(c, &ill, 1)) { fail("flag"); return; }
  if (ill) *c->m_ill = 1;
}

extern "C" void f_homogenization_thermal_response(double*, int*, int*) {}

extern "C" void f_homogenization_mechanical_response2(double* delta_t, int* FEsolving_execIP, int* FEsolving_execElem) {
  dmk_ctx* c = g_ctx;
  if (!c || !c->m_F) { dmk_set_error("f_homogenization_mechanical_response2: no bound context"); return; }
  (void)FEsolving_execIP; (void)FEsolving_execElem;
  double* d_sum = c->d_scratch + DMK_SCRATCH_TMP;
  if (dmk_cells_tangent(c, *delta_t, d_sum, g_dT)) { fail("tangent"); return; }           // homogenised over the constituents (mechanical_homogenize)
  if (cudaMemcpyAsync(c->m_P, g_dP, 9 * (size_t)g_n * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { fail("d2h"); return; }
  if (cudaMemcpyAsync(c->m_dPdF, g_dT, 81 * (size_t)g_n * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { fail("d2h"); return; }
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) fail("sync");
}

extern "C" void f_materialpoint_forward(void) {
  if (!g_ctx) { dmk_set_error("f_materialpoint_forward: no bound context"); return; }
  if (dmk_materialpoint_forward(g_ctx)) fail("forward");
}
extern "C" void f_materialpoint_result(int*, double*) {}      // result output (DADF5) is outside the hot path

// ------------------------------------------------------------------------------------------------ pure helpers (host)
extern "C" void f_math_invert(double* InvA, int* err, double* A, int* n_) {
  const int n = *n_;
  std::vector<double> M((size_t)n * n), I((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) M[(size_t)i * n + j] = A[i + (size_t)n * j]; I[(size_t)i * n + i] = 1.0; }
  *err = 0;
  for (int k = 0; k < n; k++) {
    int p = k; double mx = fabs(M[(size_t)k * n + k]);
    for (int i = k + 1; i < n; i++) if (fabs(M[(size_t)i * n + k]) > mx) { mx = fabs(M[(size_t)i * n + k]); p = i; }
    if (mx == 0.0) { *err = 1; for (int q = 0; q < n * n; q++) InvA[q] = 0.0; return; }
    if (p != k) for (int j = 0; j < n; j++) { std::swap(M[(size_t)k * n + j], M[(size_t)p * n + j]); std::swap(I[(size_t)k * n + j], I[(size_t)p * n + j]); }
    const double d = M[(size_t)k * n + k];
    for (int j = 0; j < n; j++) { M[(size_t)k * n + j] /= d; I[(size_t)k * n + j] /= d; }
    for (int i = 0; i < n; i++) if (i != k) {
      const double f = M[(size_t)i * n + k];
      if (f != 0.0) for (int j = 0; j < n; j++) { M[(size_t)i * n + j] -= f * M[(size_t)k * n + j]; I[(size_t)i * n + j] -= f * I[(size_t)k * n + j]; }
    }
  }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) InvA[i + (size_t)n * j] = I[(size_t)i * n + j];
}
// MAPPLAIN pair order (math.f90:73-84): pair index a = 3 i + j (row-major)
extern "C" void f_math_3333to99(double* m3333, double* m99) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++)
    m99[(3 * i + j) + 9 * (3 * k + l)] = m3333[i + 3 * j + 9 * k + 27 * l];
}
extern "C" void f_math_99to3333(double* m99, double* m3333) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++)
    m3333[i + 3 * j + 9 * k + 27 * l] = m99[(3 * i + j) + 9 * (3 * k + l)];
}
extern "C" void f_math_mul3333xx33(double* A, double* B, double* res) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double s = 0.0;
    for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) s += A[i + 3 * j + 9 * k + 27 * l] * B[k + 3 * l];
    res[i + 3 * j] = s;
  }
}
// rotations.f90:288-359: passive R T R^T (default), active R^T T R when *active == 1
extern "C" void f_rotate_tensor2(double* qu, double* T, double* rotated, int* active) {
  double R[9], Tr[9], t[9], o[9];
  dmk_qu2om(qu, R);
  f2c33(T, Tr);
  const bool act = active && *active == 1;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += (act ? R[3 * k + i] : R[3 * i + k]) * Tr[3 * k + j]; t[3 * i + j] = s; }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += t[3 * i + k] * (act ? R[3 * k + j] : R[3 * j + k]); o[3 * i + j] = s; }
  c2f33(o, rotated);
}
extern "C" void f_rotate_tensor4(double* qu, double* T, double* rotated, int* active) {
  double R[9], Q[9], Tr[81], o[81];
  dmk_qu2om(qu, R);
  const bool act = active && *active == 1;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q[3 * i + j] = act ? R[3 * j + i] : R[3 * i + j];
  f2c3333(T, Tr);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) {
    double s = 0;
    for (int m = 0; m < 3; m++) for (int n = 0; n < 3; n++) for (int p = 0; p < 3; p++) for (int q = 0; q < 3; q++)
      s += Q[3 * i + m] * Q[3 * j + n] * Q[3 * k + p] * Q[3 * l + q] * Tr[((m * 3 + n) * 3 + p) * 3 + q];
    o[((i * 3 + j) * 3 + k) * 3 + l] = s;
  }
  c2f3333(o, rotated);
}

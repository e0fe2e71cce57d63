// Tuning harness (not part of the library): times template variants of the register-resident FFT / Gamma kernels
// (tile width KT, threads per CTA, CTAs per SM, stage split) on one GPU with CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o tools/_build/fft_tune tools/fft_tune.cu
//   tools/_build/fft_tune 256
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <vector>
#include "../solver_cpp_b200/csrc/fft_kernels.cu"
#include "../solver_cpp_b200/csrc/fft_fast.cu"
#include "../solver_cpp_b200/csrc/gamma_kernels.cu"
#include "../solver_cpp_b200/csrc/gamma_fast.cu"

void dmk_set_error(const char* fmt, ...) { va_list a; va_start(a, fmt); vfprintf(stderr, fmt, a); va_end(a); fputc('\n', stderr); }

static double* d_real; static double2* d_hat; static DmkFftPlan pl; static int n, nxr;
static DmkGammaParams gp;
static cudaEvent_t e0, e1;
static const char* g_filter = nullptr;
static DmkPeerStore g_no_peers = {};        // local stores (no fused exchange)

template <class F> static void timeit(const char* name, F f) {
  if (g_filter && !strstr(name, g_filter)) return;
  for (int i = 0; i < 2; i++) f();
  cudaDeviceSynchronize();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { printf("%-46s FAILED %s\n", name, cudaGetErrorString(err)); return; }
  const int reps = 5;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  const double bytes = 144.0 * n * (double)n * n;
  printf("%-46s %7.3f ms  %6.0f GB/s  %5.1f %%\n", name, ms, bytes / ms / 1e6, bytes / ms / 1e6 / 6535.0 * 100);
  fflush(stdout);
}

template <class PL, int KT, int NT, int MINB> static void line_y(const char* tag) {
  DmkLineMap m; m.split = n; m.hi_stride = 0; m.o_stride = (long long)n * nxr; m.e_stride = nxr;
  char name[96]; snprintf(name, 96, "y  <%d,%d,%d> KT%d NT%d MINB%d %s", PL::R1, PL::R2, PL::R3, KT, NT, MINB, tag);
  timeit(name, [&] { launch_line<PL, KT, NT, MINB>(false, d_hat, d_hat, pl, m, m, nxr, n, 0, g_no_peers); });
}
template <class PL, int KT, int NT, int MINB> static void gamma_z(const char* tag) {
  DmkLineMap m; m.split = n; m.hi_stride = 0; m.o_stride = nxr; m.e_stride = (long long)n * nxr;
  char name[96]; snprintf(name, 96, "gz <%d,%d,%d> KT%d NT%d MINB%d %s", PL::R1, PL::R2, PL::R3, KT, NT, MINB, tag);
  timeit(name, [&] { launch_gamma<PL, KT, NT, MINB>(d_hat, pl, m, nxr, n, gp, 0, true, g_no_peers); });
}
template <class PL, int NP, int NT, int MINB> static void xr2c(const char* tag) {
  char name[96]; snprintf(name, 96, "xf <%d,%d,%d> NP%d NT%d MINB%d %s", PL::R1, PL::R2, PL::R3, NP, NT, MINB, tag);
  timeit(name, [&] { launch_xr2c<PL, NP, NT, MINB>(d_real, d_hat, pl, n * n, 0); });
}
template <class PL, int NP, int NT, int MINB> static void xc2r(const char* tag) {
  char name[96]; snprintf(name, 96, "xb <%d,%d,%d> NP%d NT%d MINB%d %s", PL::R1, PL::R2, PL::R3, NP, NT, MINB, tag);
  timeit(name, [&] { launch_xc2r<PL, NP, NT, MINB>(d_hat, d_real, pl, n * n, 1.0, 0); });
}

int main(int argc, char** argv) {
  n = argc > 1 ? atoi(argv[1]) : 256;
  g_filter = argc > 2 ? argv[2] : nullptr;
  nxr = n / 2 + 1;
  const size_t nreal = (size_t)n * n * n * 9, nhat = (size_t)n * n * nxr * 9;
  cudaMalloc(&d_real, nreal * sizeof(double)); cudaMalloc(&d_hat, nhat * sizeof(double2));
  std::vector<double> h(nreal);
  unsigned long long sd = 88172645463325252ULL;
  for (size_t i = 0; i < nreal; i++) { sd ^= sd << 13; sd ^= sd >> 7; sd ^= sd << 17; h[i] = (double)(sd >> 11) / 9007199254740992.0 - 0.5; }
  cudaMemcpy(d_real, h.data(), nreal * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemset(d_hat, 0, nhat * sizeof(double2));
  dmk_fft_plan_create(&pl, n);
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  // Gamma parameters: continuous derivative of a unit cube, isotropic-ish stiffness
  std::vector<double2> xi(n), one(n, make_double2(1.0, 0.0)); std::vector<unsigned char> nyq(n, 0);
  for (int k = 0; k < n; k++) { int f = k <= n / 2 ? k : k - n; xi[k] = make_double2(0.0, 6.283185307179586 * f); if (2 * k == n) { xi[k].y = 0; nyq[k] = 1; } }
  double2 *dxi, *done_; unsigned char* dnyq; double *dC, *ddc, *ddiv;
  cudaMalloc(&dxi, n * sizeof(double2)); cudaMalloc(&done_, n * sizeof(double2)); cudaMalloc(&dnyq, n);
  cudaMemcpy(dxi, xi.data(), n * sizeof(double2), cudaMemcpyHostToDevice); cudaMemcpy(done_, one.data(), n * sizeof(double2), cudaMemcpyHostToDevice);
  cudaMemcpy(dnyq, nyq.data(), n, cudaMemcpyHostToDevice);
  std::vector<double> C(81, 0.0);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++)
    C[((i * 3 + j) * 3 + k) * 3 + l] = 60e9 * (i == j && k == l) + 28e9 * ((i == k && j == l) + (i == l && j == k));
  C.resize(81 + 54);
  { static const int vn[6] = {0, 1, 2, 1, 0, 0}, vo[6] = {0, 1, 2, 2, 2, 1};
    for (int l = 0; l < 3; l++) for (int m = 0; m < 3; m++) for (int v = 0; v < 6; v++) {
      double s = C[((l * 3 + vn[v]) * 3 + m) * 3 + vo[v]]; if (vn[v] != vo[v]) s += C[((l * 3 + vo[v]) * 3 + m) * 3 + vn[v]];
      C[81 + (l * 3 + m) * 6 + v] = s; } }
  cudaMalloc(&dC, 135 * 8); cudaMemcpy(dC, C.data(), 135 * 8, cudaMemcpyHostToDevice);
  cudaMalloc(&ddc, 9 * 8); cudaMemset(ddc, 0, 72); cudaMalloc(&ddiv, 8); cudaMemset(ddiv, 0, 8);
  gp.C_ref = dC; gp.xi_x = gp.xi_y = gp.xi_z = dxi; gp.xo_x = gp.xo_y = gp.xo_z = done_; gp.nyq_x = gp.nyq_y = gp.nyq_z = dnyq;
  gp.rescale[0] = gp.rescale[1] = gp.rescale[2] = 1.0; gp.nx = n; gp.ky_offset = 0; gp.dc_value = ddc; gp.div_accum = ddiv; { double* dpart; cudaMalloc(&dpart, (size_t)n * n * sizeof(double)); gp.div_partial = dpart; }

  if (n == 256) {
    typedef DmkFastPlan<16, 16, 1> A; typedef DmkFastPlan<8, 8, 4> Bp; typedef DmkFastPlan<4, 8, 8> Dp;
    xr2c<A, 1, 144, 3>(""); xr2c<A, 1, 144, 2>(""); xr2c<A, 1, 144, 4>(""); xr2c<A, 2, 288, 1>("");
    line_y<A, 1, 144, 3>(""); line_y<A, 1, 144, 2>(""); line_y<A, 1, 144, 4>(""); line_y<A, 2, 288, 1>(""); line_y<A, 2, 288, 2>("");
    xc2r<A, 1, 144, 4>(""); xc2r<A, 1, 144, 3>(""); xc2r<A, 1, 144, 2>(""); xc2r<A, 2, 288, 1>(""); xc2r<A, 2, 288, 2>(""); xc2r<Bp, 1, 144, 6>(""); xc2r<Bp, 1, 144, 5>(""); xc2r<Dp, 1, 144, 6>("");
    gamma_z<A, 2, 192, 2>(""); gamma_z<A, 2, 160, 2>(""); gamma_z<A, 2, 160, 2>(""); gamma_z<A, 2, 288, 1>(""); gamma_z<A, 2, 288, 2>(""); gamma_z<A, 2, 224, 2>("");
    gamma_z<A, 1, 160, 2>(""); gamma_z<A, 1, 160, 3>(""); gamma_z<A, 1, 96, 3>(""); gamma_z<A, 1, 96, 4>(""); gamma_z<A, 1, 64, 4>("");
    gamma_z<Bp, 2, 192, 2>(""); gamma_z<Bp, 1, 160, 4>("");
  } else if (n == 512) {
    typedef DmkFastPlan<8, 8, 8> A; typedef DmkFastPlan<16, 8, 4> Bp; typedef DmkFastPlan<16, 16, 2> Cp; typedef DmkFastPlan<4, 8, 16> Dp;
    xr2c<A, 1, 288, 2>(""); xr2c<A, 1, 144, 4>(""); xr2c<A, 1, 144, 6>(""); xr2c<A, 1, 192, 4>(""); xr2c<A, 1, 192, 3>(""); xr2c<Cp, 1, 144, 3>(""); xr2c<Cp, 1, 288, 2>(""); xr2c<Bp, 1, 144, 3>("");
    line_y<A, 1, 288, 2>(""); line_y<A, 1, 144, 4>(""); line_y<A, 1, 144, 6>(""); line_y<A, 1, 192, 4>(""); line_y<A, 1, 192, 3>(""); line_y<Cp, 1, 144, 3>(""); line_y<Cp, 1, 288, 1>(""); line_y<Bp, 1, 144, 3>("");
    xc2r<A, 1, 288, 2>(""); xc2r<A, 1, 144, 4>(""); xc2r<A, 1, 144, 6>(""); xc2r<A, 1, 192, 4>(""); xc2r<Cp, 1, 144, 3>(""); xc2r<Bp, 1, 144, 3>(""); xc2r<Dp, 1, 144, 3>("");
    gamma_z<A, 1, 288, 2>(""); gamma_z<A, 1, 160, 2>(""); gamma_z<A, 1, 160, 3>(""); gamma_z<A, 1, 192, 2>(""); gamma_z<A, 1, 96, 4>(""); gamma_z<Cp, 1, 160, 2>(""); gamma_z<Cp, 1, 192, 2>(""); gamma_z<Bp, 1, 160, 2>(""); gamma_z<Bp, 1, 192, 2>("");
  } else if (n == 128) {
    typedef DmkFastPlan<16, 8, 1> A; typedef DmkFastPlan<8, 4, 4> Bp;
    xr2c<A, 4, 288, 1>(""); xr2c<A, 2, 144, 3>(""); xr2c<A, 2, 144, 4>(""); xr2c<A, 1, 72, 6>("");
    line_y<A, 4, 288, 1>(""); line_y<A, 2, 144, 3>(""); line_y<A, 2, 144, 4>(""); line_y<A, 1, 72, 6>("");
    xc2r<A, 4, 288, 1>(""); xc2r<A, 2, 144, 3>(""); xc2r<A, 2, 144, 4>(""); xc2r<Bp, 2, 144, 6>("");
    gamma_z<A, 4, 288, 1>(""); gamma_z<A, 4, 192, 2>(""); gamma_z<A, 2, 160, 2>(""); gamma_z<A, 2, 96, 3>("");
  }
  return 0;
}

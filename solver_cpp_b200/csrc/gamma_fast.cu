// Fused z pass for power-of-two nz on the register-resident FFT stages (fft_fast.cuh):
//   HBM -> [first forward stage in registers] -> shared -> [last forward stage] -> divergence RMS + Gamma contraction on
//   the digit-reversed line in shared memory -> [adjoint of the last stage] -> [adjoint of the first stage] -> HBM.
// Replaces MechBase::calculate_divergence_rms (/root/reference/src/grid/mech_base.cpp:302-398) and
// MechBase::gamma_convolution (mech_base.cpp:400-477); operator definition and per-k-point arithmetic: gamma_kernels.cu,
// gamma_point.cuh (shared with the shared-memory-only kernel, so both produce the same bits per k-point).
// One tile (N kz positions x KT kx records) per CTA; the half spectrum is read once and written once (144 B/point).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "fft_fast.cuh"
#include "gamma_point.cuh"

template <class PL, int KT, bool IMAG_XI, int NT, int MINB, bool REMOTE>
__global__ void __launch_bounds__(NT, MINB) dmk_fast_gamma_kernel(double2* __restrict__ data, const double2* __restrict__ tw,
                                                            DmkLineMap map, int n0, int tiles0, DmkGammaParams gp,
                                                            const __grid_constant__ DmkPeerStore ps) {
  constexpr int N = PL::N, R1 = PL::R1, R2 = PL::R2, R3 = PL::R3, B = KT * DMK_NC;
  static_assert(NT % 32 == 0, "whole warps: the divergence sum is reduced with full-warp shuffles");
  extern __shared__ double2 s[];
  __shared__ double sC[81];
  __shared__ double sred[NT / 32];
  const int o = blockIdx.x / tiles0, i0 = (blockIdx.x - o * tiles0) * KT;
  const int kt = min(KT, n0 - i0);
  const int bact = kt * DMK_NC;
  const bool nosplit = map.split >= N;
  // imaginary xi: the 54-entry symmetrised table behind the 81 stiffness entries (dmk_update_gamma)
  for (int i = threadIdx.x; i < (IMAG_XI ? 54 : 81); i += NT) sC[i] = gp.C_ref[IMAG_XI ? 81 + i : i];
  auto ldg = [&](int, int p, int col) { return data[dmk_line_rec_fast(map, nosplit, o, p, i0) * DMK_NC + col]; };
  auto lds = [&](int, int p, int col) { return s[p * B + col]; };
  auto sts = [&](int, int p, int col, double2 v) { s[p * B + col] = v; };
  auto stg = [&](int, int p, int col, double2 v) { dmk_line_store_ptr<REMOTE>(map, ps, data, nosplit, o, p, i0)[col] = v; };
  // ---- forward along z: natural order in HBM -> digit-reversed slots in shared memory
  fstage<N, R1, R2 * R3, B, false, false>(1, bact, tw, ldg, sts);
  __syncthreads();
  if constexpr (R3 > 1) {
    fstage<N, R2, R3, B, false, false>(1, bact, tw, lds, sts);
    __syncthreads();
    fstage<N, R3, 1, B, false, false>(1, bact, tw, lds, sts);
  } else {
    fstage<N, R2, 1, B, false, false>(1, bact, tw, lds, sts);
  }
  __syncthreads();
  // ---- per k-point: divergence contribution and Gamma contraction (slot -> kz = PL::freq(slot))
  const double resc[3] = {gp.rescale[0], gp.rescale[1], gp.rescale[2]};
  const int ky = o + gp.ky_offset;
  const int nxr = n0;
  double div_local = 0.0;
  for (int pnt = threadIdx.x; pnt < N * KT; pnt += NT) {
    const int slot = pnt / KT, j = pnt - slot * KT;
    if (j >= kt) continue;
    const int kz = PL::freq(slot);
    const int kx = i0 + j;
    double2 t[9];
    double2* ps = s + slot * B + j * DMK_NC;
#pragma unroll
    for (int c = 0; c < 9; c++) t[c] = ps[c];
    double2 xi[3];
    load_xi(gp, kx, ky, kz, xi);
    const bool is_dc = (kx == 0) && (ky == 0) && (kz == 0);
    const double wgt = ((kx == 0) ? 1.0 : 0.0) + ((kx == nxr - 1) ? 1.0 : 0.0) + ((kx > 0 && kx < nxr - 1) ? 2.0 : 0.0);
    double d;
    if (IMAG_XI) {
      const double q[3] = {xi[0].y, xi[1].y, xi[2].y};
      d = gamma_point_imag(t, sC, q, resc, is_dc, gp.dc_value);
    } else {
      d = gamma_point_cplx(t, sC, xi, resc, is_dc, gp.dc_value);
    }
    div_local += wgt * d;
#pragma unroll
    for (int c = 0; c < 9; c++) ps[c] = t[c];
  }
  __syncthreads();
  // ---- inverse along z: the adjoint network, digit-reversed slots -> natural order in HBM (unnormalised)
  if constexpr (R3 > 1) {
    fstage<N, R3, 1, B, true, true>(1, bact, tw, lds, sts);
    __syncthreads();
    fstage<N, R2, R3, B, true, true>(1, bact, tw, lds, sts);
  } else {
    fstage<N, R2, 1, B, true, true>(1, bact, tw, lds, sts);
  }
  __syncthreads();
  fstage<N, R1, R2 * R3, B, true, true>(1, bact, tw, lds, stg);
  // ---- block reduction of the divergence sum; the CTAs' sums are added up in a fixed order afterwards (dmk_sum_div_partials)
  for (int off = 16; off > 0; off >>= 1) div_local += __shfl_down_sync(0xffffffffu, div_local, off);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = div_local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) tot += sred[w];
    gp.div_partial[blockIdx.x] = tot;
  }
}

// ------------------------------------------------------------------------------------------------ pipelined variant
// Same arithmetic, tile shape and per-tile divergence sums as dmk_fast_gamma_kernel (bit-identical results), but a persistent
// CTA per SM that keeps the NEXT tile's loads in flight while it transforms the current one: the tile (N kz positions x KT kx
// records, 73.7 kB) arrives in one of two shared-memory buffers through bulk asynchronous copies (cp.async.bulk, one per kz
// row, completion counted in bytes on an mbarrier), issued as soon as the other buffer's tile has been started.  The kernel
// above has one tile in flight per CTA and loads nothing while it runs its five barrier-separated compute phases: 44 % of the
// HBM rate at 256^3 with 2 CTAs per SM (profiles/r01_summary.md); the pass executes ~700 FP64 instructions per k-point, which
// alone is 0.34 ms against the 0.37 ms the 144 B/point take at the measured copy rate, so loads and arithmetic must overlap.
__device__ __forceinline__ unsigned dmk_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dmk_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dmk_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void dmk_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dmk_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dmk_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DMK_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DMK_WAIT_DONE;\n"
      "bra DMK_WAIT_LOOP;\n"
      "DMK_WAIT_DONE:\n"
      "}\n" ::"r"(dmk_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void dmk_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dmk_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(dmk_smem_u32(bar)) : "memory");
}

template <class PL, int KT, bool IMAG_XI, int NT, bool REMOTE>
__global__ void __launch_bounds__(NT, 1) dmk_fast_gamma_pipe_kernel(double2* __restrict__ data, const double2* __restrict__ tw,
                                                                    DmkLineMap map, int n0, int tiles0, int n_tiles, DmkGammaParams gp,
                                                                    const __grid_constant__ DmkPeerStore ps) {
  constexpr int N = PL::N, R1 = PL::R1, R2 = PL::R2, R3 = PL::R3, B = KT * DMK_NC;
  static_assert(NT % 32 == 0, "whole warps: the divergence sum is reduced with full-warp shuffles");
  extern __shared__ __align__(128) double2 s_all[];       // two tile buffers
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ double sC[81];
  __shared__ double sred[NT / 32];
  const bool nosplit = map.split >= N;
  for (int i = threadIdx.x; i < (IMAG_XI ? 54 : 81); i += NT) sC[i] = gp.C_ref[IMAG_XI ? 81 + i : i];
  if (threadIdx.x == 0) {
    dmk_mbar_init(&bar[0], 1); dmk_mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // every thread issues the rows it owns; thread 0 arms the barrier with the tile's byte count (the phase cannot complete
  // before that arrival, whatever the order in which rows land)
  auto issue = [&](int tile, int b) {
    const int o = tile / tiles0, i0 = (tile - o * tiles0) * KT;
    const unsigned row_bytes = (unsigned)(min(KT, n0 - i0) * DMK_NC * sizeof(double2));
    if (threadIdx.x == 0) dmk_mbar_expect_tx(&bar[b], row_bytes * N);
    double2* dst = s_all + (size_t)b * N * B;
    for (int p = threadIdx.x; p < N; p += NT)
      dmk_bulk_g2s(dst + p * B, data + dmk_line_rec_fast(map, nosplit, o, p, i0) * DMK_NC, row_bytes, &bar[b]);
  };
  const double resc[3] = {gp.rescale[0], gp.rescale[1], gp.rescale[2]};
  const int nxr = n0;
  int tile = blockIdx.x;
  if (tile < n_tiles) issue(tile, 0);
  for (int it = 0; tile < n_tiles; tile += gridDim.x, it++) {
    const int b = it & 1;
    double2* s = s_all + (size_t)b * N * B;
    const int o = tile / tiles0, i0 = (tile - o * tiles0) * KT;
    const int kt = min(KT, n0 - i0);
    const int bact = kt * DMK_NC;
    // the other buffer's tile was finished behind the barrier that ended the previous trip: refill it
    if (tile + (int)gridDim.x < n_tiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(tile + gridDim.x, b ^ 1);
    }
    dmk_mbar_wait(&bar[b], (unsigned)((it >> 1) & 1));
    auto lds = [&](int, int p, int col) { return s[p * B + col]; };
    auto sts = [&](int, int p, int col, double2 v) { s[p * B + col] = v; };
    auto stg = [&](int, int p, int col, double2 v) { dmk_line_store_ptr<REMOTE>(map, ps, data, nosplit, o, p, i0)[col] = v; };
    // ---- forward along z, in place in shared memory: natural order -> digit-reversed slots
    fstage<N, R1, R2 * R3, B, false, false>(1, bact, tw, lds, sts);
    __syncthreads();
    if constexpr (R3 > 1) {
      fstage<N, R2, R3, B, false, false>(1, bact, tw, lds, sts);
      __syncthreads();
      fstage<N, R3, 1, B, false, false>(1, bact, tw, lds, sts);
    } else {
      fstage<N, R2, 1, B, false, false>(1, bact, tw, lds, sts);
    }
    __syncthreads();
    // ---- per k-point: divergence contribution and Gamma contraction (slot -> kz = PL::freq(slot))
    const int ky = o + gp.ky_offset;
    double div_local = 0.0;
    for (int pnt = threadIdx.x; pnt < N * KT; pnt += NT) {
      const int slot = pnt / KT, j = pnt - slot * KT;
      if (j >= kt) continue;
      const int kz = PL::freq(slot);
      const int kx = i0 + j;
      double2 t[9];
      double2* pp = s + slot * B + j * DMK_NC;
#pragma unroll
      for (int c = 0; c < 9; c++) t[c] = pp[c];
      double2 xi[3];
      load_xi(gp, kx, ky, kz, xi);
      const bool is_dc = (kx == 0) && (ky == 0) && (kz == 0);
      const double wgt = ((kx == 0) ? 1.0 : 0.0) + ((kx == nxr - 1) ? 1.0 : 0.0) + ((kx > 0 && kx < nxr - 1) ? 2.0 : 0.0);
      double d;
      if (IMAG_XI) {
        const double q[3] = {xi[0].y, xi[1].y, xi[2].y};
        d = gamma_point_imag(t, sC, q, resc, is_dc, gp.dc_value);
      } else {
        d = gamma_point_cplx(t, sC, xi, resc, is_dc, gp.dc_value);
      }
      div_local += wgt * d;
#pragma unroll
      for (int c = 0; c < 9; c++) pp[c] = t[c];
    }
    // the tile's divergence sum, reduced like the one-tile-per-CTA kernel does (same order: bit-identical partials)
    for (int off = 16; off > 0; off >>= 1) div_local += __shfl_down_sync(0xffffffffu, div_local, off);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = div_local;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < NT / 32; w++) tot += sred[w];
      gp.div_partial[tile] = tot;
    }
    // ---- inverse along z: the adjoint network, digit-reversed slots -> natural order in HBM (unnormalised)
    if constexpr (R3 > 1) {
      fstage<N, R3, 1, B, true, true>(1, bact, tw, lds, sts);
      __syncthreads();
      fstage<N, R2, R3, B, true, true>(1, bact, tw, lds, sts);
    } else {
      fstage<N, R2, 1, B, true, true>(1, bact, tw, lds, sts);
    }
    __syncthreads();
    fstage<N, R1, R2 * R3, B, true, true>(1, bact, tw, lds, stg);
    __syncthreads();                                   // every read of this buffer is done before the next trip refills it
  }
}

// DMK_GAMMA_PIPE=1 selects the pipelined kernel.  Off by default: measured SLOWER (profiles/r02_summary.md: 1.07 against 0.82 ms
// at 256^3, 512^3 alike) -- with the loads fully hidden the pass is bound by its barrier-separated compute phases (168
// registers allow 9-12 warps per SM), not by load latency; two de-phased CTAs per SM overlap those phases better than one
// CTA with a prefetched tile does.
static bool gamma_pipe_setting() {
  static const bool v = [] { const char* e = getenv("DMK_GAMMA_PIPE"); return e && e[0] == '1'; }();
  return v;
}

template <class PL, int KT, int NT, int MINB = 1>
static int launch_gamma(double2* data, const DmkFftPlan& pl, DmkLineMap map, int n0, int n_outer, const DmkGammaParams& gp,
                        cudaStream_t st, bool imag_xi, const DmkPeerStore& ps) {
  const size_t smem = (size_t)PL::N * KT * DMK_NC * sizeof(double2);
  const int tiles0 = (n0 + KT - 1) / KT;
  const long long grid = (long long)tiles0 * n_outer;
  if (grid > 0x7fffffffLL) { dmk_set_error("gamma pass grid too large"); return DMK_ERR_UNSUPPORTED; }
  // pipelined, persistent variant: tiles of 512 records (73.7 kB), two buffers, 288 threads = one butterfly of the widest
  // stage per thread; needs the row of a tile to be a multiple of 16 bytes (always: 144 B records)
  constexpr bool PIPE_OK = (PL::N * KT == 512);
  if (PIPE_OK && gamma_pipe_setting() && grid >= 2) {
    constexpr int NTP = 288;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const int g = (int)std::min<long long>(grid, sms);
#define DMK_GAMMA_PIPE_LAUNCH(IMAG_, REM_)                                                                                   \
  do {                                                                                                                       \
    DMK_CUDA_CHECK(cudaFuncSetAttribute(dmk_fast_gamma_pipe_kernel<PL, KT, IMAG_, NTP, REM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * smem))); \
    dmk_fast_gamma_pipe_kernel<PL, KT, IMAG_, NTP, REM_><<<g, NTP, 2 * smem, st>>>(data, pl.tw, map, n0, tiles0, (int)grid, gp, ps);  \
  } while (0)
    if (imag_xi) { if (ps.remote) DMK_GAMMA_PIPE_LAUNCH(true, true); else DMK_GAMMA_PIPE_LAUNCH(true, false); }
    else { if (ps.remote) DMK_GAMMA_PIPE_LAUNCH(false, true); else DMK_GAMMA_PIPE_LAUNCH(false, false); }
#undef DMK_GAMMA_PIPE_LAUNCH
    DMK_CUDA_CHECK(cudaGetLastError());
    return dmk_sum_div_partials(gp.div_partial, (int)grid, gp.div_accum, st);
  }
#define DMK_GAMMA_LAUNCH(IMAG_, REM_)                                                                                        \
  do {                                                                                                                       \
    if (smem > 48 * 1024) DMK_CUDA_CHECK(cudaFuncSetAttribute(dmk_fast_gamma_kernel<PL, KT, IMAG_, NT, MINB, REM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    dmk_fast_gamma_kernel<PL, KT, IMAG_, NT, MINB, REM_><<<(int)grid, NT, smem, st>>>(data, pl.tw, map, n0, tiles0, gp, ps);  \
  } while (0)
  if (imag_xi) { if (ps.remote) DMK_GAMMA_LAUNCH(true, true); else DMK_GAMMA_LAUNCH(true, false); }
  else { if (ps.remote) DMK_GAMMA_LAUNCH(false, true); else DMK_GAMMA_LAUNCH(false, false); }
#undef DMK_GAMMA_LAUNCH
  DMK_CUDA_CHECK(cudaGetLastError());
  return dmk_sum_div_partials(gp.div_partial, (int)grid, gp.div_accum, st);
}

int dmk_fast_gamma_zpass(double2* data, const DmkFftPlan& plz, DmkLineMap map, int n0, int n_outer, const DmkGammaParams& gp,
                         cudaStream_t st, bool imag_xi, bool* done, const DmkPeerStore* psp) {
  *done = true;
  if (dmk_fast_disabled()) { *done = false; return DMK_OK; }
  DmkPeerStore ps;
  if (psp) ps = *psp; else memset(&ps, 0, sizeof ps);
  // the Gamma section needs ~168 registers, so CTAs are small (<= 192 threads, 2 per SM): tools/fft_tune.cu sweeps
  switch (plz.n) {
    case 16: return launch_gamma<DmkFastPlan<4, 4, 1>, 4, 160, 2>(data, plz, map, n0, n_outer, gp, st, imag_xi, ps);
    case 32: return launch_gamma<DmkFastPlan<8, 4, 1>, 4, 160, 2>(data, plz, map, n0, n_outer, gp, st, imag_xi, ps);
    case 64: return launch_gamma<DmkFastPlan<8, 8, 1>, 4, 160, 2>(data, plz, map, n0, n_outer, gp, st, imag_xi, ps);
    case 128: return launch_gamma<DmkFastPlan<16, 8, 1>, 4, 192, 2>(data, plz, map, n0, n_outer, gp, st, imag_xi, ps);
    case 256: {
      static const int v = [] { const char* e = getenv("DMK_GZ_VARIANT"); return e ? atoi(e) : 0; }();        // comparison runs
      if (v == 1) return launch_gamma<DmkFastPlan<16, 16, 1>, 2, 192, 3>(data, plz, map, n0, n_outer, gp, st, imag_xi, ps);
      if (v == 2) return launch_gamma<DmkFastPlan<16, 16, 1>, 1, 96, 5>(data, plz, map, n0, n_outer, gp, st, imag_xi, ps);
      if (v == 3) return launch_gamma<DmkFastPlan<16, 16, 1>, 1, 96, 6>(data, plz, map, n0, n_outer, gp, st, imag_xi, ps);
      if (v == 4) return launch_gamma<DmkFastPlan<16, 16, 1>, 1, 160, 3>(data, plz, map, n0, n_outer, gp, st, imag_xi, ps);
      return launch_gamma<DmkFastPlan<16, 16, 1>, 2, 192, 2>(data, plz, map, n0, n_outer, gp, st, imag_xi, ps);
    }
    case 512: return launch_gamma<DmkFastPlan<8, 8, 8>, 1, 192, 2>(data, plz, map, n0, n_outer, gp, st, imag_xi, ps);
    default: *done = false; return DMK_OK;
  }
}

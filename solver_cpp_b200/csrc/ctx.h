// Internal definition of the solver context shared by the translation units of libdamask_grid_b200.so.
#pragma once
#include <vector>
#include <string>
#include "../../include/dmk_grid.h"
#include "fft_host.h"

struct DmkPointArrays;   // constitutive.cuh
struct DmkMech;          // mech_solver.cu

struct dmk_ctx {
  dmk_grid_desc g;
  int nx, ny, nz, nxr;
  int P, rank;                    // world size / rank
  int nz_loc, z_off, ny_loc, y_off;
  long long n_local;              // cells in the local z-slab
  long long n_global;
  double wgt;                     // 1 / n_global
  cudaStream_t stream; bool own_stream;
  dmk_comm comm; bool has_comm;
  void* nccl;                     // native NCCL binding (nccl_comm.cpp), or null: the collectives go through the callback table
  long long launches;
  // optional section timing of one residual evaluation (dmk_set_profiling): events on the context's stream
  bool profiling; cudaEvent_t ev[8]; double prof_ms[8]; long long prof_n;

  // ---- spectral
  DmkFftPlan plx, ply, plz;
  double2* xi_own[3]; double2* xi_oth[3]; unsigned char* nyq[3];
  bool xi_imag;
  double scaled_size[3], rescale[3];
  int kt_y, kt_z;
  double2* bufA;                  // [nz_loc][ny][nxr][9] after x/y passes; [nz][ny_loc][nxr][9] in the z pass
  double2* bufB;                  // exchange buffer [peer][nz_loc][ny_loc][nxr][9] (world_size > 1 only)
  size_t hat_records;             // records in bufA
  // fused exchange over peer memory (dmk_p2p_import): peers' bufA / bufB, opened through CUDA IPC.  When on, the forward y
  // pass stores into the peers' bufB (the z-pass layout lives in bufB), the inverse z pass stores into the peers' bufA.
  bool p2p; double2* peerA[DMK_MAX_PEERS]; double2* peerB[DMK_MAX_PEERS];
  double* d_Cref;                 // 81, row-major C_ref(l,n,m,o)
  double* d_dc;                   // 9, record order
  double* d_div;                  // 1
  double* d_div_partial;          // per-CTA divergence sums of the fused z pass (nxr * ny_loc doubles: one per tile at most)
  double* d_scratch;              // small device scratch (>= 4096 doubles)
  double* h_pinned;               // small pinned host scratch (>= 4096 doubles)
  bool gamma_set;

  // ---- constitutive
  DmkPointArrays* pts;
  // input staging of dmk_mech_form_residual_host: F arrives in `stage_chunks` z-chunks on `copy_stream`, chunk k is complete
  // when stage_ev[k] fires; the lane-persistent kernel is launched per chunk behind its event (0: F is resident)
  cudaStream_t copy_stream; cudaEvent_t stage_ev[8]; cudaEvent_t stage_begin; int stage_chunks; bool stage_ready;
  const double* stage_dst;        // device field the staged copy fills (the chunked launches apply to this input only)
  // forward half of the distributed FFT pipelined under the per-point kernel (spectral.cu: dmk_pipe_*): the per-point kernel is
  // launched per z-chunk, and chunk k's x r2c + y pass (with its stores into the peers' z-pass buffers) run on `pipe_stream`
  // while chunk k+1 is being integrated
  bool pipe_on; int pipe_chunks, pipe_done; const double* pipe_src; bool pipe_ready;
  cudaStream_t pipe_stream; cudaEvent_t pipe_ev[8]; cudaEvent_t pipe_end;
  cudaStream_t chunk_stream[2]; cudaEvent_t pipe_begin;   // the per-point kernel's z-chunks alternate between two streams
  dmk_numerics num;
  int opt_mass;                   // who takes a long list of given-up points: -1 = DMK_PHENO_MASS / default; 0 nested, 1 machine, 2 lanes
  // ---- mechanical solver
  DmkMech* mech;
  // ---- Fortran-named drop-in mirrors (pinned host)
  double *m_F0, *m_F, *m_P, *m_dPdF; int* m_ill;
};

int dmk_spectral_setup(dmk_ctx* c);
void dmk_spectral_teardown(dmk_ctx* c);
void dmk_points_teardown(dmk_ctx* c);
void dmk_mech_teardown(dmk_ctx* c);
void dmk_nccl_teardown(dmk_ctx* c);
// full transforms between a real slab field (device) and the z-pass layout in c->bufA
int dmk_fft_forward_internal(dmk_ctx* c, const double* d_real);
int dmk_fft_backward_internal(dmk_ctx* c, double* d_real, double scale);
// pipelined forward half (multi-GPU, fused exchange): begin before the per-point kernel whose OUTPUT d_real is; dmk_pipe_chunk
// after the per-point launch of every z-chunk; dmk_pipe_end waits for it (*done: the forward transform of d_real is in place)
int dmk_pipe_begin(dmk_ctx* c, const double* d_real);
int dmk_pipe_chunk(dmk_ctx* c, int k, int chunks, cudaStream_t chunk_stream);
int dmk_pipe_end(dmk_ctx* c, bool* done);
int dmk_spectrum_rms(dmk_ctx* c, const double* d_real, int mode /*0 divergence, 1 curl*/, double* d_accum);   // raw sum, += into *d_accum
int dmk_gamma_internal(dmk_ctx* c);     // fused z pass on c->bufA using c->d_dc, accumulating into c->d_div
double dmk_finish_div(dmk_ctx* c, double raw_sum);
const unsigned* dmk_points_given_up_counter(dmk_ctx* c);
bool dmk_points_single_constituent(const dmk_ctx* c);        // false: isostrain homogenisation with N_constituents > 1 (constitutive.cu)
const double* dmk_points_given_up_total(dmk_ctx* c);           // ... summed over all ranks, as a double (valid when the mass-deferral decision ran)
int dmk_fft_forward_finish_pipelined(dmk_ctx* c);
int dmk_fft_forward_redo_pipelined(dmk_ctx* c, const double* d_real);
int dmk_mech_before_point_forward(dmk_ctx* c);   // computes the pending tangent reductions before state0 <- state

// Fortran-order <-> row-major helpers for small host tensors
static inline void f2c33(const double* f, double* c) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c[3 * i + j] = f[i + 3 * j]; }
static inline void c2f33(const double* c, double* f) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) f[i + 3 * j] = c[3 * i + j]; }
static inline void f2c3333(const double* f, double* c) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++)
    c[((i * 3 + j) * 3 + k) * 3 + l] = f[i + 3 * j + 9 * k + 27 * l];
}
static inline void c2f3333(const double* c, double* f) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++)
    f[i + 3 * j + 9 * k + 27 * l] = c[((i * 3 + j) * 3 + k) * 3 + l];
}

// Kernel and launcher of the register-resident slip-only phenopowerlaw formulation (pheno_fast_dev.cuh); replaces
// f_homogenization_mechanical_response (/root/reference/src/homogenization.f90:219-266) for such phases.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <string>
#include "pheno_fast_dev.cuh"
#include "pheno_lanes_dev.cuh"

// One thread per point, one warp-sized pair of warps per block (DMK_PF_THREADS points), phase tables as a
// __grid_constant__ parameter, state-integrator arrays in shared memory ([k][thread]).  Blocks are small and end
// without a barrier: the per-point work varies by a factor of several (Newton iteration counts), and with 128-thread
// blocks 18 % of all warp samples were warps parked at the final __syncthreads waiting for the slowest one
// (profiles/r01_summary.md).  Points whose Newton path is ill-conditioned are appended to pa.defer_list (pf_stress).
#define DMK_PF_THREADS 64
#ifndef DMK_PF_MINBLOCKS
#define DMK_PF_MINBLOCKS 4
#endif
template <int NSL, int NEXP, int UNR>
__global__ void __launch_bounds__(DMK_PF_THREADS, DMK_PF_MINBLOCKS)
dmk_pheno_fast_kernel(const __grid_constant__ DmkPhenoK<NSL> K, DmkPointArrays pa, int phase, DmkNumDev num, int defer_nLp,
                      const double* __restrict__ Fin, double* __restrict__ Pout, double Dt) {
  extern __shared__ double pf_sm[];
  const DmkScr scr = {pf_sm + threadIdx.x, DMK_PF_THREADS, pf_sm + PfOff<NSL>::TOTAL * DMK_PF_THREADS + threadIdx.x, DMK_PF_THREADS, &K};   // all of it in shared memory
  DmkCnt cnt = {0, 0, 0, 0, 0, 0};
  unsigned ill = 0, ndefer = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pa.n; i += (long long)gridDim.x * blockDim.x) {
    if (pa.phase[i] != phase) continue;
    const int rc = pf_point<NSL, NEXP, UNR>(K, num, defer_nLp, pa, i, Fin, Pout, Dt, scr, cnt);
    if (rc == PF_DEFER) { pa.defer_list[atomicAdd(pa.d_defer_count, 1u)] = (int)i; ndefer++; }
    else if (rc != PF_OK) ill = 1;
  }
  // counters: warp reduce, one atomic per warp and non-zero counter
  unsigned v[8] = {cnt.Lp_iter, cnt.Li_iter, cnt.stress_calls, cnt.state_iter, cnt.substeps, cnt.cutbacks, ill, ndefer};
#pragma unroll
  for (int k = 0; k < 8; k++) {
    unsigned s = v[k];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&pa.d_counters[k], (unsigned long long)s);
  }
  if (ill) atomicOr(pa.d_ill, 1);
}

// DMK_PHENO_DEFER=<n>: hand a point to the reference-order kernel when one stress integration needs more than n Newton
// iterations or a line-search step (0: never; default 8)
static int defer_setting() {
  static const int v = [] { const char* e = getenv("DMK_PHENO_DEFER"); return e ? atoi(e) : 8; }();
  return v < 0 ? 0 : v;
}

// DMK_PHENO_UNROLL=<1|4|12>: unroll factor of the slip-system loops inside the Newton iteration.  Nested schedule: default 1
// (rolled) -- the kernels are instruction-fetch sensitive, unrolled 12-fold `no_instruction` was the top stall at a 57 %
// instruction-cache hit rate (128^3 solver iterate 31.3 ms against 28.1 ms rolled), profiles/r01_summary.md.  Lane schedule:
// default DMK_PL_UNR (per-loop factors chosen for latency hiding, profiles/r02_summary.md); 1 = the rolled instantiation
static int unroll_setting() {
  static const int v = [] { const char* e = getenv("DMK_PHENO_UNROLL"); return e ? atoi(e) : 0; }();
  return v;
}

template <int NSL, int NEXP, int UNR>
static int launch_pheno_fast(dmk_ctx* c, DmkPointArrays& pa, int ph, const DmkDevPhase& d, const DmkNumDev& nd, const double* d_F,
                             double* d_P, double dt) {
  DmkPhenoK<NSL> K;
  dmk_pheno_fast_tables<NSL>(d, K);
  const size_t smem = (size_t)(PfOff<NSL>::TOTAL + PfOff<NSL>::COLD_TOTAL) * DMK_PF_THREADS * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    DMK_CUDA_CHECK(cudaFuncSetAttribute(dmk_pheno_fast_kernel<NSL, NEXP, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  { int rc = dmk_wait_staged_input(c, d_F); if (rc) return rc; }
  const long long blocks = (pa.n + DMK_PF_THREADS - 1) / DMK_PF_THREADS;
  const int grid = (int)std::max(1LL, std::min(blocks, (long long)INT32_MAX));
  DMK_CUDA_CHECK(cudaMemsetAsync(pa.d_defer_count, 0, sizeof(unsigned), c->stream));
  dmk_pheno_fast_kernel<NSL, NEXP, UNR><<<grid, DMK_PF_THREADS, smem, c->stream>>>(K, pa, ph, nd, defer_setting(), d_F, d_P, dt);
  return DMK_OK;
}


// ------------------------------------------------------------------------------------------------ lane-persistent kernel
// pheno_lanes_dev.cuh.  One resident wave of 64-thread blocks; a warp claims chunks of DMK_PL_CHUNK consecutive cells from a
// global queue (one atomic per chunk) and deals them to its lanes as they become free, so neighbouring lanes keep working
// on neighbouring cells (coalesced structure-of-arrays accesses) and no warp is left with a long tail.  No barrier anywhere.
// Block size: 2 blocks of 128 threads per SM (8 warps, bounded by 255 registers and 840 B of shared scratch per thread alike).
// With 128-thread blocks the 215 kB of scratch leave room for ONE COPY PER BLOCK OF THE PHASE TABLES in shared memory
// (DMK_PL_KSMEM): the rolled slip-system loops read their Schmid tensors and coefficients with indexed constant-bank loads
// (LDC), the indexed-constant cache holds 79 % of them (ncu idc__request_hit_rate) and the misses compete with the
// instruction-cache misses of this 60 kB kernel for the GPC-level cache, which ran at 88 % of its request rate (gcc__cache_
// requests_type_{instruction 57 %, constant 30 %}, profiles/r02_summary.md).  From shared memory every table read is one
// broadcast LDS.
#ifndef DMK_PL_THREADS
#define DMK_PL_THREADS 128
#endif
#ifndef DMK_PL_MINBLOCKS
#define DMK_PL_MINBLOCKS (256 / DMK_PL_THREADS)
#endif
#ifndef DMK_PL_KSMEM
#define DMK_PL_KSMEM 1
#endif
// DMK_PL_BLOCKVOTE: the vote that picks the next block (residual / Newton step / state block) is taken over the whole thread
// block instead of per warp, so that all warps of a block run the same code at the same time: the kernel's hot code (28 kB of
// a 48 kB executed footprint) does not fit the SM's 32 kB instruction cache when 8 warps sit in 8 different places of it
#ifndef DMK_PL_BLOCKVOTE
#define DMK_PL_BLOCKVOTE 0
#endif
#define DMK_PL_CHUNK 256
// unroll factors RRJJCC of the rate / Jacobian / correction loops of the Newton iteration (pheno_fast_dev.cuh: unr_rate ...);
// DMK_PHENO_UNROLL=1 selects the rolled instantiation for comparison
#ifndef DMK_PL_UNR
#define DMK_PL_UNR 1
#endif
template <int NSL, int NEXP, int UNR, bool VOTE3, bool CUTBACK, int MINB>
__global__ void __launch_bounds__(DMK_PL_THREADS, MINB)
dmk_pheno_lanes_kernel(const __grid_constant__ DmkPhenoK<NSL> K, DmkPointArrays pa, int phase, DmkNumDev num,
                       const double* __restrict__ Fin, double* __restrict__ Pout, double Dt, long long cell_begin,
                       long long cell_end, int3 vw, unsigned long long* __restrict__ work, const int* __restrict__ list, const unsigned* __restrict__ list_n,
                       const double* __restrict__ list_total, double list_min) {
  // cells [cell_begin, cell_end) are this launch's share (whole slab, or one z-chunk); *work, this launch's queue, counts from 0.
  // list != nullptr: the share is list[0 .. *list_n) instead (the cells an earlier launch gave up), and the launch does
  // nothing unless *list_total -- the length of that list summed over all ranks -- exceeds list_min (the decision is made on
  // the device: no host round trip; on the GLOBAL count, so that it does not depend on the slab decomposition)
  if (list) {
    cell_begin = 0; cell_end = (long long)*list_n;
    if (!(*list_total > list_min)) return;
  }
  extern __shared__ double pf_sm[];
  const unsigned FULL = 0xffffffffu, lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
#if DMK_PL_KSMEM
  // the block's copy of the phase tables (the only barrier of the kernel); the slip-system loops index this copy (pf_tab),
  // everything else (elastic constants, compliance) stays an immediate constant-bank operand
  unsigned* ksm = reinterpret_cast<unsigned*>(pf_sm + (PfOff<NSL>::TOTAL + PfOff<NSL>::COLD_TOTAL) * DMK_PL_THREADS);
  for (unsigned k = threadIdx.x; k < sizeof(DmkPhenoK<NSL>) / 4; k += DMK_PL_THREADS) ksm[k] = reinterpret_cast<const unsigned*>(&K)[k];
  int* vote_sm = reinterpret_cast<int*>(ksm + sizeof(DmkPhenoK<NSL>) / 4);        // [3][4] (block vote)
  if (threadIdx.x < 12) vote_sm[threadIdx.x] = 0;
  __syncthreads();
  const void* tab = ksm;
#else
  const void* tab = &K;
  static_assert(!DMK_PL_BLOCKVOTE, "the block vote keeps its counters behind the shared-memory tables");
#endif
  unsigned trip = 0;
  const DmkScr scr = {pf_sm + threadIdx.x, DMK_PL_THREADS, pf_sm + PfOff<NSL>::TOTAL * DMK_PL_THREADS + threadIdx.x, DMK_PL_THREADS, tab};
  PlLane<NSL> L;
  L.mode = PL_NEED;
  long long c_next = 0, c_end = 0;                      // the warp's current chunk (warp-uniform)
  bool drained = false;
  unsigned n_Lp = 0, n_calls = 0, n_sub = 0, n_cut = 0, n_ill = 0, n_defer = 0;
  unsigned t_lp = 0, l_lp = 0, t_sb = 0, l_sb = 0, t_rf = 0, l_rf = 0;     // schedule statistics (lane 0: warp totals)
  double S[6];
  auto fail = [&]() { L.ok = false; L.mode = PL_RESTART; };               // the running sub-step is cut back
  auto done = [&](bool ill) {
    n_Lp += L.nLp_pt; n_calls += L.calls_pt; n_sub += L.nsub; n_cut += L.ncut; n_ill += ill ? 1u : 0u;
    L.mode = PL_NEED;
  };
  for (;;) {
    // ---- vote: a trip runs ONE block -- the Newton residual, the Newton step, or refill + sub-step set-up + state block --
    // whichever has most lanes waiting for it (weights vw).  Running everything every trip left 12.8 of 32 lanes active per
    // instruction (ncu, profiles/r01_summary.md): the state block and the refill were executed for the few lanes that needed
    // them at that moment.  Waiting lanes accumulate until their block wins, go through it together and then iterate in step.
    const unsigned bL = __ballot_sync(FULL, L.mode == PL_LP);
    const unsigned bN = VOTE3 ? __ballot_sync(FULL, L.mode == PL_NEWTON) : 0u;
    const unsigned bS = __ballot_sync(FULL, L.mode == PL_INIT || L.mode == PL_CONV || L.mode == PL_RESTART);
    unsigned need = drained ? 0u : __ballot_sync(FULL, L.mode == PL_NEED);
    int pL = __popc(bL), pN = __popc(bN), pS = __popc(bS) + __popc(need);
#if DMK_PL_BLOCKVOTE
    {                                                  // totals over the block: triple-buffered counters, one barrier per trip
      int* cnt = vote_sm + 4 * (trip % 3u);
      if (lane == 0) { if (pL) atomicAdd(cnt, pL); if (pN) atomicAdd(cnt + 1, pN); if (pS) atomicAdd(cnt + 2, pS); }
      __syncthreads();
      pL = cnt[0]; pN = cnt[1]; pS = cnt[2];
      if (threadIdx.x < 3) vote_sm[4 * ((trip + 2u) % 3u) + threadIdx.x] = 0;      // next used after the barrier of trip + 1
      trip++;
    }
#endif
    if (!(pL | pN | pS)) break;
    if constexpr (VOTE3) {
      // the Newton step runs only for the lanes whose residual test failed (2/3 of the Newton lanes), so it is a block of
      // its own in which such lanes accumulate
      const int cL = vw.x * pL, cN = vw.y * pN, cS = vw.z * pS;
      if (cN >= cL && cN >= cS) {
        t_lp++; l_lp += __popc(bN);
        if (L.mode == PL_NEWTON) {
          if (pl_newton<NSL, UNR>(K, scr, L) == PL_FAIL) { L.nLp_pt += (unsigned)L.nLp; fail(); }
          else L.mode = PL_LP;
        }
        continue;
      }
      if (cL >= cS) {
        t_lp++; l_lp += __popc(bL);
        if (L.mode == PL_LP) {
          const int rc = pl_residual<NSL, NEXP, UNR>(K, num, scr, L);
          if (rc == PL_DONE) L.mode = PL_CONV;
          else if (rc == PL_STEP) L.mode = PL_NEWTON;
          else if (rc == PL_FAIL) { L.nLp_pt += (unsigned)L.nLp; fail(); }
        }
        continue;
      }
    } else if (pL >= pS) {
      t_lp++; l_lp += __popc(bL);
      if (L.mode == PL_LP) {
        const int rc = pl_lp_iteration<NSL, NEXP, UNR>(K, num, scr, L);
        if (rc == PL_DONE) L.mode = PL_CONV;
        else if (rc == PL_FAIL) { L.nLp_pt += (unsigned)L.nLp; fail(); }
      }
      continue;
    }
    // ---- refill: lanes without a point take the next cell of the warp's chunk
    if (need) { t_rf++; l_rf += __popc(need); }
    while (need) {
      if (c_next >= c_end) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(work, (unsigned long long)DMK_PL_CHUNK);
        base = __shfl_sync(FULL, base, 0) + (unsigned long long)cell_begin;
        if ((long long)base >= cell_end) { drained = true; break; }
        c_next = (long long)base;
        c_end = min(c_next + DMK_PL_CHUNK, cell_end);
      }
      const int avail = (int)min(32LL, c_end - c_next);
      const int rank = __popc(need & lt);
      if (L.mode == PL_NEED && rank < avail) {
        const long long i = list ? (long long)list[c_next + rank] : c_next + rank;
        if (pa.phase[i] == phase) { pl_fetch<NSL>(pa, num, i, L); L.mode = PL_RESTART; }
      }
      c_next += min(__popc(need), avail);
      need = __ballot_sync(FULL, L.mode == PL_NEED);
    }
    // ---- sub-step set-up (first, next or cut-back sub-step of the lane's point)
    if (L.mode == PL_RESTART) {
      const int rc = pl_substep_begin<NSL, CUTBACK>(num, pa, Fin, Pout, Dt, scr, L, S);
      if (rc == PL_CONTINUE) L.mode = PL_INIT;
      else if (rc == PL_ILL) done(true);
      else if (rc == PL_GIVEUP) { pa.defer_list[atomicAdd(pa.d_defer_count, 1u)] = (int)L.i; n_defer++; L.mode = PL_NEED; }
      // PL_FAIL: stays in PL_RESTART with ok = false: cut back on the next visit
    }
    // ---- state block
    const unsigned group = __ballot_sync(FULL, L.mode == PL_INIT || L.mode == PL_CONV);
    t_sb++; l_sb += __popc(group);
    if (L.mode == PL_INIT || L.mode == PL_CONV) {
      const int rc = pl_state_block<NSL, NEXP, CUTBACK>(K, num, pa, Fin, Pout, scr, L, S, L.mode == PL_INIT, group);
      if (rc == PL_CONTINUE) L.mode = PL_LP;
      else if (rc == PL_RESTART) L.mode = PL_RESTART;
      else if (rc == PL_DONE) done(false);
      else fail();
    }
  }
  // counters (d_counters: Lp, Li, stress calls, state iterations, sub-steps, cut-backs, ill, deferred)
  unsigned v[8] = {n_Lp, n_calls, n_calls, n_calls, n_sub, n_cut, n_ill, n_defer};
#pragma unroll
  for (int k = 0; k < 8; k++) {
    unsigned s = v[k];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(FULL, s, off);
    if (lane == 0 && s) atomicAdd(&pa.d_counters[k], (unsigned long long)s);
  }
  if (n_ill) atomicOr(pa.d_ill, 1);
  if (lane == 0) {
    const unsigned w[6] = {t_lp, l_lp, t_sb, l_sb, t_rf, l_rf};
#pragma unroll
    for (int k = 0; k < 6; k++) atomicAdd(&pa.d_counters[8 + k], (unsigned long long)w[k]);
  }
}

// DMK_PHENO_VOTE=2: two-way vote (whole Newton iteration / state block) instead of the default three-way vote (residual /
// Newton step / state block) in the lane-persistent kernel (128^3 solver iterate: 23.0 ms against 21.6 ms)
static bool vote3_setting() {
  static const bool v = [] { const char* e = getenv("DMK_PHENO_VOTE"); return !(e && atoi(e) == 2); }();
  return v;
}

// DMK_PHENO_VOTE_WEIGHTS=l,n,s: integer weights of the lane counts in the three-way vote.  Default 4,4,3: the state block is
// the longest of the three, letting it wait for a few more lanes pays (sweep in profiles/r01_summary.md)
static int3 vote_weights() {
  static const int3 v = [] { int3 w = make_int3(4, 4, 3); const char* e = getenv("DMK_PHENO_VOTE_WEIGHTS"); if (e) sscanf(e, "%d,%d,%d", &w.x, &w.y, &w.z); return w; }();
  return v;
}

// What happens to a point whose full first sub-step fails (the reference cuts back, phase_mechanical.f90:1040-1085):
//   DMK_PHENO_CUTBACK=lanes    cut back in-lane, in the same launch (pheno_lanes_dev.cuh, CUTBACK = true)
//   DMK_PHENO_CUTBACK=generic  always handed to the reference-order kernel (one thread per point, 9x9 LU)
//   default (auto)             handed to the reference-order kernel while such points are few (a running solve: 0 - 2 % of the
//                              cells); when more than 1/32 of the slab gave up -- the first evaluations after a virgin state,
//                              where EVERY cell cuts back and the one-thread-per-point kernel needs 3.4 s at 256^3 -- the list
//                              goes through a lane-scheduled kernel instead (DMK_PHENO_MASS, constitutive.cu).  Both launches
//                              are issued; each reads the count on the device and returns at once if the list is not its share.
static int cutback_setting() {
  static const int v = [] { const char* e = getenv("DMK_PHENO_CUTBACK"); return !e ? 0 : std::string(e) == "lanes" ? 1 : std::string(e) == "generic" ? 2 : 0; }();
  return v;
}

// DMK_PHENO_SCHEDULE=nested: one thread walks one point through the nested loops (dmk_pheno_fast_kernel); default: lanes
static bool lanes_setting() {
  static const bool v = [] { const char* e = getenv("DMK_PHENO_SCHEDULE"); return !(e && std::string(e) == "nested"); }();
  return v;
}

// CTAs per SM: 4 x 64 threads = 8 warps, bounded by 255 registers and 840 B of shared scratch per thread alike.  Round 2 tried
// more (profiles/r02_summary.md): with L1 / D of the scratch moved to an L2-resident global array, 456 B per thread, the kernel
// compiled for 5 (204 registers) and 6 CTAs (168 registers, 336 B stack, 464 B spill stores) ran the 256^3 evaluation in 179.7
// and 193.7 ms against 168.2 ms for 4 CTAs on the same layout and 160.4 ms with everything in shared memory: the spills and the
// L2 round trips cost more than 10 or 12 warps per SM hide.
template <int NSL, int NEXP, int UNR, bool VOTE3, bool CUTBACK, int MINB = DMK_PL_MINBLOCKS>
static int launch_pheno_lanes(dmk_ctx* c, DmkPointArrays& pa, int ph, const DmkDevPhase& d, const DmkNumDev& nd, const double* d_F,
                              double* d_P, double dt, bool from_list = false, double list_min = 0.0) {
  DmkPhenoK<NSL> K;
  dmk_pheno_fast_tables<NSL>(d, K);
  static_assert(sizeof(DmkPhenoK<NSL>) % 8 == 0, "phase tables: whole doubles");
  const size_t smem = (size_t)(PfOff<NSL>::TOTAL + PfOff<NSL>::COLD_TOTAL) * DMK_PL_THREADS * sizeof(double) + (DMK_PL_KSMEM ? sizeof(DmkPhenoK<NSL>) + 64 : 0);
  static int grid = 0;
  if (!grid) {
    DMK_CUDA_CHECK(cudaFuncSetAttribute(dmk_pheno_lanes_kernel<NSL, NEXP, UNR, VOTE3, CUTBACK, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dmk_pheno_lanes_kernel<NSL, NEXP, UNR, VOTE3, CUTBACK, MINB>, DMK_PL_THREADS, smem) != cudaSuccess || occ < 1) occ = 1;
    grid = sms * occ;                                  // one resident wave
  }

  if (from_list) {                                     // the cells the launch before this one gave up (pa.defer_list), if they are many
    DMK_CUDA_CHECK(cudaMemsetAsync(pa.d_work, 0, sizeof(unsigned long long), c->stream));
    dmk_pheno_lanes_kernel<NSL, NEXP, UNR, VOTE3, CUTBACK, MINB><<<grid, DMK_PL_THREADS, smem, c->stream>>>(K, pa, ph, nd, d_F, d_P, dt, 0, 0, vote_weights(), pa.d_work, pa.defer_list, pa.d_defer_count, pa.d_defer_total, list_min);
    c->launches++;
    return DMK_OK;
  }
  DMK_CUDA_CHECK(cudaMemsetAsync(pa.d_defer_count, 0, sizeof(unsigned), c->stream));
  // one launch over the slab, or one per z-chunk: staged input (chunk k of F arrives behind stage_ev[k]) and / or the pipelined
  // forward FFT (spectral.cu: dmk_pipe_*, which transforms chunk k's stress while chunk k+1 is integrated).  Chunks for the
  // pipeline go to two alternating streams, each with its own work queue: chunk k+1's CTAs move in as chunk k's retire, so
  // the tail of a chunk (lanes finishing their last cells, ~1 ms) is not paid eight times as with launches in one stream
  const bool staged = c->stage_chunks > 1 && d_F == c->stage_dst && pa.n_const == 1;
  const bool piped = c->pipe_on && d_P == c->pipe_src && pa.n_cells_of_phase[ph] == pa.n && pa.n_const == 1;     // the only phase: its chunk is complete after its launch
  const int chunks = staged ? c->stage_chunks : piped ? c->pipe_chunks : 1;
  if (staged && piped && c->stage_chunks != c->pipe_chunks) c->pipe_on = false;
  const bool conc = piped && c->pipe_on && chunks > 1;
  const long long plane = (long long)c->nx * c->ny;
  DMK_CUDA_CHECK(cudaMemsetAsync(pa.d_work, 0, 8 * sizeof(unsigned long long), c->stream));
  if (conc) {
    DMK_CUDA_CHECK(cudaEventRecord(c->pipe_begin, c->stream));
    for (int q = 0; q < 2; q++) DMK_CUDA_CHECK(cudaStreamWaitEvent(c->chunk_stream[q], c->pipe_begin, 0));
  }
  for (int k = 0; k < chunks; k++) {
    const long long z0 = (long long)c->nz_loc * k / chunks, z1 = (long long)c->nz_loc * (k + 1) / chunks;
    const long long first = z0 * plane, end = (chunks == 1) ? pa.n : z1 * plane;     // one launch: every material point (isostrain: n_const per cell)
    cudaStream_t st = conc ? c->chunk_stream[k & 1] : c->stream;
    if (end > first) {
      if (staged) DMK_CUDA_CHECK(cudaStreamWaitEvent(st, c->stage_ev[k], 0));
      const long long warps_needed = (end - first + 31) / 32;
      const int g = (int)std::max(1LL, std::min((long long)grid, (warps_needed + DMK_PL_THREADS / 32 - 1) / (DMK_PL_THREADS / 32)));
      dmk_pheno_lanes_kernel<NSL, NEXP, UNR, VOTE3, CUTBACK, MINB><<<g, DMK_PL_THREADS, smem, st>>>(K, pa, ph, nd, d_F, d_P, dt, first, end, vote_weights(), pa.d_work + k, nullptr, nullptr, nullptr, 0.0);
      if (k > 0) c->launches++;
    }
    if (piped) { int rc = dmk_pipe_chunk(c, k, chunks, st); if (rc) return rc; }
  }
  if (conc) for (int k = 0; k < chunks; k++) DMK_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->pipe_ev[k], 0));
  return DMK_OK;
}

// Runs phase `ph` through the register-resident kernel if it qualifies (*handled), else leaves it to the generic kernel.
__global__ void dmk_defer_total_kernel(const unsigned* __restrict__ count, double* __restrict__ total) { *total = (double)*count; }

int dmk_points_pheno_fast(dmk_ctx* c, int ph, const DmkNumDev& nd, const double* d_F, double* d_P, double dt, bool* handled,
                          bool* deferred, double* defer_max) {
  DmkPointArrays& pa = *c->pts;
  int nexp = 0;
  *handled = false; *deferred = false; *defer_max = -1.0;
  if (!dmk_pheno_fast_eligible(pa.h_dev_phases[ph], 12, &nexp)) return DMK_OK;       // 12 slip systems: fcc, bcc {110}
  *handled = true; *deferred = defer_setting() > 0;
  const DmkDevPhase& d = pa.h_dev_phases[ph];
  const int u = unroll_setting();
  int rc;
  const bool cb = cutback_setting() == 1;
  const bool adaptive = cutback_setting() == 0 && dmk_pheno_lanes_eligible(nd, true);
  if (lanes_setting() && dmk_pheno_lanes_eligible(nd, cb)) {
    *deferred = !cb;                                    // in-lane cut-backs: nothing is handed to the reference-order kernel
#define DMK_PL_LAUNCH(NEXP_)                                                                                   \
  (cb ? launch_pheno_lanes<12, NEXP_, DMK_PL_UNR, true, true>(c, pa, ph, d, nd, d_F, d_P, dt)                          \
      : !vote3_setting() ? launch_pheno_lanes<12, NEXP_, DMK_PL_UNR, false, false>(c, pa, ph, d, nd, d_F, d_P, dt)     \
      : (u == 1)  ? launch_pheno_lanes<12, NEXP_, 1, true, false>(c, pa, ph, d, nd, d_F, d_P, dt)                      \
                  : launch_pheno_lanes<12, NEXP_, DMK_PL_UNR, true, false>(c, pa, ph, d, nd, d_F, d_P, dt))
    rc = (nexp == 20) ? (DMK_PL_LAUNCH(20)) : (DMK_PL_LAUNCH(0));
#undef DMK_PL_LAUNCH
    if (rc) return rc;
    // the length of the given-up list over ALL ranks (one all-reduce of one double): decides who takes the list -- the same
    // cells take the same kernel whatever the slab decomposition, so a z-slab job reproduces the single-GPU job bit for bit --
    // and tells the pipelined forward FFT whether cells changed after their chunk's pass (mech_solver.cu: evaluate)
    if (cb) DMK_CUDA_CHECK(cudaMemsetAsync(pa.d_defer_total, 0, sizeof(double), c->stream));
    else {
      dmk_defer_total_kernel<<<1, 1, 0, c->stream>>>(pa.d_defer_count, pa.d_defer_total);
      c->launches++;
      if (c->P > 1) {
        if (!c->has_comm || !c->comm.allreduce_sum) { dmk_set_error("no allreduce callback (dmk_set_comm)"); return DMK_ERR_STATE; }
        if (c->comm.allreduce_sum(c->comm.user, pa.d_defer_total, 1)) { dmk_set_error("allreduce failed"); return DMK_ERR_STATE; }
      }
    }
    if (!cb && adaptive && dmk_mass_deferral_setting(c) != 0) {       // mass deferral
      const double many = (double)(c->n_global / 32);
      if (dmk_mass_deferral_setting(c) == 2) {            // ... through the in-lane cut-back kernel (else: constitutive.cu)
        rc = (nexp == 20) ? launch_pheno_lanes<12, 20, DMK_PL_UNR, true, true>(c, pa, ph, d, nd, d_F, d_P, dt, true, many)
                          : launch_pheno_lanes<12, 0, DMK_PL_UNR, true, true>(c, pa, ph, d, nd, d_F, d_P, dt, true, many);
        if (rc) return rc;
      }
      *defer_max = many;
    }
    DMK_CUDA_CHECK(cudaGetLastError());
    return DMK_OK;
  }
  if (nexp == 20) rc = (u == 12) ? launch_pheno_fast<12, 20, 12>(c, pa, ph, d, nd, d_F, d_P, dt)
                 : (u == 4) ? launch_pheno_fast<12, 20, 4>(c, pa, ph, d, nd, d_F, d_P, dt)
                            : launch_pheno_fast<12, 20, 1>(c, pa, ph, d, nd, d_F, d_P, dt);
  else rc = (u == 12) ? launch_pheno_fast<12, 0, 12>(c, pa, ph, d, nd, d_F, d_P, dt)
          : (u == 4) ? launch_pheno_fast<12, 0, 4>(c, pa, ph, d, nd, d_F, d_P, dt)
                     : launch_pheno_fast<12, 0, 1>(c, pa, ph, d, nd, d_F, d_P, dt);
  if (rc) return rc;
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}

// ------------------------------------------------------------------------------------------------ self-test (include/dmk_grid.h)
__global__ void dmk_selftest_special_kernel(long long n, unsigned long long seed, int lim, unsigned long long* bad) {
  unsigned long long nr = 0, ns = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long z = seed + 0x9e3779b97f4a7c15ull * (unsigned long long)(i + 1);       // splitmix64
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; z ^= z >> 31;
    const long long e = 1023 - lim + (long long)((z >> 52) % (unsigned long long)(2 * lim + 1));
    const double x = __longlong_as_double((long long)((z & 0x000fffffffffffffull) | ((unsigned long long)e << 52)));
    if (!pf_fast_ok(x)) continue;
    if (__double_as_longlong(pf_rcp_fast(x)) != __double_as_longlong(1.0 / x)) nr++;
    if (__double_as_longlong(pf_sqrt_fast(x)) != __double_as_longlong(sqrt(x))) ns++;
  }
  if (nr) atomicAdd(&bad[0], nr);
  if (ns) atomicAdd(&bad[1], ns);
}
extern "C" int dmk_selftest_special(long long n, unsigned long long seed, int lim, long long* bad_rcp, long long* bad_sqrt) {
  unsigned long long* d = nullptr;
  unsigned long long h[2] = {0, 0};
  DMK_CUDA_CHECK(cudaMalloc((void**)&d, sizeof h));
  DMK_CUDA_CHECK(cudaMemset(d, 0, sizeof h));
  dmk_selftest_special_kernel<<<1184, 256>>>(n, seed, lim, d);
  DMK_CUDA_CHECK(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
  cudaFree(d);
  DMK_CUDA_CHECK(cudaGetLastError());
  *bad_rcp = (long long)h[0]; *bad_sqrt = (long long)h[1];
  return DMK_OK;
}

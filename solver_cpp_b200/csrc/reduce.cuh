// Deterministic reductions over record fields (fixed launch geometry, fixed summation order) and the layout of the
// small device scratch area every context owns.
#pragma once
#include "ctx.h"

#define DMK_SCRATCH_DOUBLES (1 << 18)
#define DMK_SCRATCH_SUM9 0          // 16 doubles: result of dmk_sum_records
#define DMK_SCRATCH_TMP 64          // 448 doubles of call-local temporaries
#define DMK_SCRATCH_MECH 512        // 512 doubles: mechanical-solver device scalars (mech_solver.cu)
#define DMK_SCRATCH_PARTIAL 1024    // block partials
#define DMK_REDUCE_BLOCKS 1184      // 148 SMs x 8
#define DMK_REDUCE_THREADS 288      // 32 x 9: every thread always sees the same tensor component

// out9[c] = sum over records of field[rec*9 + c]   (stream-ordered, no host sync)
int dmk_sum_records(dmk_ctx* c, const double* d_field, long long n_records, double* d_out9);
// dot products of flat vectors: out[k] = sum_i a[i] * b_k[i], k < nvec (pointers in a device array)
__global__ void dmk_sum_partials_kernel(const double* __restrict__ partial, int nblocks, int width, double* __restrict__ out);
int dmk_multi_dot(dmk_ctx* c, const double* d_a, const double* const* d_bptrs_host, int nvec, long long n, double* d_out);

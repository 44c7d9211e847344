// Per-leaf truncation indicator of a Chebyshev tree, on the device.
//
// After tbslas::SolveSemilagInSitu the reference hands the refitted tree to PVFMM's
// RefineTree (reference src/tree/tree_utils.h:117-118), whose subdivision test looks at the
// size of the highest-degree coefficients of every leaf.  The refinement itself stays on the
// host in PVFMM (out of scope); this kernel only spares the host the download of all
// coefficients (C2: 443 MB) by returning 8 bytes per leaf:
//     tail[leaf] = sqrt( sum over dof and over i+j+k == q of  C[dof][i][j][k]^2 )
// (the l2 norm of the top shell of the packed triangular block, tree_functor.h:256-266; the
// exact PVFMM criterion is not vendored in the reference, so this is OUR definition, stated
// here, not a parity claim).  In the packed order the shell element of row (i,j) is its last
// one, k = q-i-j.  One warp per leaf; the q-dependent shell offsets come from the host.
#include <vector>

#include "common.cuh"

namespace tb {

__global__ void tail_norm_kernel(const double *__restrict__ coeff, size_t stride, int ncoef_pad, int dof,
                                 const int *__restrict__ shell, int n_shell, size_t n_leaf,
                                 double *__restrict__ out) {
  const size_t leaf = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (leaf >= n_leaf) return;
  const double *c = coeff + leaf * stride;
  double s = 0.0;
  for (int e = lane; e < n_shell * dof; e += 32) {
    const int l = e / n_shell, r = e - l * n_shell;
    const double v = c[(size_t)l * ncoef_pad + __ldg(shell + r)];
    s = fma(v, v, s);
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0) out[leaf] = sqrt(s);
}

int launch_tail_norm(tbslas_ctx *ctx, const tbslas_tree *t, double *out) {
  StageScope sc(ctx, ST_REFIT, (double)t->n_leaf, 1);
  if (!t->n_leaf) return TBSLAS_OK;
  const int d = t->q + 1;
  std::vector<int> shell;
  int off = 0;
  for (int i = 0; i < d; i++)
    for (int j = 0; i + j < d; j++) {
      off += d - i - j;         // row (i,j) holds k = 0 .. q-i-j
      shell.push_back(off - 1);  // its last element has i+j+k == q
    }
  void *buf;
  TB_TRY(ws_get(ctx, WS_MISC, sizeof(int) * shell.size(), &buf));
  // pageable source: the copy is staged before the call returns, the vector may go away
  TB_CUDA(ctx, cudaMemcpyAsync(buf, shell.data(), sizeof(int) * shell.size(), cudaMemcpyHostToDevice,
                               ctx->stream));
  const unsigned grid = (unsigned)((t->n_leaf + 7) / 8);
  tail_norm_kernel<<<grid, 256, 0, ctx->stream>>>(t->d_coeff, t->stride, (int)(t->stride / t->dof), t->dof,
                                                  (const int *)buf, (int)shell.size(), t->n_leaf, out);
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

}  // namespace tb

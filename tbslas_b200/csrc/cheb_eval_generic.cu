// Degree-generic Chebyshev evaluation (runtime q): covers the degrees that have no fully
// unrolled instantiation (15 <= q <= TBSLAS_MAX_CHEB_DEG; the reference allows q < 20,
// cheb.h:43, its scripts use q <= 14).  Same tiling, staging (bulk TMA of the leaf's
// coefficient block) and arithmetic as cheb_eval.cuh; the bases live in shared memory
// ([degree][thread], conflict free) instead of registers, so this path is LDS bound
// (about a quarter of the DFMA peak) -- correctness coverage, not the tuned kernel.
#include "cheb_eval.cuh"

namespace tb {

template <int EPI>
__global__ void __launch_bounds__(kEvalThreads)
cheb_eval_generic_kernel(const EvalParams p, int q) {
  extern __shared__ __align__(128) double s_mem[];
  __shared__ __align__(8) uint64_t s_bar;
  const int d = q + 1;
  double *s_coef = s_mem;                       // [stride]
  double *s_px = s_mem + p.stride;              // [d][kEvalThreads]
  double *s_py = s_px + d * kEvalThreads;
  double *s_pz = s_py + d * kEvalThreads;

  const unsigned n_tiles = __ldg(p.tile_start + p.n_bins);
  if (blockIdx.x >= n_tiles) return;
  const int2 tile = __ldg(p.tile_map + blockIdx.x);
  const int leaf = tile.x;
  const unsigned slot0 = (unsigned)tile.y;
  const unsigned cnt = min((unsigned)kEvalThreads, __ldg(p.bin_start + leaf + 1) - slot0);

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = p.stride * 8u;
    mbar_expect_tx(&s_bar, bytes);
    tma_bulk_g2s(s_coef, p.coeff + (size_t)leaf * p.stride, bytes, &s_bar);
  }
  const unsigned t = threadIdx.x;
  if ((t & ~31u) >= cnt) return;
  const bool ok = t < cnt;
  const unsigned idx = __ldg(p.perm + slot0 + (ok ? t : 0u));
  const double4 g = p.geom[leaf];
  const double *x = p.pos + 3 * (size_t)idx;
  const double xi[3] = {__dadd_rn(__dmul_rn(__dsub_rn(x[0], g.x), g.w), -1.0),
                        __dadd_rn(__dmul_rn(__dsub_rn(x[1], g.y), g.w), -1.0),
                        __dadd_rn(__dmul_rn(__dsub_rn(x[2], g.z), g.w), -1.0)};
  double *basis[3] = {s_px + t, s_py + t, s_pz + t};
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const bool in = fabs(xi[a]) <= 1.0;
    const double xc = in ? xi[a] : 0.0;
    double y0 = in ? 1.0 : 0.0, y1 = xc;
    basis[a][0] = y0;
    if (q >= 1) basis[a][kEvalThreads] = xc;
    for (int i = 2; i <= q; i++) {
      const double y2 = __dsub_rn(__dmul_rn(2.0 * xc, y1), y0);
      basis[a][i * kEvalThreads] = y2;
      y0 = y1;
      y1 = y2;
    }
  }
  mbar_wait(&s_bar, 0);
  for (int l = 0; l < p.dof; l++) {
    const double *C = s_coef + l * p.ncoef_pad;
    double u = 0.0;
    int ci = 0;
    for (int i = 0; i < d; i++) {
      double v = 0.0;
      for (int j = 0; j < d - i; j++) {
        double w = 0.0;
        for (int k = 0; k < d - i - j; k++) w = fma(s_px[k * kEvalThreads + t], C[ci++], w);
        v = fma(s_py[j * kEvalThreads + t], w, v);
      }
      u = fma(s_pz[i * kEvalThreads + t], v, u);
    }
    if (ok) {
      if (EPI == EPI_STORE) {
        p.out[(size_t)idx * p.dof + l] = u;
      } else {
        const size_t o = 3 * (size_t)idx + l;
        p.out[o] = __dadd_rn(p.base[o], __dmul_rn(p.alpha, u));
      }
    }
  }
}

int launch_cheb_eval_generic(tbslas_ctx *ctx, const EvalArgs &a) {
  const tbslas_tree *t = a.tree;
  EvalParams p;
  p.coeff = t->d_coeff;
  p.geom = t->d_geom;
  p.stride = (unsigned)t->stride;
  p.ncoef_pad = (unsigned)(t->stride / t->dof);
  p.dof = t->dof;
  p.n_bins = (int)t->n_leaf + 1;
  p.pos = a.pos;
  p.perm = a.perm;
  p.bin_start = a.bin_start;
  p.tile_start = a.tile_start;
  p.tile_map = a.tile_map;
  p.out = a.out;
  p.base = a.base;
  p.alpha = a.alpha;
  const size_t smem = (t->stride + 3 * (size_t)(t->q + 1) * kEvalThreads) * sizeof(double);
  if (smem > 227 * 1024)
    return fail(ctx, TBSLAS_ERR_UNSUPPORTED, "degree %d with dof %d needs %zu bytes of shared memory",
                t->q, t->dof, smem);
  const unsigned grid = (unsigned)a.max_tiles;
  if (grid == 0) return TBSLAS_OK;
  if (a.epilogue == EPI_STORE) {
    auto k = cheb_eval_generic_kernel<EPI_STORE>;
    TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kEvalThreads, smem, ctx->stream>>>(p, t->q);
  } else {
    auto k = cheb_eval_generic_kernel<EPI_AXPY>;
    TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kEvalThreads, smem, ctx->stream>>>(p, t->q);
  }
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

}  // namespace tb

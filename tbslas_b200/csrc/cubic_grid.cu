// Uniform-grid cubic interpolation: tbslas::fast_interp (reference
// src/tree/tree_functor.h:89-153).  Node-centred N_reg^3 grid on [0,1]^3 stored
// [dof][z][y][x]; a query outside the unit cube evaluates to 0 (:106-114); the 4^3
// Lagrange stencil starts at (int)(x*(N_reg-1)) - 1 clamped to [0, N_reg-4]
// (:118-124); weights and the 64-tap sum keep the reference's operation order
// (:126-151) with un-fused multiply/add, so results are bit-identical to the CPU.
//
// Roofline: gather bound.  Algorithmic bytes/point = 24 (xyz) + 8*dof (out) + 8*dof
// (each grid node read once when queries ~ nodes); neighbouring threads share stencil
// rows through L1/L2.
#include "common.cuh"

namespace tb {

struct LagrDen {
  double d[4];
};

// DOF > 0: compile-time dof (accumulators in registers, each tap weight M0*M1M2 formed once
// and shared by all components -- the same product the reference forms per component, so the
// sums stay bit-identical); DOF == 0: runtime dof, one component at a time.
template <int DOF>
__global__ void __launch_bounds__(256)
cubic_grid_kernel(const double *__restrict__ grid, int n_reg, int dof_rt, const double *__restrict__ pos,
                  size_t n, double *__restrict__ out, LagrDen den) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int dof = DOF > 0 ? DOF : dof_rt;
  const double x[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
  if (x[0] < 0 || x[0] > 1.0 || x[1] < 0 || x[1] > 1.0 || x[2] < 0 || x[2] > 1.0) {
    for (int k = 0; k < dof; k++) out[i * dof + k] = 0;
    return;
  }
  int g[3];
  double M[3][4];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    double pt = __dmul_rn(x[a], (double)(n_reg - 1));
    int gi = ((int)pt) - 1;
    gi = max(gi, 0);
    gi = min(gi, n_reg - 4);
    g[a] = gi;
    pt = __dsub_rn(pt, (double)gi);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double m = den.d[k];
#pragma unroll
      for (int l = 0; l < 4; l++)
        if (k != l) m = __dmul_rn(m, __dsub_rn(pt, (double)l));
      M[a][k] = m;
    }
  }
  const size_t n3 = (size_t)n_reg * n_reg * n_reg;
  if (DOF > 0) {
    double val[DOF > 0 ? DOF : 1];
#pragma unroll
    for (int k = 0; k < DOF; k++) val[k] = 0;
#pragma unroll
    for (int j2 = 0; j2 < 4; j2++) {
#pragma unroll
      for (int j1 = 0; j1 < 4; j1++) {
        const double m12 = __dmul_rn(M[1][j1], M[2][j2]);
        const double *row = grid + (size_t)n_reg * ((g[1] + j1) + (size_t)n_reg * (g[2] + j2)) + g[0];
#pragma unroll
        for (int j0 = 0; j0 < 4; j0++) {
          const double w = __dmul_rn(M[0][j0], m12);
#pragma unroll
          for (int k = 0; k < DOF; k++) val[k] = __dadd_rn(val[k], __dmul_rn(w, __ldg(row + k * n3 + j0)));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < DOF; k++) out[i * DOF + k] = val[k];
  } else {
    for (int k = 0; k < dof; k++) {
      const double *gk = grid + k * n3;
      double val = 0;
#pragma unroll
      for (int j2 = 0; j2 < 4; j2++) {
#pragma unroll
        for (int j1 = 0; j1 < 4; j1++) {
          const double m12 = __dmul_rn(M[1][j1], M[2][j2]);
          const double *row = gk + (size_t)n_reg * ((g[1] + j1) + (size_t)n_reg * (g[2] + j2)) + g[0];
#pragma unroll
          for (int j0 = 0; j0 < 4; j0++)
            val = __dadd_rn(val, __dmul_rn(__dmul_rn(M[0][j0], m12), __ldg(row + j0)));
        }
      }
      out[i * dof + k] = val;
    }
  }
}

int launch_cubic_grid(tbslas_ctx *ctx, const double *grid, int n_reg, int dof, const double *pos,
                      size_t n, double *out) {
  StageScope sc(ctx, ST_CUBIC, (double)n, 1);
  if (!n) return TBSLAS_OK;
  LagrDen den;
  for (int i = 0; i < 4; i++) {  // tree_functor.h:93-99, sequential IEEE divisions
    volatile double d = 1;
    for (int j = 0; j < 4; j++)
      if (i != j) d = d / (double)(i - j);
    den.d[i] = d;
  }
  const unsigned nb = (unsigned)((n + 255) / 256);
  switch (dof) {
    case 1: cubic_grid_kernel<1><<<nb, 256, 0, ctx->stream>>>(grid, n_reg, dof, pos, n, out, den); break;
    case 2: cubic_grid_kernel<2><<<nb, 256, 0, ctx->stream>>>(grid, n_reg, dof, pos, n, out, den); break;
    case 3: cubic_grid_kernel<3><<<nb, 256, 0, ctx->stream>>>(grid, n_reg, dof, pos, n, out, den); break;
    default: cubic_grid_kernel<0><<<nb, 256, 0, ctx->stream>>>(grid, n_reg, dof, pos, n, out, den); break;
  }
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

}  // namespace tb

// Values -> Chebyshev coefficients on the device: tbslas::SetTreeGridValues (reference
// src/tree/tree_utils.h:500-552), the step that follows the semi-Lagrangian evaluation in
// every time step.  It is one dense FP64 GEMM per tree,
//     coeff[L*dof][Ncoef] = vals[L*dof][P] * M[P][Ncoef],   P = (q+1)^3,
// with M the point-to-coefficient matrix (pseudo-inverse of the basis matrix at the
// new_nodes grid, tbslas::GetPt2CoeffMatrix, src/utils/cheb.h:166-196), which the caller
// supplies once per degree (tbslas_b200_set_pt2coeff) so that host and device use the very
// same matrix.  This is the one real dense contraction on the path (C2: 81 348 x 3375 x 680,
// 0.37 TFLOP), so it runs on the FP64 tensor path: mma.sync m8n8k4 (DMMA), 128x64 CTA tiles,
// 32x32 warp tiles (16 DMMA per 8 fragment loads), cp.async double-buffered shared memory.
//
// The reference's GEMM consumes vals as [leaf][dof][P] (dof-major inside a leaf) although
// SolveSemilagRK2 produces [point][dof] -- identical for dof = 1, and tree_ns.h:502-513
// transposes explicitly for dof = 3.  Both layouts are accepted here (`point_major`).
#include "common.cuh"

namespace tb {

constexpr int kBM = 128, kBN = 64, kBK = 16;
constexpr int kAp = kBK + 4;   // padded row of the A tile: conflict-free 8x4 fragment reads
constexpr int kBp = kBN + 4;   // padded row of the B tile: conflict-free 4x8 fragment reads
constexpr int kRefitThreads = 256;

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp8(void *dst, const void *src, bool pred) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = pred ? 8 : 0;  // src-size 0: zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp16(void *dst, const void *src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}

// A(row, k) = vals[(row / dof) * P * dof + (row % dof) * sd + k * sk]; B = M padded to
// [Kp][Np] (zeros), Kp % 16 == 0, Np % 64 == 0; C(row, n) -> coeff[row * ldc + n].
__global__ void __launch_bounds__(kRefitThreads)
refit_gemm_kernel(const double *__restrict__ vals, const double *__restrict__ Mp, double *__restrict__ coeff,
                  long long rows, int P, int dof, int sd, int sk, int N, int Np, int Kp, int ldc) {
  extern __shared__ __align__(16) double smem[];
  double *As = smem;                          // [2][kBM][kAp]
  double *Bs = smem + 2 * kBM * kAp;          // [2][kBK][kBp]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;  // 4 x 2 warps of 32 x 32
  const long long row0 = (long long)blockIdx.y * kBM;
  const int col0 = blockIdx.x * kBN;

  // A loader: thread -> (rows tid / 16 + 16 i, i = 0..7; k = tid % 16): the 16 lanes of a half-warp
  // fetch the 128 contiguous bytes of one row of the tile (rows are only 8-byte aligned, hence
  // 8-byte cp.async), so an instruction touches 2 rows x 4-5 sectors instead of 32 scattered ones.
  // B loader: (k = tid / 16, 4 consecutive n)
  const int ar = tid >> 4, ak = tid & 15;
  const double *abase[8];
  bool arow_ok[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const long long arow = row0 + ar + 16 * i;
    arow_ok[i] = arow < rows;
    abase[i] = arow_ok[i] ? vals + (arow / dof) * (long long)P * dof + (arow % dof) * (long long)sd : vals;
  }
  const int bk = tid >> 4, bn = (tid & 15) * 4;
  auto load_tiles = [&](int kt, int buf) {
    const int k0 = kt * kBK;
    const int k = k0 + ak;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const bool ok = arow_ok[i] && k < P;
      cp8(As + (buf * kBM + ar + 16 * i) * kAp + ak, abase[i] + (ok ? (long long)k * sk : 0), ok);
    }
    cp16(Bs + (buf * kBK + bk) * kBp + bn, Mp + (size_t)(k0 + bk) * Np + col0 + bn);
    cp16(Bs + (buf * kBK + bk) * kBp + bn + 2, Mp + (size_t)(k0 + bk) * Np + col0 + bn + 2);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = Kp / kBK;
  load_tiles(0, 0);
  for (int kt = 0; kt < nk; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < nk) {
      load_tiles(kt + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const double *a_s = As + (buf * kBM + wm + (lane >> 2)) * kAp + (lane & 3);
    const double *b_s = Bs + (buf * kBK + (lane & 3)) * kBp + wn + (lane >> 2);
#pragma unroll
    for (int k4 = 0; k4 < kBK; k4 += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = a_s[i * 8 * kAp + k4];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = b_s[k4 * kBp + j * 8];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
  // C fragment: row = lane / 4, cols = 2 * (lane % 4) + {0, 1}
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const long long r = row0 + wm + i * 8 + (lane >> 2);
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int c = col0 + wn + j * 8 + 2 * (lane & 3);
      if (c < N) coeff[r * ldc + c] = acc[i][j][0];
      if (c + 1 < N) coeff[r * ldc + c + 1] = acc[i][j][1];
    }
  }
}

int launch_refit(tbslas_ctx *ctx, tbslas_tree *t, const double *vals, int point_major) {
  const int q = t->q, d = q + 1, P = d * d * d, N = (int)t->ncoef;
  const Pt2Coeff &m = ctx->pt2coeff[q];
  if (!m.d_M) return fail(ctx, TBSLAS_ERR_INVALID, "no point-to-coefficient matrix for degree %d: call "
                                                   "tbslas_b200_set_pt2coeff first", q);
  const long long rows = (long long)t->n_leaf * t->dof;
  StageScope sc(ctx, ST_REFIT, (double)rows, 1);
  if (rows == 0) return TBSLAS_OK;
  const int sd = point_major ? 1 : P, sk = point_major ? t->dof : 1;
  const size_t smem = (2 * kBM * kAp + 2 * kBK * kBp) * sizeof(double);
  TB_CUDA(ctx, cudaFuncSetAttribute(refit_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)(m.Np / kBN), (unsigned)((rows + kBM - 1) / kBM));
  refit_gemm_kernel<<<grid, kRefitThreads, smem, ctx->stream>>>(vals, m.d_M, t->d_coeff, rows, P, t->dof, sd, sk, N,
                                                              m.Np, m.Kp, (int)(t->stride / t->dof));
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

int set_pt2coeff(tbslas_ctx *ctx, int q, const double *M_host) {
  const int d = q + 1, P = d * d * d, N = d * (d + 1) * (d + 2) / 6;
  Pt2Coeff &m = ctx->pt2coeff[q];
  m.Kp = (P + kBK - 1) / kBK * kBK;
  m.Np = (N + kBN - 1) / kBN * kBN;
  std::vector<double> padded((size_t)m.Kp * m.Np, 0.0);
  for (int k = 0; k < P; k++)
    for (int n = 0; n < N; n++) padded[(size_t)k * m.Np + n] = M_host[(size_t)k * N + n];
  if (m.d_M) {
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    TB_CUDA(ctx, cudaFree(m.d_M));
    m.d_M = nullptr;
  }
  TB_CUDA(ctx, cudaMalloc(&m.d_M, padded.size() * sizeof(double)));
  TB_CUDA(ctx, cudaMemcpyAsync(m.d_M, padded.data(), padded.size() * sizeof(double), cudaMemcpyHostToDevice,
                               ctx->stream));
  TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TBSLAS_OK;
}

}  // namespace tb

// Element-wise stages around the tree evaluations: cubic-in-time interpolation of
// four velocity snapshots, two-level extrapolation, RK2 position update, leaf-index
// fix-up.  All are HBM-bandwidth bound; arithmetic is written with explicit
// round-to-nearest multiply/add/divide in the reference's expression order so the
// results are bit-identical to the CPU path given identical inputs.
#include "common.cuh"

namespace tb {

// tbslas::CubicInterpPolicy::InterpCubic1D (reference src/utils/cubic.h:28-56).
struct CubicTimeW {
  double h00, h10d, h01, h11d;  // h10*(t2-t1), h11*(t2-t1) are formed per value below
  double d10, d21, d32;         // t1-t0, t2-t1, t3-t2
  double h10, h11;
};

__device__ __forceinline__ double tangent(double dk1, double dk, double pk_1, double pk, double pk1) {
  // (pk1-pk)*0.5/(tk1-tk) + (pk-pk_1)*0.5/(tk-tk_1)      cubic.h:36-39
  const double a = __ddiv_rn(__dmul_rn(__dsub_rn(pk1, pk), 0.5), dk1);
  const double b = __ddiv_rn(__dmul_rn(__dsub_rn(pk, pk_1), 0.5), dk);
  return __dadd_rn(a, b);
}

template <bool AXPY>
__global__ void cubic_time_kernel(const double *__restrict__ v0, const double *__restrict__ v1,
                                  const double *__restrict__ v2, const double *__restrict__ v3,
                                  size_t m, CubicTimeW w, double *__restrict__ out,
                                  const double *__restrict__ base, double alpha) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double p0 = v0[i], p1 = v1[i], p2 = v2[i], p3 = v3[i];
  const double mk = tangent(w.d21, w.d10, p0, p1, p2);
  const double mk1 = tangent(w.d32, w.d21, p1, p2, p3);
  // h00*p1 + h10*(x2-x1)*mk + h01*p2 + h11*(x2-x1)*mk1, left to right   cubic.h:49-52
  double val = __dmul_rn(w.h00, p1);
  val = __dadd_rn(val, __dmul_rn(__dmul_rn(w.h10, w.d21), mk));
  val = __dadd_rn(val, __dmul_rn(w.h01, p2));
  val = __dadd_rn(val, __dmul_rn(__dmul_rn(w.h11, w.d21), mk1));
  out[i] = AXPY ? __dadd_rn(base[i], __dmul_rn(alpha, val)) : val;
}

int launch_cubic_time(tbslas_ctx *ctx, const double *v4, size_t m, const double times[4], double t,
                      double *out, const double *base, double alpha, int axpy) {
  StageScope sc(ctx, ST_COMBINE, (double)m, 1);
  if (!m) return TBSLAS_OK;
  // Hermite basis at the (uniform) query time; host arithmetic, un-contracted:
  // volatile keeps every intermediate a rounded double whatever -ffp-contract says.
  CubicTimeW w;
  volatile double d21 = times[2] - times[1];
  volatile double tt = (t - times[1]) / d21;
  volatile double t2 = tt * tt, t3;
  {
    volatile double a = 2 * tt;  // 2*t*t*t = ((2*t)*t)*t
    volatile double b = a * tt;
    volatile double c = b * tt;
    volatile double d = 3 * tt;  // 3*t*t = (3*t)*t
    volatile double e = d * tt;
    volatile double f = c - e;
    w.h00 = f + 1;
    t3 = t2 * tt;  // t*t*t = (t*t)*t
    volatile double g = 2 * tt;
    volatile double h = g * tt;  // 2*t*t
    volatile double k = t3 - h;
    w.h10 = k + tt;
    volatile double n0 = -2 * tt;  // -2*t*t*t = ((-2*t)*t)*t
    volatile double n1 = n0 * tt;
    volatile double n2 = n1 * tt;
    w.h01 = n2 + e;
    w.h11 = t3 - t2;
  }
  w.d10 = times[1] - times[0];
  w.d21 = d21;
  w.d32 = times[3] - times[2];
  w.h10d = w.h11d = 0;
  const unsigned grid = (unsigned)((m + 255) / 256);
  if (axpy)
    cubic_time_kernel<true><<<grid, 256, 0, ctx->stream>>>(v4, v4 + m, v4 + 2 * m, v4 + 3 * m, m, w,
                                                           out, base, alpha);
  else
    cubic_time_kernel<false><<<grid, 256, 0, ctx->stream>>>(v4, v4 + m, v4 + 2 * m, v4 + 3 * m, m,
                                                            w, out, base, alpha);
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

// Hermite-in-time weights of the four snapshots: InterpCubic1D (cubic.h:36-56) is linear in
// p0..p3, out = w0 p0 + w1 p1 + w2 p2 + w3 p3 with the w's below (host arithmetic).
void cubic_time_weights(const double times[4], double t, double w[4]) {
  const double d10 = times[1] - times[0], d21 = times[2] - times[1], d32 = times[3] - times[2];
  const double tt = (t - times[1]) / d21, t2 = tt * tt, t3 = t2 * tt;
  const double h00 = 2 * t3 - 3 * t2 + 1, h10 = t3 - 2 * t2 + tt, h01 = -2 * t3 + 3 * t2, h11 = t3 - t2;
  const double H10 = h10 * d21, H11 = h11 * d21;  // multiply the tangents m1, m2
  // m1 = (p2-p1)*0.5/d21 + (p1-p0)*0.5/d10,  m2 = (p3-p2)*0.5/d32 + (p2-p1)*0.5/d21
  w[0] = -H10 * 0.5 / d10;
  w[1] = h00 + H10 * (0.5 / d10 - 0.5 / d21) - H11 * 0.5 / d21;
  w[2] = h01 + H10 * 0.5 / d21 + H11 * (0.5 / d21 - 0.5 / d32);
  w[3] = H11 * 0.5 / d32;
}

// Coefficients of the field  sum_k w_k * tree_k  for trees that share their leaf list: the
// evaluation is linear in the coefficients, so ONE evaluation of the combined block replaces
// one evaluation per tree followed by a per-point combination.
__global__ void combine_coeff_kernel(const double *__restrict__ c0, const double *__restrict__ c1,
                                     const double *__restrict__ c2, const double *__restrict__ c3,
                                     double w0, double w1, double w2, double w3, int n_tree, size_t m,
                                     double *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double v = __dadd_rn(__dmul_rn(w0, c0[i]), __dmul_rn(w1, c1[i]));
  if (n_tree == 4) v = __dadd_rn(__dadd_rn(v, __dmul_rn(w2, c2[i])), __dmul_rn(w3, c3[i]));
  out[i] = v;
}

int launch_combine_coeff(tbslas_ctx *ctx, const double *const c[4], const double w[4], int n_tree, size_t m,
                         double *out) {
  StageScope sc(ctx, ST_COMBINE, (double)m, 1);
  if (!m) return TBSLAS_OK;
  combine_coeff_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(
      c[0], c[1], n_tree == 4 ? c[2] : c[0], n_tree == 4 ? c[3] : c[0], w[0], w[1], w[2], w[3], n_tree, m, out);
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

// tbslas::FieldExtrapFunctor (reference src/tree/tree_extrap_functor.h:70-77):
// out = 1.5*vc - 0.5*vp.
template <bool AXPY>
__global__ void extrap_kernel(const double *__restrict__ vc, const double *__restrict__ vp, size_t m,
                              double *__restrict__ out, const double *__restrict__ base,
                              double alpha) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double val = __dsub_rn(__dmul_rn(1.5, vc[i]), __dmul_rn(0.5, vp[i]));
  out[i] = AXPY ? __dadd_rn(base[i], __dmul_rn(alpha, val)) : val;
}

int launch_extrap(tbslas_ctx *ctx, const double *vc, const double *vp, size_t m, double *out,
                  const double *base, double alpha, int axpy) {
  StageScope sc(ctx, ST_COMBINE, (double)m, 1);
  if (!m) return TBSLAS_OK;
  const unsigned grid = (unsigned)((m + 255) / 256);
  if (axpy)
    extrap_kernel<true><<<grid, 256, 0, ctx->stream>>>(vc, vp, m, out, base, alpha);
  else
    extrap_kernel<false><<<grid, 256, 0, ctx->stream>>>(vc, vp, m, out, base, alpha);
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

// x' = x + alpha*v (reference src/semilag/traj.inc:34-36,41-42), multiply then add.
__global__ void axpy_kernel(const double *__restrict__ base, const double *__restrict__ v,
                            double alpha, size_t m, double *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[i] = __dadd_rn(base[i], __dmul_rn(alpha, v[i]));
}

int launch_axpy(tbslas_ctx *ctx, const double *base, const double *v, double alpha, size_t m,
                double *out) {
  StageScope sc(ctx, ST_COMBINE, (double)m, 1);
  if (!m) return TBSLAS_OK;
  axpy_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(base, v, alpha, m, out);
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

// internal leaf ids -> what the boundary reports: global Morton index, -1 = no leaf.
__global__ void leaf_fixup_kernel(int32_t *leaf, size_t n, int n_leaf, long long offset) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = leaf[i];
  if (j < 0) return;  // outsider marker, resolved by the exchange path
  leaf[i] = (j >= n_leaf) ? -1 : (int32_t)(j + offset);
}

// per-leaf point counts of an evaluation, kept with the tree (tbslas_b200_tree_last_point_counts)
__global__ void keep_counts_kernel(const uint32_t *__restrict__ count, uint32_t *__restrict__ keep, size_t n,
                                   int add) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) keep[j] = (add ? keep[j] : 0u) + count[j];
}

int launch_keep_counts(tbslas_ctx *ctx, const uint32_t *count, uint32_t *keep, size_t n_leaf, bool add) {
  if (!n_leaf) return TBSLAS_OK;
  keep_counts_kernel<<<(unsigned)((n_leaf + 255) / 256), 256, 0, ctx->stream>>>(count, keep, n_leaf, add ? 1 : 0);
  TB_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return TBSLAS_OK;
}

int launch_leaf_fixup(tbslas_ctx *ctx, int32_t *leaf, size_t n, size_t n_leaf, long long offset) {
  StageScope sc(ctx, ST_COMBINE, (double)n, 1);
  if (!n) return TBSLAS_OK;
  leaf_fixup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(leaf, n, (int)n_leaf,
                                                                          offset);
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

}  // namespace tb

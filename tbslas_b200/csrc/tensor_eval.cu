// Velocity at the ARRIVAL points by sum factorisation.
//
// The first evaluation of a semi-Lagrangian step samples the velocity at the arrival points,
// and in the tree-level call (tbslas::SolveSemilagInSitu, reference src/tree/tree_semilag.h:
// 103-124) those are not arbitrary: they are the (q+1)^3 tensor grid of every leaf of the
// advected tree (CollectChebTreeGridPoints, tree_utils.h:442-498).  Where such a leaf lies inside
// ONE leaf of the velocity tree, the local coordinates of its grid are a tensor product too,
//     xi(px) x eta(py) x zeta(pz),
// and the triangular Chebyshev sum factorises into three 1-D passes
//     A[i][j][px] = sum_k C[i][j][k] T_k(xi_px)            (120 rows x 15, <= 15 terms)
//     B[i][py][px] = sum_j A[i][j][px] T_j(eta_py)          (15 x 225, <= 15 terms)
//     u[pz][py][px] = sum_i B[i][py][px] T_i(zeta_pz)       (3375, 15 terms)
// = 88 k FMA per leaf and component at q = 14 against 3375 x 679 = 2.29 M for point-by-point
// evaluation (tree_functor.h:27-84): the same polynomial at the same points, summed in another
// order (differences ~1e-15 of the field scale).  The coordinates and bases are formed exactly
// as the generic kernels form them (same expressions, same rounding).
//
// Leaf assignment stays the reference's (tree_functor.h:190-198): a grid point belongs to the
// velocity leaf that contains the leaf's box only if its depth-15 anchor lies inside that
// leaf's integer box -- points ON an upper face belong to the neighbour (or wrap, or leave the
// domain).  Those points, and all points of leaves that are not inside one velocity leaf, are
// listed as EXCEPTIONS and evaluated by the generic locate/evaluate path afterwards.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cheb_eval.cuh"  // mbarrier / bulk-TMA helpers
#include "common.cuh"
#include "gridbase.cuh"
#include "keys.cuh"

namespace tb {

constexpr int kTensorThreads = 256;
#ifndef TB_TENSOR_MINB
#define TB_TENSOR_MINB 2
#endif
#ifndef TB_TENSOR_XSTORE_CS
#define TB_TENSOR_XSTORE_CS 0  // 1: arrival points by streaming stores (st.global.cs)
#endif
#if TB_TENSOR_XSTORE_CS
#define TB_TENSOR_XSTORE(ptr, v) __stcs((ptr), (v))
#else
#define TB_TENSOR_XSTORE(ptr, v) (*(ptr) = (v))
#endif
constexpr int kMaxD = TBSLAS_MAX_CHEB_DEG + 1;
constexpr int kMaxRows = kMaxD * (kMaxD + 1) / 2;

struct TensorTables {
  double node[kMaxD];          // tbslas::new_nodes, 1-D
  uint16_t row_off[kMaxRows];  // first coefficient of row (i,j) in the packed block
  uint16_t row_first[kMaxD + 1];  // first row of plane i
  uint8_t row_len[kMaxRows];   // terms of row (i,j): d - i - j
  uint8_t tile_k[(kMaxRows + 7) / 8];  // longest row of every 8-row tile (DMMA kernel: k-steps of pass 1)
};

// Ghost leaves (Morton-sharded velocity tree): the LAST leaf of every rank, all-gathered per call.  An
// advected leaf whose containing velocity leaf lives on another rank always lies under that rank's last
// leaf (leaves are disjoint and Morton ordered, and both trees share the split keys), so with these few
// records the sum-factorised path covers exactly the leaves it covers when the velocity tree is held
// whole -- instead of sending all their points through the generic, exchanged evaluation.  Record of
// rank r at ghost + r*ghost_rec: {uint4 box; double4 geom; u64 n_leaf; u64 pad; double coeff[vstride]}.
constexpr size_t kGhostHdr = 64;
struct TensorParams {
  const double *vcoeff;     // velocity coefficients, [leaf][dof][ncoef_pad]
  const double4 *vgeom;
  const uint4 *vbox;
  int v_nleaf;              // local velocity leaves; map values >= v_nleaf name the ghost of rank (j - v_nleaf)
  const char *ghost;        // nullptr: no ghosts
  size_t ghost_rec;
  unsigned vstride, ncoef_pad;
  const double4 *ggeom;     // grid tree, already offset to the first leaf of the range
  const uint8_t *gdepth;
  const int32_t *map;
  size_t n_leaf;
  int d, periodic;
  const double *x;          // [n_leaf * P][3] the grid points
  double *xgen;             // != nullptr: the grid points are WRITTEN here first (== x): the kernel is
                            // also tbslas::CollectChebTreeGridPoints (gridpts.cu) for these leaves
  double *out;              // [n_leaf * P][3] = x + alpha * v on regular points
  double alpha;
  unsigned *exc_count;      // exceptions: number and point ids
  uint32_t *exc_idx;
};

__device__ __forceinline__ const double *vel_coeff(const TensorParams &p, int j) {
  return j < p.v_nleaf ? p.vcoeff + (size_t)j * p.vstride
                       : reinterpret_cast<const double *>(p.ghost + (size_t)(j - p.v_nleaf) * p.ghost_rec + kGhostHdr);
}
__device__ __forceinline__ double4 vel_geom(const TensorParams &p, int j) {
  return j < p.v_nleaf ? p.vgeom[j]
                       : *reinterpret_cast<const double4 *>(p.ghost + (size_t)(j - p.v_nleaf) * p.ghost_rec + 16);
}
__device__ __forceinline__ uint4 vel_box(const TensorParams &p, int j) {
  return j < p.v_nleaf ? p.vbox[j] : *reinterpret_cast<const uint4 *>(p.ghost + (size_t)(j - p.v_nleaf) * p.ghost_rec);
}

// this rank's last leaf -> its ghost record (before the all-gather)
__global__ void ghost_pack_kernel(const uint4 *__restrict__ vbox, const double4 *__restrict__ vgeom,
                                  const double *__restrict__ vcoeff, int v_nleaf, unsigned vstride, char *rec) {
  const int t = threadIdx.x;
  if (t == 0) {
    *reinterpret_cast<uint4 *>(rec) = v_nleaf ? vbox[v_nleaf - 1] : make_uint4(0, 0, 0, 0);
    *reinterpret_cast<double4 *>(rec + 16) = v_nleaf ? vgeom[v_nleaf - 1] : make_double4(0, 0, 0, 2.0);
    *reinterpret_cast<unsigned long long *>(rec + 48) = (unsigned long long)v_nleaf;
    *reinterpret_cast<unsigned long long *>(rec + 56) = 0ull;
  }
  double *c = reinterpret_cast<double *>(rec + kGhostHdr);
  for (unsigned e = t; e < vstride; e += blockDim.x) c[e] = v_nleaf ? vcoeff[(size_t)(v_nleaf - 1) * vstride + e] : 0.0;
}

// grid leaf -> velocity leaf that contains its box (local index, or v_nleaf + r for the ghost of rank r), or -1
__global__ void grid_leaf_map_kernel(const uint4 *__restrict__ gbox, size_t n_leaf,
                                     const uint64_t *__restrict__ vkeys, const uint4 *__restrict__ vbox,
                                     const uint32_t *__restrict__ vcell, int vshift, int v_nleaf,
                                     int32_t *__restrict__ map, const char *__restrict__ ghost, size_t ghost_rec,
                                     int nranks, int me) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_leaf) return;
  const uint4 b = gbox[c];
  const uint64_t key = anchor_key(b.x, b.y, b.z);
  const unsigned cell = (unsigned)(key >> vshift);
  int lo = (int)__ldg(vcell + cell), hi = (int)__ldg(vcell + cell + 1);
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(vkeys + mid) <= key)
      lo = mid + 1;
    else
      hi = mid;
  }
  const int j = lo - 1;
  int r = -1;
  if (j >= 0 && j < v_nleaf) {
    const uint4 v = __ldg(vbox + j);
    if (b.w <= v.w && ((((b.x ^ v.x) | (b.y ^ v.y) | (b.z ^ v.z)) >> v.w) == 0u)) r = j;
  }
  if (r < 0 && ghost) {
    for (int g = 0; g < nranks; g++) {
      if (g == me) continue;
      const char *rec = ghost + (size_t)g * ghost_rec;
      if (*reinterpret_cast<const unsigned long long *>(rec + 48) == 0ull) continue;  // that rank holds no leaf
      const uint4 v = *reinterpret_cast<const uint4 *>(rec);
      if (b.w <= v.w && ((((b.x ^ v.x) | (b.y ^ v.y) | (b.z ^ v.z)) >> v.w) == 0u)) r = v_nleaf + g;
    }
  }
  map[c] = r;
}

__global__ void __launch_bounds__(kTensorThreads)
tensor_grid_eval_kernel(const TensorParams p, const TensorTables tb_) {
  extern __shared__ __align__(16) double sm[];
  const int d = p.d, dp = d | 1, P2 = d * d, P = P2 * d, n_row = d * (d + 1) / 2;
  double *sT = sm;                       // [3][d][dp]   T_k at the leaf's grid, per axis
  double *sC = sT + 3 * d * dp;          // [ncoef_pad]
  double *sA = sC + p.ncoef_pad;         // [n_row][dp]
  double *sB = sA + n_row * dp;          // [d][P2]
  __shared__ unsigned s_ok[3];           // bit i: node i of the axis lies inside the velocity leaf's box
  __shared__ unsigned s_exc_base, s_exc_n;
  const int t = threadIdx.x;
  for (size_t leaf = blockIdx.x; leaf < p.n_leaf; leaf += gridDim.x) {
    const int j = p.map[leaf];
    const size_t gp0 = leaf * (size_t)P;
    __syncthreads();  // previous leaf done with the shared arrays
    if (p.xgen) {  // arrival points of this leaf, formed exactly as gridpts.cu forms them
      const double4 gc = p.ggeom[leaf];
      const double glen = 1.0 / (double)(1u << p.gdepth[leaf]);
      double *o = p.xgen + 3 * gp0;
      for (int e = t; e < 3 * P; e += kTensorThreads) {
        const int pt = e / 3, a = e - 3 * pt;
        const int pz = pt / (d * d), rem = pt - pz * d * d, py = rem / d, px = rem - py * d;
        const double c = a == 0 ? gc.x : (a == 1 ? gc.y : gc.z);
        o[e] = __dadd_rn(c, __dmul_rn(glen, tb_.node[a == 0 ? px : (a == 1 ? py : pz)]));
      }
    }
    if (t < 3) s_ok[t] = 0u;
    if (t == 0) s_exc_n = 0u;
    __syncthreads();
    if (j >= 0 && t < 3 * d) {
      const int a = t / d, i = t - a * d;
      const double4 gc = p.ggeom[leaf], gv = vel_geom(p, j);
      const uint4 vb = vel_box(p, j);
      const double len = 1.0 / (double)(1u << p.gdepth[leaf]);
      const double c = a == 0 ? gc.x : (a == 1 ? gc.y : gc.z), vc = a == 0 ? gv.x : (a == 1 ? gv.y : gv.z);
      const unsigned vba = a == 0 ? vb.x : (a == 1 ? vb.y : vb.z);
      const double x = __dadd_rn(c, __dmul_rn(len, tb_.node[i]));            // gridpts.cu
      const double xi = __dadd_rn(__dmul_rn(__dsub_rn(x, vc), gv.w), -1.0);  // cheb_eval.cuh
      const bool in = fabs(xi) <= 1.0;
      const double xc = in ? xi : 0.0, x2 = 2.0 * xc;
      double t0 = in ? 1.0 : 0.0, t1 = xc;
      double *T = sT + a * d * dp + i;
      T[0] = t0;
      if (d > 1) T[dp] = t1;
      for (int k = 2; k < d; k++) {
        const double t2 = __dsub_rn(__dmul_rn(x2, t1), t0);
        T[k * dp] = t2;
        t0 = t1;
        t1 = t2;
      }
      // anchor of the point along this axis (locate.cu): inside the velocity leaf's box?
      const double xs = x * 32768.0;
      int jx = __double2int_rd(xs);
      if (!p.periodic && xs == 32768.0) jx = 32767;
      if ((unsigned)jx < 32768u && ((((unsigned)jx ^ vba) >> vb.w) == 0u)) atomicOr(&s_ok[a], 1u << i);
    }
    __syncthreads();
    const unsigned okx = s_ok[0], oky = s_ok[1], okz = s_ok[2];
    const unsigned n_reg = (unsigned)(__popc(okx) * __popc(oky) * __popc(okz));
    if (t == 0 && n_reg < (unsigned)P) s_exc_base = atomicAdd(p.exc_count, (unsigned)P - n_reg);
    if (j >= 0 && n_reg) {
      for (int l = 0; l < 3; l++) {
        const double *C = vel_coeff(p, j) + (size_t)l * p.ncoef_pad;
        for (int e = t; e < (int)p.ncoef_pad; e += kTensorThreads) sC[e] = C[e];
        __syncthreads();
        // pass 1: rows (i,j) contracted with T_k(x) -- A[row][px]
        for (int e = t; e < n_row * d; e += kTensorThreads) {
          const int r = e / d, px = e - r * d;
          // plane of the row: rows of plane i are row_first[i] .. row_first[i+1]-1, row length d-i-jj
          int i = 0;
          while (tb_.row_first[i + 1] <= r) i++;
          const int len_r = d - i - (r - tb_.row_first[i]);
          const double *c = sC + tb_.row_off[r], *T = sT + px;
          double acc = 0.0;
          for (int k = 0; k < len_r; k++) acc = fma(c[k], T[k * dp], acc);
          sA[r * dp + px] = acc;
        }
        __syncthreads();
        // pass 2: B[i][py][px] = sum_jj A[(i,jj)][px] T_jj(y_py)
        for (int e = t; e < P; e += kTensorThreads) {
          const int i = e / P2, rem = e - i * P2, py = rem / d, px = rem - py * d;
          const double *A = sA + tb_.row_first[i] * dp + px, *T = sT + d * dp + py;
          double acc = 0.0;
          for (int jj = 0; jj < d - i; jj++) acc = fma(A[jj * dp], T[jj * dp], acc);
          sB[e] = acc;
        }
        __syncthreads();
        // pass 3: u[pz][py][px] = sum_i B[i][py][px] T_i(z_pz); x' = x + alpha * u on regular points
        for (int e = t; e < P; e += kTensorThreads) {
          const int pz = e / P2, rem = e - pz * P2, py = rem / d, px = rem - py * d;
          if (!(((okx >> px) & (oky >> py) & (okz >> pz)) & 1u)) continue;
          const double *B = sB + rem, *T = sT + 2 * d * dp + pz;
          double acc = 0.0;
          for (int i = 0; i < d; i++) acc = fma(B[i * P2], T[i * dp], acc);
          const size_t o = 3 * (gp0 + e) + l;
          p.out[o] = __dadd_rn(p.x[o], __dmul_rn(p.alpha, acc));  // traj.inc:36
        }
        __syncthreads();
      }
    }
    if (n_reg < (unsigned)P) {  // list the exceptions of this leaf (CTA-uniform condition)
      __syncthreads();          // s_exc_base is visible
      for (int e = t; e < P; e += kTensorThreads) {
        const int pz = e / P2, rem = e - pz * P2, py = rem / d, px = rem - py * d;
        if (j >= 0 && (((okx >> px) & (oky >> py) & (okz >> pz)) & 1u)) continue;
        const unsigned k = atomicAdd(&s_exc_n, 1u);
        p.exc_idx[s_exc_base + k] = (uint32_t)(gp0 + e);
      }
    }
  }
}

// The same for a compile-time degree (q <= 14): the three passes are register blocked so that the
// FP64 pipe, not shared memory, bounds them.
//   passes 1+2, thread = (plane i, px): T_k(x_px) in 15 registers; for every row j of the plane
//     a_j = sum_k C[i][j][k] T_k(x_px) (coefficients: shared-memory reads shared by the px lanes),
//     kept in registers; then B[i][py][px] = sum_j a_j T_j(y_py) for the 15 py (T_j(y_py) read as
//     warp-wide broadcasts).  The intermediate A never touches shared memory.
//   pass 3, thread = column (py,px): B[0..q][py][px] in 15 registers, u[pz] = sum_i B_i T_i(z_pz)
//     for the 15 pz with broadcast reads of T_i(z_pz); epilogue x' = x + alpha u on regular points.
template <int D>
__global__ void __launch_bounds__(kTensorThreads, TB_TENSOR_MINB)
tensor_grid_eval_kernel_t(const TensorParams p, const TensorTables tb_) {
  constexpr int DP = D | 1, DJ = (D + 1) & ~1, P2 = D * D, P = P2 * D, TOTAL = D * (D + 1) * (D + 2) / 6;
  static_assert(P2 <= kTensorThreads, "one thread per (py,px) column");
  extern __shared__ __align__(16) double sm[];
  double *sT0 = sm;                  // [D][DP]  T_k(x_px), degree major
  double *sT1 = sT0 + D * DP + (D * DP & 1);  // [D][DJ]  T_j(y_py), point major (rows 16-byte aligned)
  double *sT2 = sT1 + D * DJ;        // [D][DJ]  T_i(z_pz), point major
  double *sC3 = sT2 + D * DJ;        // [3][ncoef_pad] the velocity leaf's block, all components
  double *sB = sC3 + p.vstride;      // [D][P2]
  __shared__ unsigned s_ok[3];
  __shared__ unsigned s_exc_base, s_exc_n;
  const int t = threadIdx.x;
  for (size_t leaf = blockIdx.x; leaf < p.n_leaf; leaf += gridDim.x) {
    const int j = p.map[leaf];
    const size_t gp0 = leaf * (size_t)P;
    __syncthreads();
    if (p.xgen) {  // arrival points of this leaf, formed exactly as gridpts.cu forms them
      const double4 gc = p.ggeom[leaf];
      const double glen = 1.0 / (double)(1u << p.gdepth[leaf]);
      double *o = p.xgen + 3 * gp0;
      for (int e = t; e < 3 * P; e += kTensorThreads) {
        const int pt = e / 3, a = e - 3 * pt;
        const int pz = pt / (D * D), rem = pt - pz * D * D, py = rem / D, px = rem - py * D;
        const double c = a == 0 ? gc.x : (a == 1 ? gc.y : gc.z);
        o[e] = __dadd_rn(c, __dmul_rn(glen, tb_.node[a == 0 ? px : (a == 1 ? py : pz)]));
      }
    }
    if (t < 3) s_ok[t] = 0u;
    if (t == 0) s_exc_n = 0u;
    if (j >= 0) {  // in flight while the bases are built
      const double *C = vel_coeff(p, j);
      for (int e = t; e < (int)p.vstride; e += kTensorThreads) sC3[e] = C[e];
    }
    __syncthreads();
    if (j >= 0 && t < 3 * D) {
      const int a = t / D, i = t - a * D;
      const double4 gc = p.ggeom[leaf], gv = vel_geom(p, j);
      const uint4 vb = vel_box(p, j);
      const double len = 1.0 / (double)(1u << p.gdepth[leaf]);
      const double c = a == 0 ? gc.x : (a == 1 ? gc.y : gc.z), vc = a == 0 ? gv.x : (a == 1 ? gv.y : gv.z);
      const unsigned vba = a == 0 ? vb.x : (a == 1 ? vb.y : vb.z);
      const double x = __dadd_rn(c, __dmul_rn(len, tb_.node[i]));            // gridpts.cu
      const double xi = __dadd_rn(__dmul_rn(__dsub_rn(x, vc), gv.w), -1.0);  // cheb_eval.cuh
      const bool in = fabs(xi) <= 1.0;
      const double xc = in ? xi : 0.0, x2 = 2.0 * xc;
      double t0 = in ? 1.0 : 0.0, t1 = xc;
      // axis 0: [degree][point]; axes 1, 2: [point][degree]
      double *T = a == 0 ? sT0 + i : (a == 1 ? sT1 : sT2) + i * DJ;
      const int st = a == 0 ? DP : 1;
      T[0] = t0;
      if (D > 1) T[st] = t1;
#pragma unroll
      for (int k = 2; k < D; k++) {
        const double t2 = __dsub_rn(__dmul_rn(x2, t1), t0);
        T[k * st] = t2;
        t0 = t1;
        t1 = t2;
      }
      if (a != 0 && DJ > D) T[D] = 0.0;  // padding read by the paired loads
      const double xs = x * 32768.0;
      int jx = __double2int_rd(xs);
      if (!p.periodic && xs == 32768.0) jx = 32767;
      if ((unsigned)jx < 32768u && ((((unsigned)jx ^ vba) >> vb.w) == 0u)) atomicOr(&s_ok[a], 1u << i);
    }
    __syncthreads();
    const unsigned okx = s_ok[0], oky = s_ok[1], okz = s_ok[2];
    const unsigned n_reg = (unsigned)(__popc(okx) * __popc(oky) * __popc(okz));
    if (t == 0 && n_reg < (unsigned)P) s_exc_base = atomicAdd(p.exc_count, (unsigned)P - n_reg);
    if (j >= 0 && n_reg) {
      const int pi = t / D, px = t - pi * D;      // passes 1+2: (plane, px)
      const int py3 = t / D, px3 = t - py3 * D;   // pass 3: column (py, px) == t
      double tx[D];
      if (t < P2) {
#pragma unroll
        for (int k = 0; k < D; k++) tx[k] = sT0[k * DP + px];
      }
      // the grid point of (px3, py3, pz), recomputed as gridpts.cu forms it (bit-identical)
      const double4 gc = p.ggeom[leaf];
      const double glen = 1.0 / (double)(1u << p.gdepth[leaf]);
      const double xq = __dadd_rn(gc.x, __dmul_rn(glen, tb_.node[px3 < D ? px3 : 0]));
      const double yq = __dadd_rn(gc.y, __dmul_rn(glen, tb_.node[py3 < D ? py3 : 0]));
      for (int l = 0; l < 3; l++) {
        const double *sC = sC3 + l * p.ncoef_pad;
        if (t < P2) {
          const int m = D - pi;                                   // rows of plane pi
          const int base = TOTAL - m * (m + 1) * (m + 2) / 6;     // its first coefficient
          double a[D];
#pragma unroll
          for (int jj = 0; jj < D; jj++) {
            const int len = m - jj;                               // <= 0: the row does not exist
            const double *c = sC + base + jj * m - jj * (jj - 1) / 2;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < D - jj; k++)
              if (k < len) acc = fma(c[k], tx[k], acc);
            a[jj] = acc;
          }
#pragma unroll 3
          for (int py = 0; py < D; py++) {
            const double *Ty = sT1 + py * DJ;
            double acc = 0.0;
#pragma unroll
            for (int jj = 0; jj < D; jj++) acc = fma(a[jj], Ty[jj], acc);  // a[jj] = 0 past the plane
            sB[pi * P2 + py * D + px] = acc;
          }
        }
        __syncthreads();
        if (t < P2) {
          double b[D];
#pragma unroll
          for (int i = 0; i < D; i++) b[i] = sB[i * P2 + t];
          const bool col_ok = ((okx >> px3) & (oky >> py3)) & 1u;
#pragma unroll 3
          for (int pz = 0; pz < D; pz++) {
            const double *Tz = sT2 + pz * DJ;
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < D; i++) acc = fma(b[i], Tz[i], acc);
            if (col_ok && ((okz >> pz) & 1u)) {
              const size_t o = 3 * (gp0 + (size_t)pz * P2 + t) + l;
              const double x0 = l == 0 ? xq : (l == 1 ? yq : __dadd_rn(gc.z, __dmul_rn(glen, tb_.node[pz])));
              p.out[o] = __dadd_rn(x0, __dmul_rn(p.alpha, acc));  // traj.inc:36
            }
          }
        }
        __syncthreads();
      }
    }
    if (n_reg < (unsigned)P) {  // list the exceptions of this leaf (CTA-uniform condition)
      __syncthreads();
      for (int e = t; e < P; e += kTensorThreads) {
        const int pz = e / P2, rem = e - pz * P2, py = rem / D, px = rem - py * D;
        if (j >= 0 && (((okx >> px) & (oky >> py) & (okz >> pz)) & 1u)) continue;
        const unsigned k = atomicAdd(&s_exc_n, 1u);
        p.exc_idx[s_exc_base + k] = (uint32_t)(gp0 + e);
      }
    }
  }
}

template <int D>
static int launch_tensor_t(tbslas_ctx *ctx, const TensorParams &p, const TensorTables &tt, size_t n_leaf) {
  constexpr int DP = D | 1, DJ = (D + 1) & ~1;
  const size_t smem = sizeof(double) * ((size_t)D * DP + (D * DP & 1) + 2 * D * DJ + p.vstride + (size_t)D * D * D);
  auto k = tensor_grid_eval_kernel_t<D>;
  TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  const size_t want = n_leaf < (size_t)ctx->n_sm * 16 ? n_leaf : (size_t)ctx->n_sm * 16;
  k<<<(unsigned)want, kTensorThreads, smem, ctx->stream>>>(p, tt);
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}


// ---------------------------------------------------------------------------------------------
// The same three passes as FP64 tensor-core GEMMs (mma.sync m8n8k4, SASS DMMA.8x8x4): unlike the
// point-by-point contraction (a triangular [points x 120] x [120 x 45] product whose
// padding eats the tensor path's advantage) these ARE small dense matrix products per leaf,
//   pass 1   A1[(i,j)][px]  = Cpad[(i,j)][k]  x Tx[k][px]     [n_row x D] x [D x D]   (rows zero padded;
//                                                               k-steps beyond a tile's longest row skipped)
//   pass 2   B2[i][py][px]  = Ty^T[py][j]     x A1_i[j][px]    [D x (D-i)] x [(D-i) x D]  per plane i
//   pass 3   U[pz][(py,px)] = Tz^T[pz][i]     x B2[i][(py,px)] [D x D] x [D x D^2]
// with D padded to 16.  The scalar kernel above reads one shared-memory operand per FMA and is bound by
// the LDS pipe (FP64 pipe 23 %); a DMMA does 256 FMAs for two 8-byte operands per lane (one in pass 3,
// where the Tz fragments stay in registers).  Fragments (PTX ISA, m8n8k4 .row.col f64): lane l holds
// A[l/4][l%4], B[l%4][l/4], C[l/4][2(l%4) .. +1].  8 warps per CTA share the tiles of a pass; one
// component at a time: 70 KB of shared memory per CTA, 3 CTAs per SM.  Basis tables, exception masks,
// the arrival-point generation and the epilogue are the scalar kernel's.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Fragment loads address (row = 4*step + l%4, col = base + l/4): with 16-double rows the four rows of
// a fragment fall on the same banks (a 4-way conflict on every load).  XOR-ing the column with
// 4*(row mod 4) spreads them over the banks without padding.
__device__ __forceinline__ int swz(int row, int col) { return row * 16 + (col ^ ((row & 3) << 2)); }

// ---------------------------------------------------------------------------------------------
// Second DMMA kernel (round 2; runs q = 15, and q <= 14 as an A/B switch): the three GEMM passes with the
// per-leaf work AROUND the GEMMs taken off the critical path.  ncu of the first version (round 2's first
// session; removed) showed 4 % FP64
// and 46 % issue utilisation with 82 k SM-clocks per three leaves: the time went into (a) the arrival-point
// loop (four integer divisions and a lane-divergent constant-bank load per element), (b) lane-divergent
// constant-bank loads of the row tables and nodes inside the passes (an LDC with different indices in a
// warp is replayed per distinct index), (c) a synchronous 16 KB coefficient copy per leaf, (d) k-step loops
// of data-dependent length whose loads were issued one DMMA at a time.  Here:
//   * tables (nodes, row offsets/lengths) live in shared memory, filled once per CTA;
//   * the coordinates of the leaf's grid are 3 x D numbers (sX) computed once per leaf by the threads that
//     build the bases; the arrival points are a table expansion with per-thread selectors fixed for the
//     whole launch (no division per element), and the epilogue takes its base coordinate from sX too
//     (same expression, same bits);
//   * the NEXT leaf's coefficient block arrives by one bulk-TMA copy (cp.async.bulk + mbarrier), issued
//     as soon as the last component's pass 1 has consumed the current block: no copy loop, no latency;
//   * the k-steps of passes 1 and 2 are unrolled: all fragment loads first, then the DMMAs;
//   * persistent CTAs (3 per SM), one barrier less per leaf (exception counters double buffered).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_cta() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int D>
__global__ void __launch_bounds__(kTensorThreads, 3)
tensor_grid_dmma2_kernel(const TensorParams p, const TensorTables tb_) {
  constexpr int P2 = D * D, P = P2 * D, NROW = D * (D + 1) / 2, MT1 = (NROW + 7) / 8, KS = (D + 3) / 4;
  constexpr int XG = (3 * P2 + kTensorThreads - 1) / kTensorThreads;  // arrival-point slots per thread and z-plane
  static_assert(D <= 16, "one 16-wide tile per axis");
  extern __shared__ __align__(16) double sm[];
  double *sT = sm;                    // [3][16][16]  T_deg(point) per axis, degree major, zero padded
  double *sC3 = sT + 3 * 256;         // [3][ncoef_pad]  (bulk-TMA destination, 16-byte aligned)
  double *sA1 = sC3 + p.vstride;      // [MT1*8][16]
  double *sB2 = sA1 + MT1 * 8 * 16;   // [16][16][16]  (plane, py, px)
  __shared__ double sX[3 * 16];       // coordinates of the leaf's grid, per axis
  __shared__ double sNode[16];
  __shared__ uint32_t sRow[MT1 * 8];  // row (i,j): first coefficient | length << 16
  __shared__ unsigned s_ok[2][4];     // [leaf parity][axis]
  __shared__ unsigned s_exc[2][2];    // [leaf parity]{base, n}
  __shared__ __align__(8) uint64_t s_bar;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int lr = lane >> 2, lc = lane & 3;  // fragment row / column index of this lane
  for (int e = D * 256 + t; e < 16 * 256; e += kTensorThreads) sB2[e] = 0.0;  // planes >= D stay zero (pass 2 writes whole planes < D)
  for (int e = t; e < 3 * 256; e += kTensorThreads) sT[e] = 0.0;    // degrees / points >= D stay zero
  if (t < 16) sNode[t] = t < D ? tb_.node[t] : 0.0;
  if (t < MT1 * 8) sRow[t] = t < NROW ? ((uint32_t)tb_.row_off[t] | ((uint32_t)tb_.row_len[t] << 16)) : 0u;
  if (t < 8) (&s_ok[0][0])[t] = 0u;
  if (t < 4) (&s_exc[0][0])[t] = 0u;
  if (t == 0) mbar_init(&s_bar, 1);
  // selectors of this thread's arrival-point slots inside one z-plane: element r = 3*(py*D + px) + axis
  int xsel[XG];  // >= 0: index into sX (x: px, y: 16 + py), -1: the plane's z, -2: no element
#pragma unroll
  for (int m = 0; m < XG; m++) {
    const int r = t + m * kTensorThreads;
    const int pt = r / 3, a = r - 3 * pt, py = pt / D, px = pt - py * D;
    xsel[m] = r >= 3 * P2 ? -2 : (a == 0 ? px : (a == 1 ? 16 + py : -1));
  }
  __syncthreads();
  auto fetch_coeff = [&](int jv) {  // one thread: the block of velocity leaf jv -> sC3
    fence_proxy_async_cta();        // earlier generic-proxy reads of sC3 vs. the async-proxy write
    const uint32_t bytes = p.vstride * 8u;
    mbar_expect_tx(&s_bar, bytes);
    tma_bulk_g2s(sC3, vel_coeff(p, jv), bytes, &s_bar);
  };
  uint32_t phase = 0;
  bool prefetched = false;  // CTA-uniform: this leaf's block was requested during the previous leaf
  int par = 0;
  for (size_t leaf = blockIdx.x; leaf < p.n_leaf; leaf += gridDim.x, par ^= 1) {
    const int j = p.map[leaf];
    const size_t gp0 = leaf * (size_t)P;
    const double4 gc = p.ggeom[leaf];
    const double glen = 1.0 / (double)(1u << p.gdepth[leaf]);
    __syncthreads();  // the previous leaf's passes are done with sT, sX, sC3
    if (j >= 0 && !prefetched && t == 0) fetch_coeff(j);
    if (t < 3 * D) {
      const int a = t / D, i = t - a * D;
      const double c = a == 0 ? gc.x : (a == 1 ? gc.y : gc.z);
      const double x = __dadd_rn(c, __dmul_rn(glen, sNode[i]));  // gridpts.cu
      sX[a * 16 + i] = x;
      if (j >= 0) {
        const double4 gv = vel_geom(p, j);
        const uint4 vb = vel_box(p, j);
        const double vc = a == 0 ? gv.x : (a == 1 ? gv.y : gv.z);
        const unsigned vba = a == 0 ? vb.x : (a == 1 ? vb.y : vb.z);
        const double xi = __dadd_rn(__dmul_rn(__dsub_rn(x, vc), gv.w), -1.0);  // cheb_eval.cuh
        const bool in = fabs(xi) <= 1.0;
        const double xc = in ? xi : 0.0, x2 = 2.0 * xc;
        double t0 = in ? 1.0 : 0.0, t1 = xc;
        double *T = sT + a * 256;  // [degree][point], swizzled
        T[swz(0, i)] = t0;
        if (D > 1) T[swz(1, i)] = t1;
#pragma unroll
        for (int k = 2; k < D; k++) {
          const double t2 = __dsub_rn(__dmul_rn(x2, t1), t0);
          T[swz(k, i)] = t2;
          t0 = t1;
          t1 = t2;
        }
        const double xs = x * 32768.0;
        int jx = __double2int_rd(xs);
        if (!p.periodic && xs == 32768.0) jx = 32767;
        if ((unsigned)jx < 32768u && ((((unsigned)jx ^ vba) >> vb.w) == 0u)) atomicOr(&s_ok[par][a], 1u << i);
      }
    }
    __syncthreads();
    const unsigned okx = s_ok[par][0], oky = s_ok[par][1], okz = s_ok[par][2];
    const unsigned n_reg = (unsigned)(__popc(okx) * __popc(oky) * __popc(okz));
    if (t == 0 && n_reg < (unsigned)P) s_exc[par][0] = atomicAdd(p.exc_count, (unsigned)P - n_reg);
    if (t < 3) s_ok[par ^ 1][t] = 0u;  // the next leaf's masks and counter (last read before this leaf's first barrier)
    if (t == 0) s_exc[par ^ 1][1] = 0u;
    if (p.xgen) {  // arrival points of this leaf: a table expansion, fully coalesced
      double xv[XG];
#pragma unroll
      for (int m = 0; m < XG; m++) xv[m] = xsel[m] >= 0 ? sX[xsel[m]] : 0.0;
      double *o = p.xgen + 3 * gp0 + t;
#pragma unroll 5
      for (int pz = 0; pz < D; pz++) {
        const double zq = sX[32 + pz];
#pragma unroll
        for (int m = 0; m < XG; m++)
          if (xsel[m] != -2) TB_TENSOR_XSTORE(o + pz * 3 * P2 + m * kTensorThreads, xsel[m] == -1 ? zq : xv[m]);
      }
    }
    if (j >= 0) {  // the coefficient block has landed (waited for even when no pass runs: phases stay in step)
      mbar_wait(&s_bar, phase);
      phase ^= 1u;
    }
    prefetched = false;
    if (j >= 0 && n_reg) {
      // Tz fragments of pass 3 (the same for every column tile and component)
      double az[2][KS];
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int ks = 0; ks < KS; ks++) az[mt][ks] = sT[512 + swz(ks * 4 + lc, mt * 8 + lr)];
      for (int l = 0; l < 3; l++) {
        const double *C = sC3 + l * p.ncoef_pad;
        // ---- pass 1: A1[r][px] = sum_k Cpad[r][k] Tx[k][px]
        // (a warp keeps its column tile: the Tx fragments are loaded once per pass; every fragment address
        //  is base + 64 * k-step, the swizzle term does not depend on the k-step)
        {
          const int nt = warp & 1;
          const double *pb = sT + lc * 16 + ((nt * 8 + lr) ^ (lc << 2));
          double b[KS];
#pragma unroll
          for (int ks = 0; ks < KS; ks++) b[ks] = pb[ks * 64];
          for (int mt = warp >> 1; mt < MT1; mt += kTensorThreads / 64) {
            const int r = mt * 8 + lr;
            const uint32_t rw = sRow[r];
            const int rlen = (int)(rw >> 16);
            const double *pa = C + (rw & 0xffffu) + lc;
            const int nks = (tb_.tile_k[mt] + 3) >> 2;  // warp-uniform index
            double a[KS];
#pragma unroll
            for (int ks = 0; ks < KS; ks++) a[ks] = ks * 4 + lc < rlen ? pa[ks * 4] : 0.0;
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
              if (ks < nks) dmma884(c0, c1, a[ks], b[ks]);
            *reinterpret_cast<double2 *>(sA1 + swz(r, nt * 8 + 2 * lc)) = make_double2(c0, c1);
          }
        }
        __syncthreads();
        if (l == 2) {  // sC3 is free: request the next leaf's block now, it lands under passes 2 and 3
          const size_t next = leaf + gridDim.x;
          if (next < p.n_leaf) {
            const int jn = p.map[next];
            if (jn >= 0 && t == 0) fetch_coeff(jn);
            prefetched = true;
          }
        }
        // ---- pass 2: B2[i][py][px] = sum_j Ty[j][py] A1[(i,j)][px]
        // (a warp keeps its output tile (py tile, px tile) and takes every other plane: the Ty fragments
        //  are loaded once per pass)
        {
          const int mt = warp & 1, nt = (warp >> 1) & 1;
          const double *pa = sT + 256 + lc * 16 + ((mt * 8 + lr) ^ (lc << 2));
          double a[KS];
#pragma unroll
          for (int ks = 0; ks < KS; ks++) a[ks] = pa[ks * 64];
          for (int i = warp >> 2; i < D; i += kTensorThreads / 128) {
            const int nj = D - i, rb = tb_.row_first[i] + lc;  // warp-uniform index
            const double *pb = sA1 + rb * 16 + ((nt * 8 + lr) ^ ((rb & 3) << 2));
            double b[KS];
#pragma unroll
            for (int ks = 0; ks < KS; ks++) b[ks] = ks * 4 + lc < nj ? pb[ks * 64] : 0.0;
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
              if (ks * 4 < nj) dmma884(c0, c1, a[ks], b[ks]);
            *reinterpret_cast<double2 *>(sB2 + i * 256 + (mt * 8 + lr) * 16 + ((nt * 8 + 2 * lc) ^ ((i & 3) << 2))) = make_double2(c0, c1);
          }
        }
        __syncthreads();
        // ---- pass 3: U[pz][(py,px)] = sum_i Tz[i][pz] B2[i][(py,px)];  x' = x + alpha U on regular points
        const double *sXl = sX + l * 16;
        double *const ob = p.out + 3 * gp0 + l;
        const unsigned okxd = okx & ((1u << D) - 1u), okzd = okz & ((1u << D) - 1u);  // (bits >= D are never set)
        for (int nt = warp; nt < 2 * D; nt += kTensorThreads / 32) {
          const int py = nt >> 1, px0 = (nt & 1) * 8 + 2 * lc;
          if (!((oky >> py) & 1u)) continue;  // warp-uniform: the whole row is exceptions
          double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
          const double *pb = sB2 + lc * 256 + ((nt * 8 + lr) ^ (lc << 2));
#pragma unroll
          for (int ks = 0; ks < KS; ks++) {
            const double b = pb[ks * 1024];
            dmma884(c[0][0], c[0][1], az[0][ks], b);
            dmma884(c[1][0], c[1][1], az[1][ks], b);
          }
          double *o = ob + 3 * (lr * P2 + py * D + px0);
#pragma unroll
          for (int mt = 0; mt < 2; mt++) {
            const int pz = mt * 8 + lr;
            if (!((okzd >> pz) & 1u)) continue;
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int px = px0 + h;
              if (!((okxd >> px) & 1u)) continue;
              const double x0 = sXl[l == 0 ? px : (l == 1 ? py : pz)];
              o[mt * (24 * P2) + 3 * h] = __dadd_rn(x0, __dmul_rn(p.alpha, c[mt][h]));  // traj.inc:36
            }
          }
        }
        // (pass 1 of the next component only writes sA1; its barrier orders pass 2 after this pass 3)
      }
    }
    if (n_reg < (unsigned)P) {  // list the exceptions of this leaf (CTA-uniform condition)
      __syncthreads();
      const unsigned base = s_exc[par][0];
      for (int e = t; e < P; e += kTensorThreads) {
        const int pz = e / P2, rem = e - pz * P2, py = rem / D, px = rem - py * D;
        if (j >= 0 && (((okx >> px) & (oky >> py) & (okz >> pz)) & 1u)) continue;
        const unsigned k = atomicAdd(&s_exc[par][1], 1u);
        p.exc_idx[base + k] = (uint32_t)(gp0 + e);
      }
    }
  }
}

template <int D>
static int launch_tensor_dmma2(tbslas_ctx *ctx, const TensorParams &p, const TensorTables &tt, size_t n_leaf) {
  constexpr int NROW = D * (D + 1) / 2, MT1 = (NROW + 7) / 8;
  const size_t smem = sizeof(double) * ((size_t)3 * 256 + p.vstride + (size_t)MT1 * 8 * 16 + 16 * 256);
  auto k = tensor_grid_dmma2_kernel<D>;
  TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  const size_t cap = (size_t)ctx->n_sm * (size_t)(ctx->opt.tensor_ctas_per_sm > 0 ? ctx->opt.tensor_ctas_per_sm : 600);  // one CTA per leaf measured best
  const size_t want = n_leaf < cap ? n_leaf : cap;
  k<<<(unsigned)want, kTensorThreads, smem, ctx->stream>>>(p, tt);
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

// ---------------------------------------------------------------------------------------------
// Third DMMA kernel (q <= 14): the same passes, but pass 3 runs ONCE for the three velocity components, with
// the component as the fastest index of its column dimension,
//     U[pz][(py, px, l)] = sum_i Tz[i][pz] B[i][(py, px, l)],
// so that a warp's 8 x 8 output tile is, for every pz, 64 contiguous bytes of the AoS position array and the
// three components of a point are written together -- kernel 2 fills every 32-byte sector 8 bytes at a time in
// three passes microseconds apart (18.3 GB written + 5.0 GB of read-modify-write reads for 13.2 GB).  The price
// is the B array of all three components resident in shared memory ([D][3 D^2] doubles, 81.6 KB at q = 14), i.e.
// two CTAs per SM instead of three; the coefficient block is staged one component at a time to make room.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kTensorThreads, 2)
tensor_grid_dmma3_kernel(const TensorParams p, const TensorTables tb_) {
  constexpr int P2 = D * D, P = P2 * D, NROW = D * (D + 1) / 2, MT1 = (NROW + 7) / 8, KS = (D + 3) / 4;
  constexpr int NC = 3 * P2, NB = (NC + 7) & ~7;  // columns of pass 3 (py, px, l), padded to whole tiles
  constexpr int XG = (3 * P2 + kTensorThreads - 1) / kTensorThreads;
  static_assert(D <= 15, "the B array of three components must fit two CTAs per SM");
  extern __shared__ __align__(16) double sm[];
  double *sT = sm;                    // [3][16][16]  T_deg(point) per axis, degree major, zero padded
  double *sC = sT + 3 * 256;          // [ncoef_pad]  one component at a time (bulk-TMA destination)
  double *sA1 = sC + p.ncoef_pad;     // [MT1*8][16]
  double *sB = sA1 + MT1 * 8 * 16;    // [D][NB]
  __shared__ double sX[3 * 16];
  __shared__ double sNode[16];
  __shared__ uint32_t sRow[MT1 * 8];  // row (i,j): first coefficient | length << 16
  __shared__ uint32_t sCol[NB];       // column (py, px, l): px | py << 8 | index of its base coordinate in sX << 16 (0xff: z)
  __shared__ unsigned s_ok[2][4];
  __shared__ unsigned s_exc[2][2];
  __shared__ __align__(8) uint64_t s_bar;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  for (int e = t; e < 3 * 256; e += kTensorThreads) sT[e] = 0.0;
  for (int e = NC + t; e < D * NB; e += kTensorThreads)  // the padding columns of every row stay zero
    if (e % NB >= NC) sB[e] = 0.0;
  if (t < 16) sNode[t] = t < D ? tb_.node[t] : 0.0;
  if (t < MT1 * 8) sRow[t] = t < NROW ? ((uint32_t)tb_.row_off[t] | ((uint32_t)tb_.row_len[t] << 16)) : 0u;
  for (int c = t; c < NB; c += kTensorThreads) {
    const int pt = c / 3, l = c - 3 * pt, py = pt / D, px = pt - py * D;
    sCol[c] = c < NC ? ((uint32_t)px | ((uint32_t)py << 8) | ((uint32_t)(l == 0 ? px : (l == 1 ? 16 + py : 0xff)) << 16))
                     : 0xffffffffu;
  }
  if (t < 8) (&s_ok[0][0])[t] = 0u;
  if (t < 4) (&s_exc[0][0])[t] = 0u;
  if (t == 0) mbar_init(&s_bar, 1);
  int xsel[XG];
#pragma unroll
  for (int m = 0; m < XG; m++) {
    const int r = t + m * kTensorThreads;
    const int pt = r / 3, a = r - 3 * pt, py = pt / D, px = pt - py * D;
    xsel[m] = r >= 3 * P2 ? -2 : (a == 0 ? px : (a == 1 ? 16 + py : -1));
  }
  __syncthreads();
  uint32_t phase = 0;
  bool prefetched = false;  // CTA-uniform: component 0 of this leaf's block was requested during the previous leaf
  int par = 0;
  for (size_t leaf = blockIdx.x; leaf < p.n_leaf; leaf += gridDim.x, par ^= 1) {
    const int j = p.map[leaf];
    const size_t gp0 = leaf * (size_t)P;
    const double4 gc = p.ggeom[leaf];
    const double glen = 1.0 / (double)(1u << p.gdepth[leaf]);
    __syncthreads();  // the previous leaf's passes are done with sT, sX, sB
    if (t < 3 * D) {
      const int a = t / D, i = t - a * D;
      const double c = a == 0 ? gc.x : (a == 1 ? gc.y : gc.z);
      const double x = __dadd_rn(c, __dmul_rn(glen, sNode[i]));  // gridpts.cu
      sX[a * 16 + i] = x;
      if (j >= 0) {
        const double4 gv = vel_geom(p, j);
        const uint4 vb = vel_box(p, j);
        const double vc = a == 0 ? gv.x : (a == 1 ? gv.y : gv.z);
        const unsigned vba = a == 0 ? vb.x : (a == 1 ? vb.y : vb.z);
        const double xi = __dadd_rn(__dmul_rn(__dsub_rn(x, vc), gv.w), -1.0);  // cheb_eval.cuh
        const bool in = fabs(xi) <= 1.0;
        const double xc = in ? xi : 0.0, x2 = 2.0 * xc;
        double t0 = in ? 1.0 : 0.0, t1 = xc;
        double *T = sT + a * 256;
        T[swz(0, i)] = t0;
        if (D > 1) T[swz(1, i)] = t1;
#pragma unroll
        for (int k = 2; k < D; k++) {
          const double t2 = __dsub_rn(__dmul_rn(x2, t1), t0);
          T[swz(k, i)] = t2;
          t0 = t1;
          t1 = t2;
        }
        const double xs = x * 32768.0;
        int jx = __double2int_rd(xs);
        if (!p.periodic && xs == 32768.0) jx = 32767;
        if ((unsigned)jx < 32768u && ((((unsigned)jx ^ vba) >> vb.w) == 0u)) atomicOr(&s_ok[par][a], 1u << i);
      }
    }
    __syncthreads();
    const unsigned okx = s_ok[par][0], oky = s_ok[par][1], okz = s_ok[par][2];
    const unsigned n_reg = (unsigned)(__popc(okx) * __popc(oky) * __popc(okz));
    const bool passes = j >= 0 && n_reg;  // CTA-uniform
    auto fetch_comp = [&](int jv, int l) {  // one thread: component l of velocity leaf jv's block -> sC
      fence_proxy_async_cta();
      const uint32_t bytes = p.ncoef_pad * 8u;
      mbar_expect_tx(&s_bar, bytes);
      tma_bulk_g2s(sC, vel_coeff(p, jv) + (size_t)l * p.ncoef_pad, bytes, &s_bar);
    };
    if (passes && !prefetched && t == 0) fetch_comp(j, 0);
    if (!passes && prefetched) {  // requested for nothing (every point of the leaf is an exception): keep the phases in step
      mbar_wait(&s_bar, phase);
      phase ^= 1u;
    }
    prefetched = false;
    if (t == 0 && n_reg < (unsigned)P) s_exc[par][0] = atomicAdd(p.exc_count, (unsigned)P - n_reg);
    if (t < 3) s_ok[par ^ 1][t] = 0u;
    if (t == 0) s_exc[par ^ 1][1] = 0u;
    if (p.xgen) {  // arrival points of this leaf: a table expansion, fully coalesced
      double xv[XG];
#pragma unroll
      for (int m = 0; m < XG; m++) xv[m] = xsel[m] >= 0 ? sX[xsel[m]] : 0.0;
      double *o = p.xgen + 3 * gp0 + t;
#pragma unroll 5
      for (int pz = 0; pz < D; pz++) {
        const double zq = sX[32 + pz];
#pragma unroll
        for (int m = 0; m < XG; m++)
          if (xsel[m] != -2) TB_TENSOR_XSTORE(o + pz * 3 * P2 + m * kTensorThreads, xsel[m] == -1 ? zq : xv[m]);
      }
    }
    if (passes) {
      for (int l = 0; l < 3; l++) {
        mbar_wait(&s_bar, phase);  // component l has landed
        phase ^= 1u;
        // ---- pass 1: A1[r][px] = sum_k Cpad[r][k] Tx[k][px]
        {
          const int nt = warp & 1;
          const double *pb = sT + lc * 16 + ((nt * 8 + lr) ^ (lc << 2));
          double b[KS];
#pragma unroll
          for (int ks = 0; ks < KS; ks++) b[ks] = pb[ks * 64];
          for (int mt = warp >> 1; mt < MT1; mt += kTensorThreads / 64) {
            const int r = mt * 8 + lr;
            const uint32_t rw = sRow[r];
            const int rlen = (int)(rw >> 16);
            const double *pa = sC + (rw & 0xffffu) + lc;
            const int nks = (tb_.tile_k[mt] + 3) >> 2;
            double a[KS];
#pragma unroll
            for (int ks = 0; ks < KS; ks++) a[ks] = ks * 4 + lc < rlen ? pa[ks * 4] : 0.0;
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
              if (ks < nks) dmma884(c0, c1, a[ks], b[ks]);
            *reinterpret_cast<double2 *>(sA1 + swz(r, nt * 8 + 2 * lc)) = make_double2(c0, c1);
          }
        }
        __syncthreads();
        if (l < 2) {  // sC is free: the next component lands under pass 2 ...
          if (t == 0) fetch_comp(j, l + 1);
        } else {      // ... and after the last one, component 0 of the next leaf under passes 2 and 3
          const size_t next = leaf + gridDim.x;
          if (next < p.n_leaf) {
            const int jn = p.map[next];
            if (jn >= 0) {
              if (t == 0) fetch_comp(jn, 0);
              prefetched = true;
            }
          }
        }
        // ---- pass 2: B[i][(py, px, l)] = sum_j Ty[j][py] A1[(i,j)][px]
        {
          const int mt = warp & 1, nt = (warp >> 1) & 1;
          const double *pa = sT + 256 + lc * 16 + ((mt * 8 + lr) ^ (lc << 2));
          double a[KS];
#pragma unroll
          for (int ks = 0; ks < KS; ks++) a[ks] = pa[ks * 64];
          const int py = mt * 8 + lr, px = nt * 8 + 2 * lc;
          for (int i = warp >> 2; i < D; i += kTensorThreads / 128) {
            const int nj = D - i, rb = tb_.row_first[i] + lc;
            const double *pbb = sA1 + rb * 16 + ((nt * 8 + lr) ^ ((rb & 3) << 2));
            double b[KS];
#pragma unroll
            for (int ks = 0; ks < KS; ks++) b[ks] = ks * 4 + lc < nj ? pbb[ks * 64] : 0.0;
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
              if (ks * 4 < nj) dmma884(c0, c1, a[ks], b[ks]);
            if (py < D) {
              double *o = sB + i * NB + 3 * (py * D + px) + l;
              if (px < D) o[0] = c0;
              if (px + 1 < D) o[3] = c1;
            }
          }
        }
        __syncthreads();
      }
      // ---- pass 3, all components: U[pz][c] = sum_i Tz[i][pz] B[i][c];  x' = x + alpha U on regular points
      double az[2][KS];
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int ks = 0; ks < KS; ks++) az[mt][ks] = sT[512 + swz(ks * 4 + lc, mt * 8 + lr)];
      double *const ob = p.out + 3 * gp0;
      for (int nt = warp; nt < NB / 8; nt += kTensorThreads / 32) {
        double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        const double *pb = sB + lc * NB + nt * 8 + lr;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
          const double b = ks * 4 + lc < D ? pb[ks * 4 * NB] : 0.0;
          dmma884(c[0][0], c[0][1], az[0][ks], b);
          dmma884(c[1][0], c[1][1], az[1][ks], b);
        }
        const int col0 = nt * 8 + 2 * lc;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const uint32_t ci = sCol[col0 + h];
          if (ci == 0xffffffffu) continue;
          const int px = (int)(ci & 0xffu), py = (int)((ci >> 8) & 0xffu), sel = (int)(ci >> 16);
          if (!((okx >> px) & (oky >> py) & 1u)) continue;
#pragma unroll
          for (int mt = 0; mt < 2; mt++) {
            const int pz = mt * 8 + lr;
            if (pz >= D || !((okz >> pz) & 1u)) continue;
            const double x0 = sX[sel == 0xff ? 32 + pz : sel];
            ob[pz * NC + col0 + h] = __dadd_rn(x0, __dmul_rn(p.alpha, c[mt][h]));  // traj.inc:36
          }
        }
      }
    }
    if (n_reg < (unsigned)P) {  // list the exceptions of this leaf (CTA-uniform condition)
      __syncthreads();
      const unsigned base = s_exc[par][0];
      for (int e = t; e < P; e += kTensorThreads) {
        const int pz = e / P2, rem = e - pz * P2, py = rem / D, px = rem - py * D;
        if (j >= 0 && (((okx >> px) & (oky >> py) & (okz >> pz)) & 1u)) continue;
        const unsigned k = atomicAdd(&s_exc[par][1], 1u);
        p.exc_idx[base + k] = (uint32_t)(gp0 + e);
      }
    }
  }
}

template <int D>
static int launch_tensor_dmma3(tbslas_ctx *ctx, const TensorParams &p, const TensorTables &tt, size_t n_leaf) {
  if constexpr (D > 15) {
    return launch_tensor_dmma2<D>(ctx, p, tt, n_leaf);
  } else {
    constexpr int NROW = D * (D + 1) / 2, MT1 = (NROW + 7) / 8, NB = (3 * D * D + 7) & ~7;
    const size_t smem = sizeof(double) * ((size_t)3 * 256 + p.ncoef_pad + (size_t)MT1 * 8 * 16 + (size_t)D * NB);
    auto k = tensor_grid_dmma3_kernel<D>;
    TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    // 2 resident CTAs per SM; flat between 2 and 16 CTAs per SM (4.80-4.83 ms on C2), 5.2 at 32, 5.4 per leaf
    const int per_sm = ctx->opt.tensor_ctas_per_sm > 0 ? ctx->opt.tensor_ctas_per_sm : 8;
    const size_t cap = (size_t)ctx->n_sm * (size_t)per_sm;
    const size_t want = n_leaf < cap ? n_leaf : cap;
    k<<<(unsigned)want, kTensorThreads, smem, ctx->stream>>>(p, tt);
    TB_CUDA(ctx, cudaGetLastError());
    return TBSLAS_OK;
  }
}

__global__ void publish_count_kernel(const unsigned *__restrict__ src, unsigned *host_word) {
  *reinterpret_cast<volatile unsigned *>(host_word) = *src;
  __threadfence_system();
}

// exceptions: positions out, values back in
__global__ void gather_points_kernel(const double *__restrict__ x, const uint32_t *__restrict__ idx, size_t m,
                                     double *__restrict__ out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * m) return;
  const size_t s = e / 3;
  out[e] = x[3 * (size_t)idx[s] + (e - 3 * s)];
}
// the same for arrival points that were never written (virtual x): rebuilt from the grid
__global__ void gather_grid_points_kernel(const GridBase gb, const uint32_t *__restrict__ idx, size_t m,
                                          double *__restrict__ out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * m) return;
  const size_t s = e / 3;
  const unsigned i = idx[s];
  GridBase raw = gb;
  raw.periodic = 0;  // the evaluation that follows wraps, exactly as it would wrap a stored point
  out[e] = grid_base_coord(raw, gb.ggeom[i / gb.P], i, (int)(e - 3 * s));
}
__global__ void scatter_update_kernel(const double *__restrict__ pos, const double *__restrict__ val,
                                      const uint32_t *__restrict__ idx, size_t m, double alpha, int periodic,
                                      double *__restrict__ x, double *__restrict__ out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * m) return;
  const size_t s = e / 3, o = 3 * (size_t)idx[s] + (e - 3 * s);
  const double b = pos[e];  // wrapped in place by the evaluation when periodic (tree_functor.h:442-449)
  if (periodic && x) x[o] = b;
  out[o] = __dadd_rn(b, __dmul_rn(alpha, val[e]));
}

// Stage 1 of the first RK2 sub-step on the grid points of leaves [leaf0, leaf0 + n_leaf) of `grid`:
// out = x + alpha * vel(x).  `x` is rewritten only where the periodic wrap changes it -- or, with
// gen_points, WRITTEN: the kernel then also produces the arrival points (gridpts.cu's job).
// Returns TBSLAS_ERR_UNSUPPORTED (without touching anything) when the shortcut does not apply; in a
// multi-rank context that decision only uses what every rank knows, because the generic pass over the
// exceptions of a Morton-sharded velocity tree is a collective evaluation.
int eval_tree_dev_points(tbslas_tree *t, int bc, double *pos, size_t n, double *out);  // api.cu
int comm_allgather_bytes(tbslas_ctx *ctx, const void *mine, void *all, size_t bytes);       // comm.cu

void make_grid_base(const tbslas_tree *grid, size_t leaf0, int bc, GridBase *gb) {
  const int d = grid->q + 1;
  gb->ggeom = grid->d_geom + leaf0;
  gb->D = (unsigned)d;
  gb->P = (unsigned)(d * d * d);
  gb->periodic = bc == TBSLAS_PERIODIC;
  new_nodes_host(grid->q, gb->node);
}

// virtual x: only the kernels that rebuild the grid points themselves (compile-time degrees)
bool tensor_grid_supports_virtual_x(const tbslas_tree *vel) { return vel->q + 1 <= 15; }

int launch_tensor_grid_eval(tbslas_ctx *ctx, tbslas_tree *vel, const tbslas_tree *grid, size_t leaf0,
                            size_t n_leaf, int bc, double *x, double *out, double alpha, bool gen_points,
                            bool virtual_x) {
  const bool collective = ctx->nranks > 1 && !vel->replicated;
  if (vel->dof != 3 || vel->q != grid->q || !vel->boxes_all || !grid->boxes_all || (!collective && !vel->n_leaf))
    return TBSLAS_ERR_UNSUPPORTED;
  const int d = vel->q + 1, dp = d | 1, P = d * d * d, n_row = d * (d + 1) / 2;
  const size_t n = n_leaf * (size_t)P;
  if (!n && !collective) return TBSLAS_OK;
  if (n >= (size_t)0xfffffff0u) return TBSLAS_ERR_UNSUPPORTED;
  TB_TRY(tree_coeff_ready(vel));
  const int periodic = (bc == TBSLAS_PERIODIC);
  void *exc_idx = nullptr, *misc = nullptr;
  unsigned *exc_count = nullptr;
  // Morton-sharded velocity tree: every rank's last leaf, all-gathered (see TensorParams)
  const char *ghost = nullptr;
  const size_t ghost_rec = kGhostHdr + sizeof(double) * vel->stride;
  if (collective) {
    void *g;
    TB_TRY(ws_get(ctx, WS_GHOST, ghost_rec * (ctx->nranks + 1), &g));
    char *mine = (char *)g + ghost_rec * ctx->nranks;
    StageScope sc(ctx, ST_EXCHANGE, (double)(ghost_rec * ctx->nranks), 1);
    ghost_pack_kernel<<<1, 256, 0, ctx->stream>>>(vel->d_box, vel->d_geom, vel->d_coeff, (int)vel->n_leaf,
                                                 (unsigned)vel->stride, mine);
    TB_CUDA(ctx, cudaGetLastError());
    TB_TRY(comm_allgather_bytes(ctx, mine, g, ghost_rec));
    ghost = (const char *)g;
  }
  if (n) {
    TensorTables tt;
    new_nodes_host(vel->q, tt.node);
    int off = 0, r = 0;
    memset(tt.tile_k, 0, sizeof(tt.tile_k));
    for (int i = 0; i < d; i++) {
      tt.row_first[i] = (uint16_t)r;
      for (int j = 0; i + j < d; j++) {
        tt.row_len[r] = (uint8_t)(d - i - j);
        if (tt.row_len[r] > tt.tile_k[r >> 3]) tt.tile_k[r >> 3] = tt.row_len[r];
        tt.row_off[r++] = (uint16_t)off;
        off += d - i - j;
      }
    }
    tt.row_first[d] = (uint16_t)r;
    void *map;
    TB_TRY(ws_get(ctx, WS_GRIDMAP, sizeof(int32_t) * n_leaf, &map));
    TB_TRY(ws_get(ctx, WS_EXC_IDX, sizeof(uint32_t) * (n + 1), &exc_idx));
    TB_TRY(ws_get(ctx, WS_MISC, 64, &misc));
    exc_count = (unsigned *)misc;
    StageScope sc(ctx, ST_TENSOR, (double)n, 2);
    TB_CUDA(ctx, cudaMemsetAsync(exc_count, 0, sizeof(unsigned), ctx->stream));
    grid_leaf_map_kernel<<<(unsigned)((n_leaf + 255) / 256), 256, 0, ctx->stream>>>(
        grid->d_box + leaf0, n_leaf, vel->d_key, vel->d_box, vel->d_cell, vel->cell_shift, (int)vel->n_leaf,
        (int32_t *)map, ghost, ghost_rec, ctx->nranks, ctx->rank);
    TB_CUDA(ctx, cudaGetLastError());
    TensorParams p;
    p.vcoeff = vel->d_coeff;
    p.vgeom = vel->d_geom;
    p.vbox = vel->d_box;
    p.v_nleaf = (int)vel->n_leaf;
    p.ghost = ghost;
    p.ghost_rec = ghost_rec;
    p.vstride = (unsigned)vel->stride;
    p.ncoef_pad = (unsigned)(vel->stride / vel->dof);
    p.ggeom = grid->d_geom + leaf0;
    p.gdepth = grid->d_depth + leaf0;
    p.map = (const int32_t *)map;
    p.n_leaf = n_leaf;
    p.d = d;
    p.periodic = periodic;
    p.x = x;
    p.xgen = (gen_points && !virtual_x) ? x : nullptr;  // virtual x: the arrival points are never written
    p.out = out;
    p.alpha = alpha;
    p.exc_count = exc_count;
    p.exc_idx = (uint32_t *)exc_idx;
    const bool force_generic = ctx->opt.tensor_generic;
    const int dmma_mode = ctx->opt.tensor_dmma;  // 0: scalar kernels, 2: the second DMMA kernel, 3: the default (2 for q = 15)
    bool launched = false;
    if (!force_generic && dmma_mode) {
      switch (d) {
#define TB_CASE(DD) \
  case DD:          \
    TB_TRY(dmma_mode >= 3 ? launch_tensor_dmma3<DD>(ctx, p, tt, n_leaf) : launch_tensor_dmma2<DD>(ctx, p, tt, n_leaf)); \
    launched = true; \
    break;
        TB_CASE(9) TB_CASE(10) TB_CASE(11) TB_CASE(12) TB_CASE(13) TB_CASE(14) TB_CASE(15) TB_CASE(16)
#undef TB_CASE
        default: break;
      }
    }
    if (!launched) switch (force_generic ? 0 : d) {
#define TB_CASE(DD) \
  case DD:          \
    TB_TRY(launch_tensor_t<DD>(ctx, p, tt, n_leaf)); \
    break;
      TB_CASE(2) TB_CASE(3) TB_CASE(4) TB_CASE(5) TB_CASE(6) TB_CASE(7) TB_CASE(8) TB_CASE(9)
      TB_CASE(10) TB_CASE(11) TB_CASE(12) TB_CASE(13) TB_CASE(14) TB_CASE(15)
#undef TB_CASE
      default: {  // degree-generic kernel (q = 15..19)
        const size_t smem = sizeof(double) * ((size_t)3 * d * dp + p.ncoef_pad + (size_t)n_row * dp + (size_t)d * d * d);
        TB_CUDA(ctx, cudaFuncSetAttribute(tensor_grid_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const size_t want = n_leaf < (size_t)ctx->n_sm * 16 ? n_leaf : (size_t)ctx->n_sm * 16;
        tensor_grid_eval_kernel<<<(unsigned)want, kTensorThreads, smem, ctx->stream>>>(p, tt);
        TB_CUDA(ctx, cudaGetLastError());
      }
    }
  }
  // The exceptions go through the generic path, whose launches the host sizes: it needs their number.
  // Which arrival points are exceptions is decided by the two leaf lists and the boundary condition
  // alone (integer boxes and node positions; no coefficient enters), so the count is read back from
  // the device ONCE per (grid leaf range, velocity tree, bc) and remembered: a step on unchanged trees
  // never waits for the device.  The first time, a kernel stores the count into a device-accessible
  // pinned word (a cudaMemcpy would queue on the device-to-host copy engine behind the previous
  // chunk's values in the pipelined host calls).
  size_t m = 0;
  if (n) {
    // (global hash: with ghost leaves the exception set also depends on the other ranks' shards)
    ExcCount key{vel->struct_hash ^ (vel->global_hash * 0x9e3779b97f4a7c15ull), grid->struct_hash, vel->n_leaf,
                 grid->n_leaf, leaf0, n_leaf, vel->q, periodic, 0};
    ExcCount *hit = nullptr;
    for (ExcCount &e : ctx->exc_cache)
      if (e.vel_hash == key.vel_hash && e.grid_hash == key.grid_hash && e.vel_leaves == key.vel_leaves &&
          e.grid_leaves == key.grid_leaves && e.leaf0 == key.leaf0 && e.n_leaf == key.n_leaf && e.q == key.q &&
          e.periodic == key.periodic)
        hit = &e;
    if (hit) {
      m = hit->count;
    } else {
      publish_count_kernel<<<1, 1, 0, ctx->stream>>>(exc_count, ctx->h_exc);
      TB_CUDA(ctx, cudaGetLastError());
      TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      m = key.count = ctx->h_exc[0];
      if (ctx->exc_cache.size() >= 256) ctx->exc_cache.erase(ctx->exc_cache.begin());
      ctx->exc_cache.push_back(key);
    }
  }
  ctx->last_exceptions = m;
  if (!m && !collective) return TBSLAS_OK;
  void *epos, *eval;
  // sized with headroom: the count differs from chunk to chunk and a growing workspace slot is a
  // cudaFree + cudaMalloc
  const size_t cap = m > n / 8 + 4096 ? m + m / 4 : n / 8 + 4096;
  TB_TRY(ws_get(ctx, WS_EXC_POS, sizeof(double) * 3 * cap, &epos));
  TB_TRY(ws_get(ctx, WS_EXC_VAL, sizeof(double) * 3 * cap, &eval));
  const unsigned g3 = (unsigned)((3 * m + 255) / 256);
  if (m) {
    StageScope sc(ctx, ST_TENSOR, 0.0, 1);
    if (virtual_x) {
      GridBase gb;
      make_grid_base(grid, leaf0, bc, &gb);
      gather_grid_points_kernel<<<g3, 256, 0, ctx->stream>>>(gb, (const uint32_t *)exc_idx, m, (double *)epos);
    } else {
      gather_points_kernel<<<g3, 256, 0, ctx->stream>>>(x, (const uint32_t *)exc_idx, m, (double *)epos);
    }
    TB_CUDA(ctx, cudaGetLastError());
  }
  TB_TRY(eval_tree_dev_points(vel, bc, (double *)epos, m, (double *)eval));
  if (m) {
    StageScope sc(ctx, ST_TENSOR, 0.0, 1);
    scatter_update_kernel<<<g3, 256, 0, ctx->stream>>>((const double *)epos, (const double *)eval,
                                                       (const uint32_t *)exc_idx, m, alpha,
                                                       periodic, virtual_x ? nullptr : x, out);
    TB_CUDA(ctx, cudaGetLastError());
  }
  return TBSLAS_OK;
}

}  // namespace tb

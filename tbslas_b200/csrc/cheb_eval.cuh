// Tensor-product Chebyshev evaluation of one leaf's polynomial at the points binned
// into that leaf.
//
// Replaces tbslas::EvalNodesLocal steps d-e (reference src/tree/tree_functor.h:248-383):
// coefficient unpack (:248-268), rescale to [-1,1] (:283-294), pvfmm::cheb_poly
// (:328-330), pvfmm::vec_eval (:27-84) and the AoS store (:376-383); optionally fused
// with the RK2 position update of tbslas::IntegrateRK2 (src/semilag/traj.inc:34-42).
//
// Design (B200, measured in profiles/r01_microbench_fp64_lds.txt):
//   * FP64-pipe bound: Ncoef - 1 DFMA per point per dof (one per coefficient, see RowPair).  DFMA
//     issues every 2 cycles per sub-partition with 8 cycles latency -> >= 4 independent chains.
//   * A warp-broadcast LDS.64 costs 1 SM-cycle, a warp DFMA 0.5: every coefficient read
//     from shared memory must feed >= 3-4 DFMAs, so each thread owns PPT points and the
//     T_k(x), T_j(y) of all of them live in registers (loops fully unrolled per degree);
//     T_i(z) is carried by its recurrence along the outer loop.
//   * One CTA = one (leaf, tile of THREADS*PPT points).  The leaf's contiguous
//     coefficient block (dof * Ncoef doubles, 1.3-23 KB) is staged into shared memory
//     by ONE bulk-TMA copy (cp.async.bulk + mbarrier) issued by thread 0 while all
//     threads gather their points and build the bases.
//   * Points are read through the leaf grouping (perm) and results scattered back to the
//     caller's AoS order; runs of 32 consecutive slots are consecutive points.
// Arithmetic parity: local coordinates and the Chebyshev recurrences use separate
// multiply/subtract exactly as the reference (bit-identical bases); the triangular
// contraction uses FMA (the reference: mul then add), so values differ from the CPU
// path only by rounding of the accumulation (<< 1e-12 relative to the field scale).
#pragma once
#include "common.cuh"
#include "gridbase.cuh"

namespace tb {

constexpr int kEvalThreads = 128;

struct EvalParams {
  const double *coeff;   // [(n_leaf+1)][dof][ncoef_pad]
  const double4 *geom;   // [(n_leaf+1)] {cx,cy,cz,2*2^depth}
  unsigned stride;       // doubles per leaf block = dof * ncoef_pad
  unsigned ncoef_pad;
  int dof;
  int n_bins;            // n_leaf + 1
  const double *pos;
  const uint32_t *perm;
  const uint32_t *bin_start;
  const uint32_t *tile_start;
  const int2 *tile_map;
  double *out;
  const double *base;
  double alpha;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// 1-D bulk TMA: global -> shared, completion counted in bytes on the mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                             uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// T_0..T_Q at xi, pvfmm::cheb_poly semantics: all zero when |xi| > 1; recurrence with
// separate multiply and subtract (bit-identical to the CPU path).
// Returns whether xi lies inside; T[0] is then 1 -- the contraction below never reads it (RowPair) and the
// callers zero the value of a point that is outside in any axis, which is what all-zero bases amount to.
template <int Q>
__device__ __forceinline__ bool cheb_basis(double xi, double (&T)[Q + 1]) {
  const bool in = fabs(xi) <= 1.0;
  const double x = in ? xi : 0.0;
  T[0] = in ? 1.0 : 0.0;
  if (Q >= 1) T[1] = x;
  const double x2 = 2.0 * x;
#pragma unroll
  for (int i = 2; i <= Q; i++) T[i] = __dsub_rn(__dmul_rn(x2, T[i - 1]), T[i - 2]);
  return in;
}

// ---- compile-time expansion of the triangular contraction ------------------------
// The loops of pvfmm::vec_eval (tree_functor.h:38-77) are expanded by template recursion
// (not `#pragma unroll`, which nvcc abandons for bodies this large): every coefficient
// offset, basis index and trip count below is a compile-time constant, so T_k(x), T_j(y)
// stay in registers and coefficient reads are LDS.128 at immediate offsets.
//
// Rows (i,j) and (i,j+1) are contracted together so a thread always has 2*PPT independent
// DFMA chains in flight (DFMA latency 8 cycles, issue interval 2).
struct CoefG4 { const double *p; };  // 32-byte aligned coefficient block in global memory
struct CoefReg { double a, b; };     // no loads at all (ceiling of the contraction)
#ifdef TB_EVALBENCH_CONST  // tools/evalbench.cu only
__constant__ double g_coef_const[2 * 1024];
template <int B> struct CoefConst {};  // constant bank at a compile-time offset: the DFMA's own operand
#endif

template <int Q, int PPT, bool PYS, bool PAIR, int I, int J, int CI>
struct RowPair {
  static constexpr int D = Q + 1;
  static constexpr int N0 = D - I - J;          // terms in row J
  static constexpr bool TWO = PAIR && (J + 1 < D - I);  // row J+1 exists (N0-1 >= 1 terms)
  static constexpr int NEXT = CI + N0 + (TWO ? N0 - 1 : 0);
  static __device__ __forceinline__ double coef(const double2 *C2, int i) {
    return (i & 1) ? C2[i >> 1].y : C2[i >> 1].x;
  }
  static __device__ __forceinline__ double coef(const double *C1, int i) { return C1[i]; }
  // experiments (tools/evalbench.cu): coefficients by 256-bit broadcast loads from global memory through
  // L1 (four per instruction; identical loads of one quad are merged by the compiler), or from registers
  static __device__ __forceinline__ double coef(CoefG4 c, int i) {
    double q0, q1, q2, q3;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(q0), "=d"(q1), "=d"(q2), "=d"(q3) : "l"(c.p + (i & ~3)));
    return (i & 3) == 0 ? q0 : ((i & 3) == 1 ? q1 : ((i & 3) == 2 ? q2 : q3));
  }
  static __device__ __forceinline__ double coef(CoefReg c, int i) { return (i & 1) ? c.b : c.a; }
#ifdef TB_EVALBENCH_CONST
  template <int B>
  static __device__ __forceinline__ double coef(CoefConst<B>, int i) { return g_coef_const[B + i]; }
#endif
  template <class CP>
  static __device__ __forceinline__ void run(CP C2,
                                             const double (&px)[PPT][D],
                                             const double (&py)[PYS ? 1 : PPT][D],
                                             const double *s_py, double (&v)[PPT]) {
    // T_0 = 1 wherever the point lies inside the leaf (and the caller zeroes the value of a point
    // outside, where every basis value is 0: cheb_poly semantics), so the k = 0 term of a row is the
    // coefficient itself, the j = 0 row of a plane is its row sum, and plane 0 is its own sum: fma(1, c, 0)
    // == c exactly, i.e. the same bits for rows + planes + 1 fewer DFMAs: Ncoef - 1 per point and component
    // (815 -> 679 at q = 14, 219 -> 164 at q = 8).
    double w0[PPT], w1[PPT];
#pragma unroll
    for (int k = 0; k < N0; k++) {
      const double c0 = coef(C2, CI + k);
#pragma unroll
      for (int s = 0; s < PPT; s++) w0[s] = (k == 0) ? c0 : fma(px[s][k], c0, w0[s]);
      if (TWO && k < N0 - 1) {
        const double c1 = coef(C2, CI + N0 + k);
#pragma unroll
        for (int s = 0; s < PPT; s++) w1[s] = (k == 0) ? c1 : fma(px[s][k], c1, w1[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < PPT; s++) {
      v[s] = (J == 0) ? w0[s] : fma(PYS ? s_py[(J * PPT + s) * kEvalThreads] : py[PYS ? 0 : s][J], w0[s], v[s]);
      if (TWO)
        v[s] = fma(PYS ? s_py[((J + 1) * PPT + s) * kEvalThreads] : py[PYS ? 0 : s][TWO ? J + 1 : J],
                   w1[s], v[s]);
    }
    if constexpr (J + (PAIR ? 2 : 1) < D - I)
      RowPair<Q, PPT, PYS, PAIR, I, J + (PAIR ? 2 : 1), NEXT>::run(C2, px, py, s_py, v);
  }
};

template <int Q, int PPT, bool PYS, bool PAIR, int I, int CI>
struct ZLevel {
  static constexpr int D = Q + 1;
  template <class CP>
  static __device__ __forceinline__ void run(CP C2,
                                             const double (&px)[PPT][D],
                                             const double (&py)[PYS ? 1 : PPT][D],
                                             const double *s_py, const double (&zc)[PPT],
                                             double (&tz0)[PPT], double (&tz1)[PPT], double (&u)[PPT]) {
    double pz[PPT], v[PPT];
#pragma unroll
    for (int s = 0; s < PPT; s++) {  // T_I(z) by its recurrence (mul, sub: bit-exact basis)
      if (I == 0)
        pz[s] = 1.0;  // (inside the leaf; see RowPair)
      else if (I == 1)
        pz[s] = zc[s];
      else
        pz[s] = __dsub_rn(__dmul_rn(2.0 * zc[s], tz1[s]), tz0[s]);
      tz0[s] = (I == 0) ? 0.0 : tz1[s];
      tz1[s] = pz[s];
      v[s] = 0.0;
    }
    RowPair<Q, PPT, PYS, PAIR, I, 0, CI>::run(C2, px, py, s_py, v);
#pragma unroll
    for (int s = 0; s < PPT; s++) u[s] = (I == 0) ? v[s] : fma(pz[s], v[s], u[s]);
    if constexpr (I + 1 < D)
      ZLevel<Q, PPT, PYS, PAIR, I + 1, CI + (D - I) * (D - I + 1) / 2>::run(C2, px, py, s_py, zc, tz0, tz1, u);
  }
};

// PYS: keep T_j(y) in shared memory ([j][point slot][thread], conflict free) instead of
// registers -- frees 2*(Q+1)*PPT registers so high degrees can run more points per thread.
// NB: batches of 32*PPT points a warp works through for the same leaf; the coordinates of
// batch b+1 are fetched while batch b is being contracted (hides the gather latency that
// two resident CTAs per SM cannot).
template <int Q, int PPT, int EPI, bool PYS, int NB, bool PAIR>
__global__ void __launch_bounds__(kEvalThreads)
cheb_eval_kernel(const EvalParams p) {
  constexpr int D = Q + 1;
  constexpr int NW = kEvalThreads / 32;
  extern __shared__ __align__(128) double s_coef[];
  __shared__ __align__(8) uint64_t s_bar;

  const unsigned n_tiles = __ldg(p.tile_start + p.n_bins);
  if (blockIdx.x >= n_tiles) return;
  const int2 tile = __ldg(p.tile_map + blockIdx.x);
  const int leaf = tile.x;
  const unsigned slot0 = (unsigned)tile.y;
  const unsigned bin_end = __ldg(p.bin_start + leaf + 1);
  const unsigned cnt = min((unsigned)(kEvalThreads * PPT * NB), bin_end - slot0);

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = p.stride * 8u;
    mbar_expect_tx(&s_bar, bytes);
    tma_bulk_g2s(s_coef, p.coeff + (size_t)leaf * p.stride, bytes, &s_bar);
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned wbase = warp * 32 * PPT;  // batch t = b*NW + warp starts at slot t*32*PPT
  if (wbase >= cnt) return;          // whole warp has no points (no barrier follows)

  const double4 g = p.geom[leaf];
  double *s_py = s_coef + p.stride + threadIdx.x;  // [(j*PPT + s)*kEvalThreads]
  const uint32_t *perm0 = p.perm + slot0;

  // coordinates of the first batch
  unsigned idx_n[PPT];
  double xn[PPT][3];
#pragma unroll
  for (int s = 0; s < PPT; s++) {
    const unsigned slot = wbase + s * 32 + lane;
    idx_n[s] = __ldg(perm0 + (slot < cnt ? slot : wbase));
    const double *x = p.pos + 3 * (size_t)idx_n[s];
    xn[s][0] = x[0];
    xn[s][1] = x[1];
    xn[s][2] = x[2];
  }
  bool waited = false;

#pragma unroll 1
  for (int b = 0; b < NB; b++) {
    unsigned idx[PPT];
    bool ok[PPT];
    bool inside[PPT];
    double px[PPT][D], py[PYS ? 1 : PPT][D], zc[PPT];
#pragma unroll
    for (int s = 0; s < PPT; s++) {
      idx[s] = idx_n[s];
      ok[s] = wbase + s * 32 + lane < cnt;
      // xi = (x - c) * 2 * 2^depth - 1, left to right (tree_functor.h:288-293); the
      // power-of-two scalings are exact, so one multiply by 2*2^depth is the same value.
      const double xi = __dadd_rn(__dmul_rn(__dsub_rn(xn[s][0], g.x), g.w), -1.0);
      const double yi = __dadd_rn(__dmul_rn(__dsub_rn(xn[s][1], g.y), g.w), -1.0);
      const double zi = __dadd_rn(__dmul_rn(__dsub_rn(xn[s][2], g.z), g.w), -1.0);
      bool in = cheb_basis<Q>(xi, px[s]);
      if (PYS) {
        in = cheb_basis<Q>(yi, py[0]) && in;
#pragma unroll
        for (int j = 0; j < D; j++) s_py[(j * PPT + s) * kEvalThreads] = py[0][j];
      } else {
        in = cheb_basis<Q>(yi, py[s]) && in;
      }
      const bool inz = fabs(zi) <= 1.0;
      zc[s] = inz ? zi : 0.0;
      inside[s] = in && inz;  // a point outside its leaf in any axis evaluates to 0
    }
    // prefetch the next batch of this warp (consumed after the contraction below)
    const unsigned wnext = wbase + NW * 32 * PPT;
    const bool more = (b + 1 < NB) && (wnext < cnt);
    if (more) {
#pragma unroll
      for (int s = 0; s < PPT; s++) {
        const unsigned slot = wnext + s * 32 + lane;
        idx_n[s] = __ldg(perm0 + (slot < cnt ? slot : wnext));
        const double *x = p.pos + 3 * (size_t)idx_n[s];
        xn[s][0] = x[0];
        xn[s][1] = x[1];
        xn[s][2] = x[2];
      }
    }
    if (!waited) {
      mbar_wait(&s_bar, 0);
      waited = true;
    }

#pragma unroll 1
    for (int l = 0; l < p.dof; l++) {
      const double2 *C2 = reinterpret_cast<const double2 *>(s_coef + l * p.ncoef_pad);
      double u[PPT], tz0[PPT], tz1[PPT];
#pragma unroll
      for (int s = 0; s < PPT; s++) u[s] = tz0[s] = tz1[s] = 0.0;
      ZLevel<Q, PPT, PYS, PAIR, 0, 0>::run(C2, px, py, s_py, zc, tz0, tz1, u);
#pragma unroll
      for (int s = 0; s < PPT; s++) {
        if (!inside[s]) u[s] = 0.0;
        if (ok[s]) {
          if (EPI == EPI_STORE) {
            p.out[(size_t)idx[s] * p.dof + l] = u[s];
          } else {  // x' = x0 + alpha * v, multiply then add as traj.inc:36,42
            const size_t o = 3 * (size_t)idx[s] + l;
            p.out[o] = __dadd_rn(p.base[o], __dmul_rn(p.alpha, u[s]));
          }
        }
      }
    }
    if (!more) break;
    wbase = wnext;
  }
}

constexpr int kEvalBatches = 1;  // batches of 32*PPT points per warp and tile (measured: 1 is best)

template <int Q, int PPT, bool PYS = false, int NB = kEvalBatches, bool PAIR = false>
int launch_cheb_eval_q(tbslas_ctx *ctx, const EvalArgs &a) {
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
  const size_t smem =
      (t->stride + (PYS ? (size_t)(Q + 1) * PPT * kEvalThreads : 0)) * sizeof(double);
  const unsigned grid = (unsigned)a.max_tiles;
  if (grid == 0) return TBSLAS_OK;
  if (a.epilogue == EPI_STORE) {
    auto k = cheb_eval_kernel<Q, PPT, EPI_STORE, PYS, NB, PAIR>;
    if (smem > 48 * 1024)
      TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kEvalThreads, smem, ctx->stream>>>(p);
  } else {
    auto k = cheb_eval_kernel<Q, PPT, EPI_AXPY, PYS, NB, PAIR>;
    if (smem > 48 * 1024)
      TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kEvalThreads, smem, ctx->stream>>>(p);
  }
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

// points per thread for each degree (register budget: 2*(Q+1)*PPT doubles of bases)
#ifndef TB_PPT14
#define TB_PPT14 2
#endif
constexpr int eval_ppt(int q) { return q <= 9 ? 4 : (q <= 12 ? 3 : (q <= 14 ? TB_PPT14 : 2)); }

}  // namespace tb

// tools/evalbench.cu -- steady-state rate of the Chebyshev contraction (no gathers, no
// tile prologue): the inner loops of cheb_eval.cuh run `iters` times per thread over two
// alternating coefficient blocks resident in shared memory.  Separates "the inner loop
// itself" from "tile prologue / occupancy" when reading the eval kernel's FP64-pipe
// utilisation.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
//        -Itbslas_b200/csrc -Iinclude tools/evalbench.cu -o tools/evalbench
#include <cstdio>
#include <cstdlib>
#include <vector>

#define TB_EVALBENCH_CONST 1
#include "cheb_eval.cuh"

using namespace tb;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

template <int Q, int PPT, bool PYS, bool PAIR, int MINB>
__global__ void __launch_bounds__(kEvalThreads, MINB)
steady_kernel(const double *__restrict__ coef, unsigned stride, double *__restrict__ out, int iters) {
  constexpr int D = Q + 1;
  extern __shared__ __align__(128) double s_coef[];
  for (unsigned i = threadIdx.x; i < 2 * stride; i += blockDim.x) s_coef[i] = coef[i];
  double *s_py = s_coef + 2 * stride + threadIdx.x;
  __syncthreads();
  double px[PPT][D], py[PYS ? 1 : PPT][D], zc[PPT];
#pragma unroll
  for (int s = 0; s < PPT; s++) {
    const double xi = -0.9 + 1.7e-3 * threadIdx.x + 0.11 * s, yi = 0.8 - 2.1e-3 * threadIdx.x - 0.07 * s;
    cheb_basis<Q>(xi, px[s]);
    if (PYS) {
      cheb_basis<Q>(yi, py[0]);
#pragma unroll
      for (int j = 0; j < D; j++) s_py[(j * PPT + s) * kEvalThreads] = py[0][j];
    } else {
      cheb_basis<Q>(yi, py[s]);
    }
    zc[s] = 0.3 - 1e-3 * threadIdx.x;
  }
  double acc[PPT];
#pragma unroll
  for (int s = 0; s < PPT; s++) acc[s] = 0;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    const double2 *C2 = reinterpret_cast<const double2 *>(s_coef + (it & 1) * stride);
    double u[PPT], tz0[PPT], tz1[PPT];
#pragma unroll
    for (int s = 0; s < PPT; s++) u[s] = tz0[s] = tz1[s] = 0.0;
    ZLevel<Q, PPT, PYS, PAIR, 0, 0>::run(C2, px, py, s_py, zc, tz0, tz1, u);
#pragma unroll
    for (int s = 0; s < PPT; s++) acc[s] += u[s];
  }
  double r = 0;
#pragma unroll
  for (int s = 0; s < PPT; s++) r += acc[s];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// the same loop with the coefficients (SRC 1) read by 256-bit broadcast loads from global memory through
// L1, or (SRC 2) taken from registers: what the contraction does when no load competes with the DFMAs
template <int Q, int PPT, int SRC, int MINB>
__global__ void __launch_bounds__(kEvalThreads, MINB)
steady_src_kernel(const double *__restrict__ coef, unsigned stride, double *__restrict__ out, int iters) {
  constexpr int D = Q + 1;
  double px[PPT][D], py[PPT][D], zc[PPT];
#pragma unroll
  for (int s = 0; s < PPT; s++) {
    const double xi = -0.9 + 1.7e-3 * threadIdx.x + 0.11 * s, yi = 0.8 - 2.1e-3 * threadIdx.x - 0.07 * s;
    cheb_basis<Q>(xi, px[s]);
    cheb_basis<Q>(yi, py[s]);
    zc[s] = 0.3 - 1e-3 * threadIdx.x;
  }
  double acc[PPT];
#pragma unroll
  for (int s = 0; s < PPT; s++) acc[s] = 0;
  const unsigned gstride = (stride + 3) & ~3u;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    double u[PPT], tz0[PPT], tz1[PPT];
#pragma unroll
    for (int s = 0; s < PPT; s++) {
      u[s] = tz0[s] = tz1[s] = 0.0;
      px[s][1] += 1e-13;  // every row of two or more terms reads px[1]: the contraction is not loop invariant
    }
    if (SRC == 1) {
      CoefG4 c{coef + (size_t)((it + blockIdx.x) & 7) * gstride};
      ZLevel<Q, PPT, false, false, 0, 0>::run(c, px, py, nullptr, zc, tz0, tz1, u);
    } else if (SRC == 3) {  // two blocks, alternating (every DFMA takes its coefficient as a constant-bank operand)
      if (it & 1)
        ZLevel<Q, PPT, false, false, 0, 0>::run(CoefConst<1024>(), px, py, nullptr, zc, tz0, tz1, u);
      else
        ZLevel<Q, PPT, false, false, 0, 0>::run(CoefConst<0>(), px, py, nullptr, zc, tz0, tz1, u);
    } else {
      CoefReg c{1e-3 * it, 0.5 - 1e-3 * it};
      ZLevel<Q, PPT, false, false, 0, 0>::run(c, px, py, nullptr, zc, tz0, tz1, u);
    }
#pragma unroll
    for (int s = 0; s < PPT; s++) acc[s] += u[s];
  }
  double r = 0;
#pragma unroll
  for (int s = 0; s < PPT; s++) r += acc[s];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// Register-footprint proxies: the real contraction code, but the basis arrays are filled from a few values so
// that the compiler keeps fewer registers (PXM: px[k] = px[k mod PXM], PYM likewise).  Arithmetic is meaningless;
// the instruction mix, the dependency chains and the coefficient loads are the real ones.  Answers "what would
// three or four points per thread reach if the bases needed fewer registers".
template <int Q, int PPT, int PXM, int PYM, int MINB, bool NOLOAD = false>
__global__ void __launch_bounds__(kEvalThreads, MINB)
steady_proxy_kernel(const double *__restrict__ coef, unsigned stride, double *__restrict__ out, int iters) {
  constexpr int D = Q + 1;
  extern __shared__ __align__(128) double s_coef[];
  for (unsigned i = threadIdx.x; i < 2 * stride; i += blockDim.x) s_coef[i] = coef[i];
  __syncthreads();
  double bx[PPT][PXM], by[PPT][PYM], zc[PPT];
#pragma unroll
  for (int s = 0; s < PPT; s++) {
#pragma unroll
    for (int k = 0; k < PXM; k++) bx[s][k] = 0.9 - 1.7e-3 * threadIdx.x + 0.011 * s - 0.05 * k;
#pragma unroll
    for (int k = 0; k < PYM; k++) by[s][k] = 0.8 - 2.1e-3 * threadIdx.x - 0.007 * s - 0.04 * k;
    zc[s] = 0.3 - 1e-3 * threadIdx.x;
  }
  double acc[PPT];
#pragma unroll
  for (int s = 0; s < PPT; s++) acc[s] = 0;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    double px[PPT][D], py[PPT][D];
#pragma unroll
    for (int s = 0; s < PPT; s++) {
      bx[s][1 % PXM] += 1e-13;
#pragma unroll
      for (int k = 0; k < D; k++) {
        px[s][k] = bx[s][k % PXM];
        py[s][k] = by[s][k % PYM];
      }
    }
    const double2 *C2 = reinterpret_cast<const double2 *>(s_coef + (it & 1) * stride);
    double u[PPT], tz0[PPT], tz1[PPT];
#pragma unroll
    for (int s = 0; s < PPT; s++) u[s] = tz0[s] = tz1[s] = 0.0;
    if (NOLOAD) {  // coefficients from two registers: the DFMA stream alone
      CoefReg c{1e-3 * it + 0.25, 0.5 - 1e-3 * it};
      ZLevel<Q, PPT, false, false, 0, 0>::run(c, px, py, nullptr, zc, tz0, tz1, u);
    } else {
      ZLevel<Q, PPT, false, false, 0, 0>::run(C2, px, py, nullptr, zc, tz0, tz1, u);
    }
#pragma unroll
    for (int s = 0; s < PPT; s++) acc[s] += u[s];
  }
  double r = 0;
#pragma unroll
  for (int s = 0; s < PPT; s++) r += acc[s];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int Q, int PPT, int PXM, int PYM, int MINB, bool NOLOAD = false>
void run_proxy(const char *name, int ctas_per_sm, int n_sm, const double *d_coef, double *d_out) {
  constexpr int D = Q + 1;
  const unsigned ncoef = D * (D + 1) * (D + 2) / 6, stride = ncoef + (ncoef & 1);
  const size_t smem = 2 * stride * sizeof(double);
  auto k = steady_proxy_kernel<Q, PPT, PXM, PYM, MINB, NOLOAD>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, k));
  const int iters = 400, grid = n_sm * ctas_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0));
    k<<<grid, kEvalThreads, smem>>>(d_coef, stride, d_out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  const double dfma = (double)(ncoef - 1) * PPT * kEvalThreads * (double)grid * iters;
  printf("%-28s q=%d ppt=%d px regs %d py regs %d regs=%3d local=%zu ctas/sm=%d : %.3f ms  %.2f TFLOP/s (executed DFMA)\n", name, Q,
         PPT, PXM, PYM, fa.numRegs, (size_t)fa.localSizeBytes, ctas_per_sm, best, 2 * dfma / best * 1e-9);
}

// ---------------------------------------------------------------------------------------------------
// "Plane at a time, k outermost": the experiment behind DESIGN 3.2's operand-delivery argument.  For a
// fixed plane i and a fixed k all rows j with more than k terms are advanced together, one point after the
// other, so consecutive DFMAs share the MULTIPLICAND T_k(x) of that point (operand reuse) and read two fresh
// operands (coefficient, accumulator) instead of three.  Same chains, same order inside every chain, same
// bits; it needs one accumulator per (point, row of the plane) and the plane's coefficients in k-major order
// (LAYOUT 1: [k][j], paired LDS.128) -- or single LDS.64 from the shipped [j][k] order (LAYOUT 0).
template <int Q, int PPT, int LAYOUT, int I, int CI>
struct PlaneK {
  static constexpr int D = Q + 1, M = D - I;  // rows of the plane: j = 0..M-1, row j has M - j terms
  static __device__ __forceinline__ void run(const double *C, const double (&px)[PPT][D], const double (&py)[PPT][D],
                                             const double (&zc)[PPT], double (&tz0)[PPT], double (&tz1)[PPT],
                                             double (&u)[PPT]) {
    double acc[PPT][M];
#pragma unroll
    for (int k = 0; k < M; k++) {
      constexpr int dummy = 0;
      (void)dummy;
      const int nrow = M - k;  // rows with a k-th term
      double c[M];
      if (LAYOUT == 1) {       // [k][j]: the nrow coefficients of this k are contiguous
        int off = CI;
#pragma unroll
        for (int kk = 0; kk < k; kk++) off += M - kk;
#pragma unroll
        for (int j = 0; j < nrow; j++) {
          const int e = off + j;
          const double2 t = reinterpret_cast<const double2 *>(C)[e >> 1];
          c[j] = (e & 1) ? t.y : t.x;
        }
      } else {                 // shipped order: row j starts at CI + sum_{j' < j} (M - j')
#pragma unroll
        for (int j = 0; j < nrow; j++) {
          int off = CI;
#pragma unroll
          for (int jj = 0; jj < j; jj++) off += M - jj;
          c[j] = C[off + k];
        }
      }
#pragma unroll
      for (int s = 0; s < PPT; s++)
#pragma unroll
        for (int j = 0; j < nrow; j++) acc[s][j] = (k == 0) ? c[j] : fma(px[s][k], c[j], acc[s][j]);
    }
#pragma unroll
    for (int s = 0; s < PPT; s++) {
      double pz;
      if (I == 0) pz = 1.0;
      else if (I == 1) pz = zc[s];
      else pz = __dsub_rn(__dmul_rn(2.0 * zc[s], tz1[s]), tz0[s]);
      tz0[s] = (I == 0) ? 0.0 : tz1[s];
      tz1[s] = pz;
      double v = acc[s][0];
#pragma unroll
      for (int j = 1; j < M; j++) v = fma(py[s][j], acc[s][j], v);
      u[s] = (I == 0) ? v : fma(pz, v, u[s]);
    }
    if constexpr (I + 1 < D) PlaneK<Q, PPT, LAYOUT, I + 1, CI + M * (M + 1) / 2>::run(C, px, py, zc, tz0, tz1, u);
  }
};

template <int Q, int PPT, int LAYOUT, int MINB>
__global__ void __launch_bounds__(kEvalThreads, MINB)
steady_planek_kernel(const double *__restrict__ coef, unsigned stride, double *__restrict__ out, int iters) {
  constexpr int D = Q + 1;
  extern __shared__ __align__(128) double s_coef[];
  for (unsigned i = threadIdx.x; i < 2 * stride; i += blockDim.x) s_coef[i] = coef[i];
  __syncthreads();
  double px[PPT][D], py[PPT][D], zc[PPT];
#pragma unroll
  for (int s = 0; s < PPT; s++) {
    const double xi = -0.9 + 1.7e-3 * threadIdx.x + 0.11 * s, yi = 0.8 - 2.1e-3 * threadIdx.x - 0.07 * s;
    cheb_basis<Q>(xi, px[s]);
    cheb_basis<Q>(yi, py[s]);
    zc[s] = 0.3 - 1e-3 * threadIdx.x;
  }
  double acc[PPT];
#pragma unroll
  for (int s = 0; s < PPT; s++) acc[s] = 0;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    double u[PPT], tz0[PPT], tz1[PPT];
#pragma unroll
    for (int s = 0; s < PPT; s++) {
      u[s] = tz0[s] = tz1[s] = 0.0;
      px[s][1] += 1e-13;
    }
    PlaneK<Q, PPT, LAYOUT, 0, 0>::run(s_coef + (it & 1) * stride, px, py, zc, tz0, tz1, u);
#pragma unroll
    for (int s = 0; s < PPT; s++) acc[s] += u[s];
  }
  double r = 0;
#pragma unroll
  for (int s = 0; s < PPT; s++) r += acc[s];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int Q, int PPT, int LAYOUT, int MINB>
void run_planek(const char *name, int ctas_per_sm, int n_sm, const double *d_coef, double *d_out) {
  constexpr int D = Q + 1;
  const unsigned ncoef = D * (D + 1) * (D + 2) / 6, stride = ncoef + (ncoef & 1);
  const size_t smem = 2 * stride * sizeof(double);
  auto k = steady_planek_kernel<Q, PPT, LAYOUT, MINB>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, k));
  const int iters = 400, grid = n_sm * ctas_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0));
    k<<<grid, kEvalThreads, smem>>>(d_coef, stride, d_out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  const double dfma = (double)(ncoef - 1) * PPT * kEvalThreads * (double)grid * iters;
  printf("%-28s q=%d ppt=%d layout=%d regs=%3d local=%zu ctas/sm=%d : %.3f ms  %.2f TFLOP/s (executed DFMA)\n", name, Q, PPT,
         LAYOUT, fa.numRegs, (size_t)fa.localSizeBytes, ctas_per_sm, best, 2 * dfma / best * 1e-9);
}

template <int Q, int PPT, int SRC, int MINB>
void run_src(const char *name, int ctas_per_sm, int n_sm, const double *d_coef, double *d_out) {
  constexpr int D = Q + 1;
  const unsigned ncoef = D * (D + 1) * (D + 2) / 6, stride = ncoef + (ncoef & 1);
  auto k = steady_src_kernel<Q, PPT, SRC, MINB>;
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, k));
  const int iters = 400, grid = n_sm * ctas_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0));
    k<<<grid, kEvalThreads>>>(d_coef, stride, d_out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  const double dfma = (double)(ncoef - 1) * PPT  /* T_0 = 1: Ncoef - 1 DFMAs per point */ * kEvalThreads * (double)grid * iters;
  printf("%-28s q=%d ppt=%d src=%d regs=%3d ctas/sm=%d : %.3f ms  %.2f TFLOP/s (executed DFMA)\n", name, Q, PPT,
         SRC, fa.numRegs, ctas_per_sm, best, 2 * dfma / best * 1e-9);
}

template <int Q, int PPT, bool PYS, bool PAIR, int MINB>
void run(const char *name, int ctas_per_sm, int n_sm, const double *d_coef, double *d_out) {
  constexpr int D = Q + 1;
  const unsigned ncoef = D * (D + 1) * (D + 2) / 6, stride = ncoef + (ncoef & 1);
  const size_t smem = (2 * stride + (PYS ? (size_t)D * PPT * kEvalThreads : 0)) * sizeof(double);
  auto k = steady_kernel<Q, PPT, PYS, PAIR, MINB>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kEvalThreads, smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, k));
  const int iters = 400, grid = n_sm * ctas_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0));
    k<<<grid, kEvalThreads, smem>>>(d_coef, stride, d_out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  const double dfma = (double)(ncoef - 1) * PPT  /* T_0 = 1: Ncoef - 1 DFMAs per point */ * kEvalThreads * (double)grid * iters;
  printf("%-28s q=%d ppt=%d pys=%d pair=%d regs=%3d occ=%d ctas/sm=%d : %.3f ms  %.2f TFLOP/s (executed DFMA)\n",
         name, Q, PPT, (int)PYS, (int)PAIR, fa.numRegs, occ, ctas_per_sm, best, 2 * dfma / best * 1e-9);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int n_sm = prop.multiProcessorCount;
  std::vector<double> h(8192);
  for (size_t i = 0; i < h.size(); i++) h[i] = 1e-3 * ((i * 2654435761u) % 1000) - 0.5;
  double *d_coef, *d_out;
  CK(cudaMalloc(&d_coef, h.size() * 8));
  CK(cudaMemcpy(d_coef, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_out, sizeof(double) * kEvalThreads * n_sm * 8));
  printf("device %s, %d SMs\n", prop.name, n_sm);
  run_planek<14, 2, 1, 1>("plane-k, [k][j] layout", 1, n_sm, d_coef, d_out);
  run_planek<14, 2, 1, 1>("plane-k, [k][j] layout", 2, n_sm, d_coef, d_out);
  run_planek<14, 2, 0, 1>("plane-k, shipped layout", 2, n_sm, d_coef, d_out);
  run_planek<8, 4, 1, 1>("q8 plane-k, [k][j] layout", 2, n_sm, d_coef, d_out);
  run_planek<8, 2, 1, 1>("q8 ppt2 plane-k, [k][j]", 2, n_sm, d_coef, d_out);
  run_proxy<14, 2, 15, 15, 1>("proxy ppt2 full bases", 2, n_sm, d_coef, d_out);
  run_proxy<14, 3, 15, 15, 1>("proxy ppt3 full bases", 2, n_sm, d_coef, d_out);
  run_proxy<14, 3, 8, 15, 1>("proxy ppt3 px/2", 2, n_sm, d_coef, d_out);
  run_proxy<14, 3, 15, 2, 1>("proxy ppt3 py by 2 regs", 2, n_sm, d_coef, d_out);
  run_proxy<14, 3, 8, 8, 1>("proxy ppt3 px/2 py/2", 2, n_sm, d_coef, d_out);
  run_proxy<14, 4, 8, 8, 1>("proxy ppt4 px/2 py/2", 2, n_sm, d_coef, d_out);
  run_proxy<14, 4, 8, 2, 1>("proxy ppt4 px/2 py 2 regs", 2, n_sm, d_coef, d_out);
  run_proxy<14, 4, 2, 2, 1>("proxy ppt4 2+2 regs", 2, n_sm, d_coef, d_out);
  run_proxy<14, 2, 15, 15, 1, true>("proxy ppt2 full, NO LOADS", 2, n_sm, d_coef, d_out);
  run_proxy<14, 4, 8, 8, 1, true>("proxy ppt4 px/2 py/2, NO LOADS", 2, n_sm, d_coef, d_out);
  run_proxy<14, 4, 2, 2, 1, true>("proxy ppt4 2+2, NO LOADS", 2, n_sm, d_coef, d_out);
  run_proxy<14, 4, 2, 2, 1, true>("proxy ppt4 2+2, NO LOADS", 4, n_sm, d_coef, d_out);
  run_proxy<14, 4, 2, 2, 3, true>("proxy ppt4 2+2, NO LOADS", 3, n_sm, d_coef, d_out);
  // occupancy sweep of the shipped q=14 shape (255 regs -> 2 CTAs/SM)
  run<14, 2, false, false, 1>("ppt2 regs", 1, n_sm, d_coef, d_out);
  run_src<14, 2, 1, 1>("ppt2 LDG.256 via L1", 1, n_sm, d_coef, d_out);
  run_src<14, 2, 1, 1>("ppt2 LDG.256 via L1", 2, n_sm, d_coef, d_out);
  CK(cudaMemcpyToSymbol(g_coef_const, h.data(), sizeof(double) * 2048));
  run_src<14, 2, 3, 1>("ppt2 constant operands", 1, n_sm, d_coef, d_out);
  run_src<14, 2, 3, 1>("ppt2 constant operands", 2, n_sm, d_coef, d_out);
  run_src<14, 3, 3, 1>("ppt3 constant operands", 2, n_sm, d_coef, d_out);
  run_src<8, 4, 3, 1>("q8 ppt4 constant operands", 2, n_sm, d_coef, d_out);
  run_src<8, 4, 1, 1>("q8 ppt4 LDG.256 via L1", 2, n_sm, d_coef, d_out);
  run<14, 2, false, false, 1>("ppt2 regs", 2, n_sm, d_coef, d_out);
  run<14, 2, false, true, 1>("ppt2 regs pair", 1, n_sm, d_coef, d_out);
  run<14, 2, false, true, 1>("ppt2 regs pair", 2, n_sm, d_coef, d_out);
  run<14, 2, false, false, 3>("ppt2 regs minb3", 3, n_sm, d_coef, d_out);
  run<14, 2, true, false, 3>("ppt2 pys minb3", 3, n_sm, d_coef, d_out);
  run<14, 2, true, false, 4>("ppt2 pys minb4", 4, n_sm, d_coef, d_out);
  run<14, 3, false, false, 1>("ppt3 regs", 2, n_sm, d_coef, d_out);
  run<14, 3, true, false, 1>("ppt3 pys", 2, n_sm, d_coef, d_out);
  run<14, 3, true, false, 3>("ppt3 pys minb3", 3, n_sm, d_coef, d_out);
  run<14, 4, true, false, 1>("ppt4 pys", 2, n_sm, d_coef, d_out);
  run<14, 4, true, true, 1>("ppt4 pys pair", 2, n_sm, d_coef, d_out);
  run<8, 4, false, false, 1>("q8 ppt4 regs", 2, n_sm, d_coef, d_out);
  run<8, 4, false, false, 3>("q8 ppt4 regs minb3", 3, n_sm, d_coef, d_out);
  run<8, 4, false, false, 4>("q8 ppt4 regs minb4", 4, n_sm, d_coef, d_out);
  return 0;
}

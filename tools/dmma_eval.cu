// tools/dmma_eval.cu -- prototype + steady-state rate of a DMMA (mma.sync m8n8k4 f64)
// formulation of the dof-3, q=14 Chebyshev evaluation, checked against a host sum.
//
// Scalar kernel (cheb_eval_wt.cuh): every coefficient is a broadcast LDS.128 feeding 2*PPT
// DFMAs; the register-file write port is shared between the LDS return data and the FP64
// results, which caps the contraction at ~83 % of the DFMA peak.
//
// Tensor formulation.  For one leaf and a batch of points p,
//     W[p][(i,l)] = sum over (j,k), j+k <= q-i, of  (Ty_j(p) Tx_k(p)) * C[l][i][j][k]
// is a GEMM  [points x 120 (j,k) pairs] x [120 x 45 (i,l) columns]; ordering the (j,k) pairs by
// j+k makes the non-zero part of every column a PREFIX of the K range, so an 8-column tile
// needs only ceil(prefix/4) k-steps: 30+23+14+7+4+1 = 79 DMMAs per 8 points for all three
// components (2528 MAC/point against 2040 useful, 81 %), against 3*815 = 2445 DFMAs/point in the
// scalar kernel.  The A fragment (the products Ty_j Tx_k) is formed once per k-step and shared by
// the column tiles; the B fragment is one conflict-free lane-private LDS.64 per DMMA.  The last
// contraction, u[p][l] = sum_i Tz_i(p) W[p][(i,l)], runs on the accumulator fragments.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 tools/dmma_eval.cu -o tools/dmma_eval
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int Q = 14, D = Q + 1, NCOEF = D * (D + 1) * (D + 2) / 6, DOF = 3;
constexpr int NJK = D * (D + 1) / 2;   // 120 (j,k) pairs
constexpr int KSTEPS = NJK / 4;        // 30
constexpr int NTILE = 6;               // 45 columns (i,l) padded to 48
constexpr int S = 36;                  // row stride (doubles) of the basis tables: == 4 mod 16
constexpr int kThreads = 128;

__host__ __device__ constexpr int tile_ksteps(int T) {
  const int imin = (8 * T) / 3, s = Q - imin, cnt = (s + 1) * (s + 2) / 2;
  return (cnt + 3) / 4;
}
__host__ __device__ constexpr int tiles_at(int t) {  // column tiles still active at k-step t
  int n = 0;
  for (int T = 0; T < NTILE; T++) n += tile_ksteps(T) > t ? 1 : 0;
  return n;
}
__host__ __device__ constexpr int block_of(int t, int T) {  // index of the (t,T) B block
  int n = 0;
  for (int u = 0; u < t; u++) n += tiles_at(u);
  return n + T;  // tiles are active in the order 0..tiles_at(t)-1
}
constexpr int NBLOCK = block_of(KSTEPS, 0);  // 79
static_assert(NBLOCK == 79, "k-step count");

__constant__ uint32_t c_jk[KSTEPS];  // 4 x (j << 4 | k) per k-step, byte kk = lane % 4

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void basis_to_smem(double xi, double *s_t, int lane) {
  const bool in = fabs(xi) <= 1.0;
  const double x = in ? xi : 0.0, x2 = 2.0 * x;
  double t0 = in ? 1.0 : 0.0, t1 = x;
  s_t[lane] = t0;
  s_t[S + lane] = t1;
#pragma unroll
  for (int i = 2; i <= Q; i++) {
    const double t2 = __dsub_rn(__dmul_rn(x2, t1), t0);
    s_t[i * S + lane] = t2;
    t0 = t1;
    t1 = t2;
  }
}

template <int T0, int T>
struct TileLoop {
  static __device__ __forceinline__ void run(const double *s_B, int lane, const double (&a)[4],
                                             double (&acc)[4][NTILE][2]) {
    if (tile_ksteps(T) > T0) {
      const double b = s_B[block_of(T0, T) * 32 + lane];
#pragma unroll
      for (int r = 0; r < 4; r++) dmma(acc[r][T][0], acc[r][T][1], a[r], b);
    }
    TileLoop<T0, T + 1>::run(s_B, lane, a, acc);
  }
};
template <int T0>
struct TileLoop<T0, NTILE> {
  static __device__ __forceinline__ void run(const double *, int, const double (&)[4], double (&)[4][NTILE][2]) {}
};

template <int T0>
struct KLoop {
  static __device__ __forceinline__ void run(const double *s_B, const double *s_tx, const double *s_ty,
                                             int lane, int prow, int kk, double (&acc)[4][NTILE][2]) {
    const uint32_t jk = (c_jk[T0] >> (8 * kk)) & 0xffu;
    const double *ty = s_ty + (jk >> 4) * S + prow, *tx = s_tx + (jk & 15u) * S + prow;
    double a[4];
#pragma unroll
    for (int r = 0; r < 4; r++) a[r] = __dmul_rn(ty[8 * r], tx[8 * r]);
    TileLoop<T0, 0>::run(s_B, lane, a, acc);
    KLoop<T0 + 1>::run(s_B, s_tx, s_ty, lane, prow, kk, acc);
  }
};
template <>
struct KLoop<KSTEPS> {
  static __device__ __forceinline__ void run(const double *, const double *, const double *, int, int, int,
                                             double (&)[4][NTILE][2]) {}
};

// pts: SoA [3][n] local coordinates in [-1,1]; out: [n][3]
#ifndef MINB
#define MINB 3
#endif
__global__ void __launch_bounds__(kThreads, MINB)
dmma_eval_kernel(const double *__restrict__ coef, const uint16_t *__restrict__ bidx,
                 const double *__restrict__ pts, size_t n, double *__restrict__ out, int tiles_per_cta) {
  extern __shared__ __align__(16) double smem[];
  double *s_B = smem;                                       // [79][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *s_tx = smem + NBLOCK * 32 + warp * 3 * D * S;     // [15][S] per warp
  double *s_ty = s_tx + D * S, *s_tz = s_ty + D * S;
  for (int e = threadIdx.x; e < NBLOCK * 32; e += kThreads) {
    const uint16_t src = bidx[e];
    s_B[e] = src == 0xffffu ? 0.0 : coef[src];
  }
  __syncthreads();
  const int kk = lane & 3, prow = lane >> 2;
  for (int it = 0; it < tiles_per_cta; it++) {
    const size_t p = ((size_t)blockIdx.x * tiles_per_cta + it) * kThreads + threadIdx.x;
    const bool ok = p < n;
    basis_to_smem(ok ? pts[p] : 2.0, s_tx, lane);
    basis_to_smem(ok ? pts[n + p] : 2.0, s_ty, lane);
    basis_to_smem(ok ? pts[2 * n + p] : 2.0, s_tz, lane);
    __syncwarp();
    double acc[4][NTILE][2];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int T = 0; T < NTILE; T++) acc[r][T][0] = acc[r][T][1] = 0.0;
    KLoop<0>::run(s_B, s_tx, s_ty, lane, prow, kk, acc);
    // u[p][l] = sum_i Tz_i(p) W[p][(i,l)]; this lane holds columns n = 8T + 2kk + e of its rows
    const int r0 = (2 * kk) % 3;
    const size_t pbase = p - lane;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      double v[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int T = 0; T < NTILE; T++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int nn = 8 * T + e + 2 * kk;
          const int i = min((nn * 43) >> 7, Q);  // nn / 3; the 3 padding columns hold zeros
          v[(2 * T + e) % 3] = fma(s_tz[i * S + prow + 8 * r], acc[r][T][e], v[(2 * T + e) % 3]);
        }
      // accumulator v[m] belongs to component l = (m + r0) % 3
      double u0 = r0 == 0 ? v[0] : (r0 == 1 ? v[2] : v[1]);
      double u1 = r0 == 0 ? v[1] : (r0 == 1 ? v[0] : v[2]);
      double u2 = r0 == 0 ? v[2] : (r0 == 1 ? v[1] : v[0]);
#pragma unroll
      for (int d = 1; d <= 2; d <<= 1) {
        u0 += __shfl_xor_sync(0xffffffffu, u0, d);
        u1 += __shfl_xor_sync(0xffffffffu, u1, d);
        u2 += __shfl_xor_sync(0xffffffffu, u2, d);
      }
      const size_t q = pbase + prow + 8 * r;
      if (kk < 3 && q < n) out[3 * q + kk] = kk == 0 ? u0 : (kk == 1 ? u1 : u2);
    }
    __syncwarp();
  }
}

static int tri(int i, int j, int k) {  // packed order: i (z) outer, j (y), k (x) inner, i+j+k <= q
  int off = 0;
  for (int a = 0; a < i; a++) off += (D - a) * (D - a + 1) / 2;
  for (int b = 0; b < j; b++) off += D - i - b;
  return off + k;
}

int main(int argc, char **argv) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int n_sm = prop.multiProcessorCount;
  // (j,k) pairs by s = j+k ascending
  std::vector<int> J, K;
  for (int s = 0; s <= Q; s++)
    for (int j = 0; j <= s; j++) {
      J.push_back(j);
      K.push_back(s - j);
    }
  uint32_t h_jk[KSTEPS];
  for (int t = 0; t < KSTEPS; t++) {
    h_jk[t] = 0;
    for (int kk = 0; kk < 4; kk++) h_jk[t] |= (uint32_t)((J[4 * t + kk] << 4) | K[4 * t + kk]) << (8 * kk);
  }
  CK(cudaMemcpyToSymbol(c_jk, h_jk, sizeof(h_jk)));
  std::vector<uint16_t> bidx((size_t)NBLOCK * 32, 0xffff);
  for (int t = 0; t < KSTEPS; t++)
    for (int T = 0; T < NTILE; T++) {
      if (tile_ksteps(T) <= t) continue;
      const int b = block_of(t, T);
      for (int lane = 0; lane < 32; lane++) {
        const int kk = lane & 3, nn = 8 * T + (lane >> 2), m = 4 * t + kk;
        if (nn >= 45) continue;
        const int i = nn / 3, l = nn % 3, j = J[m], k = K[m];
        if (i + j + k > Q) continue;
        bidx[(size_t)b * 32 + lane] = (uint16_t)(l * NCOEF + tri(i, j, k));
      }
    }
  std::vector<double> h_coef(DOF * NCOEF);
  srand(7);
  for (auto &c : h_coef) c = (rand() / (double)RAND_MAX - 0.5) * 0.1;
  const int tiles_per_cta = argc > 1 ? atoi(argv[1]) : 64, ctas = MINB * n_sm;
  const size_t n = (size_t)ctas * tiles_per_cta * kThreads;
  std::vector<double> h_pts(3 * n);
  for (auto &x : h_pts) x = 2.0 * (rand() / (double)RAND_MAX) - 1.0;
  double *d_coef, *d_pts, *d_out;
  uint16_t *d_bidx;
  CK(cudaMalloc(&d_coef, h_coef.size() * 8));
  CK(cudaMalloc(&d_pts, h_pts.size() * 8));
  CK(cudaMalloc(&d_out, 3 * n * 8));
  CK(cudaMalloc(&d_bidx, bidx.size() * 2));
  CK(cudaMemcpy(d_coef, h_coef.data(), h_coef.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_pts, h_pts.data(), h_pts.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_bidx, bidx.data(), bidx.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = (NBLOCK * 32 + (kThreads / 32) * 3 * D * S) * sizeof(double);
  CK(cudaFuncSetAttribute(dmma_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dmma_eval_kernel, kThreads, smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, dmma_eval_kernel));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0));
    dmma_eval_kernel<<<ctas, kThreads, smem>>>(d_coef, d_bidx, d_pts, n, d_out, tiles_per_cta);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  std::vector<double> h_out(3 * n);
  CK(cudaMemcpy(h_out.data(), d_out, 3 * n * 8, cudaMemcpyDeviceToHost));
  // host check on a sample
  double max_err = 0, scale = 0;
  for (size_t s = 0; s < 2000; s++) {
    const size_t p = (s * 7919) % n;
    double T[3][D];
    for (int a = 0; a < 3; a++) {
      const double x = h_pts[a * n + p];
      T[a][0] = 1;
      T[a][1] = x;
      for (int i = 2; i <= Q; i++) T[a][i] = 2 * x * T[a][i - 1] - T[a][i - 2];
    }
    for (int l = 0; l < DOF; l++) {
      long double u = 0;
      for (int i = 0; i <= Q; i++)
        for (int j = 0; i + j <= Q; j++)
          for (int k = 0; i + j + k <= Q; k++)
            u += (long double)h_coef[l * NCOEF + tri(i, j, k)] * T[0][k] * T[1][j] * T[2][i];
      max_err = fmax(max_err, fabs((double)u - h_out[3 * p + l]));
      scale = fmax(scale, fabs((double)u));
    }
  }
  const double model = (9.0 * D + 2.0 * DOF * NCOEF) * (double)n;
  printf("device %s, %d SMs; regs %d, occupancy %d CTAs/SM, smem %zu B\n", prop.name, n_sm, fa.numRegs, occ, smem);
  printf("dmma eval q=14 dof=3: %zu points in %.3f ms = %.2f G pts/s, %.2f TFLOP/s (model flops), "
         "%.2f TFLOP/s executed in DMMA; max err %.3e (scale %.3e, rel %.2e)\n",
         n, best, n / best * 1e-6, model / best * 1e-9, 2.0 * NBLOCK * 256 / 8 * n / best * 1e-9, max_err, scale,
         max_err / scale);
  return max_err / scale < 1e-12 ? 0 : 1;
}

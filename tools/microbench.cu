// tools/microbench.cu -- B200 design-point probes for the Chebyshev evaluation kernel.
//
// Questions this answers (numbers go to DESIGN.md / profiles/):
//   1. DFMA peak (TFLOP/s) and how many independent chains per SM sub-partition it
//      takes to reach it (=> DFMA latency / issue interval).
//   2. Shared-memory broadcast-load delivery rate for 8/16-byte loads (all lanes read
//      the same address): is a broadcast cheaper than a full-width load?
//   3. Mixed loop: one broadcast LDS feeding P DFMAs, P = 1,2,4,8 -- where is the knee?
//   4. DMMA m8n8k4 rate, alone and interleaved with DFMA (separate pipe or shared?).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/microbench.cu -o tools/microbench
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) {                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                 \
    }                                                                          \
  } while (0)

// ---------------------------------------------------------------- 1. DFMA peak
template <int CHAINS>
__global__ void dfma_kernel(double *out, int iters, double a, double b, long long *cyc) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) acc[c] = threadIdx.x * 1e-3 + c;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
#pragma unroll
      for (int c = 0; c < CHAINS; c++) acc[c] = fma(acc[c], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// ------------------------------------------- 1b. DFMA with the contraction's operand pattern
// acc[s] = fma(x[s][k], c[k], acc[s]): a FRESH multiplicand register per instruction, the coefficient register
// shared by the PPT instructions of one k (operand reuse), the accumulator -- everything in registers, no
// loads.  dfma_kernel above reads the same two registers (a, b) in every instruction; this one asks the
// register file for two fresh 64-bit operands per DFMA, as the Chebyshev contraction does.
template <int PPT, int NX, int NC>
__global__ void dfma_pattern_kernel(double *out, int iters, double a, long long *cyc) {
  double x[PPT][NX], c[NC], acc[PPT];
#pragma unroll
  for (int s = 0; s < PPT; s++) {
    acc[s] = threadIdx.x * 1e-3 + s;
#pragma unroll
    for (int k = 0; k < NX; k++) x[s][k] = 1.0 + 1e-9 * (threadIdx.x + 3 * k + 7 * s) * a;
  }
#pragma unroll
  for (int k = 0; k < NC; k++) c[k] = 1e-9 * (k + 1) * a;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < NX; k++) {
#pragma unroll
      for (int s = 0; s < PPT; s++) acc[s] = fma(x[s][k], c[k % NC], acc[s]);
    }
  }
  long long t1 = clock64();
  double r = 0;
#pragma unroll
  for (int s = 0; s < PPT; s++) r += acc[s];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// ------------------------------------------------------- 2/3. LDS broadcast mix
// Each iteration: one broadcast load of VEC doubles from shared memory (address is
// warp-uniform but data dependent on the loop counter so it cannot be hoisted),
// then P DFMAs per loaded double on independent accumulators.
template <int VEC, int P>
__global__ void lds_mix_kernel(double *out, int iters, const double *src, long long *cyc) {
  extern __shared__ __align__(16) double sm[];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = src[i];
  __syncthreads();
  double acc[(P > 0 ? P : 1) * VEC];
#pragma unroll
  for (int c = 0; c < (P > 0 ? P : 1) * VEC; c++) acc[c] = threadIdx.x * 1e-3 + c;
  double sink = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 32; u++) {
      const int idx = ((it * 32 + u) * VEC) & 2047;
      double c[VEC];
      if (VEC == 2) {
        double2 t = *reinterpret_cast<const double2 *>(&sm[idx]);
        c[0] = t.x;
        c[VEC - 1] = t.y;
      } else {
        c[0] = sm[idx];
      }
      if (P == 0) {
#pragma unroll
        for (int v = 0; v < VEC; v++) sink += c[v];  // 1 DADD per double, minimal
      } else {
#pragma unroll
        for (int v = 0; v < VEC; v++)
#pragma unroll
          for (int p = 0; p < P; p++) acc[v * P + p] = fma(acc[v * P + p], c[v], 1e-9);
      }
    }
  }
  long long t1 = clock64();
  double s = sink;
#pragma unroll
  for (int c = 0; c < (P > 0 ? P : 1) * VEC; c++) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// Same, but the per-lane operand comes from a NON-broadcast (lane-private) LDS.64.
template <int P>
__global__ void lds_private_kernel(double *out, int iters, const double *src, long long *cyc) {
  extern __shared__ __align__(16) double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = src[i & 2047];
  __syncthreads();
  double acc[P];
#pragma unroll
  for (int c = 0; c < P; c++) acc[c] = threadIdx.x * 1e-3 + c;
  const int lane_off = threadIdx.x & 255;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 32; u++) {
      const int idx = (((it * 32 + u) & 7) * 256 + lane_off);
      const double c = sm[idx];
#pragma unroll
      for (int p = 0; p < P; p++) acc[p] = fma(acc[p], c, 1e-9);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < P; c++) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// ---------------------------------------------------------------- 4. DMMA
template <int CHAINS, int DFMA_PER>
__global__ void dmma_kernel(double *out, int iters, double a, double b, long long *cyc) {
  double d0[CHAINS], d1[CHAINS];
  double f[DFMA_PER > 0 ? DFMA_PER : 1];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) {
    d0[c] = threadIdx.x * 1e-3;
    d1[c] = c;
  }
#pragma unroll
  for (int c = 0; c < (DFMA_PER > 0 ? DFMA_PER : 1); c++) f[c] = c + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int c = 0; c < CHAINS; c++) {
        asm volatile(
            "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
            : "+d"(d0[c]), "+d"(d1[c])
            : "d"(a), "d"(b));
#pragma unroll
        for (int p = 0; p < DFMA_PER; p++) f[p] = fma(f[p], a, b);
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += d0[c] + d1[c];
#pragma unroll
  for (int c = 0; c < (DFMA_PER > 0 ? DFMA_PER : 1); c++) s += f[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// ---------------------------------------------------------------- harness
struct Timer {
  cudaEvent_t a, b;
  Timer() {
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
  }
  void start() { CK(cudaEventRecord(a)); }
  float stop() {
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
  }
};

static double *d_out;
static long long *d_cyc;
static double *d_src;
static int n_sm;

template <int CHAINS>
void run_dfma(int threads, int ctas_per_sm) {
  const int iters = 4096;
  Timer t;
  dfma_kernel<CHAINS><<<n_sm * ctas_per_sm, threads>>>(d_out, 64, 1.0000001, 1e-9, d_cyc);
  CK(cudaDeviceSynchronize());
  t.start();
  dfma_kernel<CHAINS><<<n_sm * ctas_per_sm, threads>>>(d_out, iters, 1.0000001, 1e-9, d_cyc);
  float ms = t.stop();
  long long cyc;
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  double fma_per_thread = (double)iters * 16 * CHAINS;
  double total = fma_per_thread * threads * ctas_per_sm * n_sm;
  double per_sm_clk = fma_per_thread * threads * ctas_per_sm / (double)cyc;
  printf("dfma chains=%d threads=%d ctas/sm=%d warps/smsp=%.1f : %.2f TFLOP/s  %.1f DFMA/clk/SM  (%.0f MHz eff)\n",
         CHAINS, threads, ctas_per_sm, threads * ctas_per_sm / 128.0, 2 * total / ms * 1e-9,
         per_sm_clk, cyc / ms * 1e-3);
}

template <int PPT, int NX, int NC>
void run_dfma_pattern(int threads, int ctas_per_sm) {
  const int iters = 4096;
  Timer t;
  auto k = dfma_pattern_kernel<PPT, NX, NC>;
  k<<<n_sm * ctas_per_sm, threads>>>(d_out, 8, 1.0000001, d_cyc);
  CK(cudaDeviceSynchronize());
  t.start();
  k<<<n_sm * ctas_per_sm, threads>>>(d_out, iters, 1.0000001, d_cyc);
  float ms = t.stop();
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, k));
  double total = (double)iters * NX * PPT * threads * ctas_per_sm * n_sm;
  printf("dfma pattern: %d points x %d fresh multiplicands, %d coefficient regs, regs=%d threads=%d ctas/sm=%d warps/smsp=%.1f : %.2f TFLOP/s\n",
         PPT, NX, NC, fa.numRegs, threads, ctas_per_sm, threads * ctas_per_sm / 128.0, 2 * total / ms * 1e-9);
}

template <int VEC, int P>
void run_mix(int threads, int ctas_per_sm) {
  const int iters = 1024;
  auto k = lds_mix_kernel<VEC, P>;
  k<<<n_sm * ctas_per_sm, threads, 2048 * 8>>>(d_out, 8, d_src, d_cyc);
  CK(cudaDeviceSynchronize());
  Timer t;
  t.start();
  k<<<n_sm * ctas_per_sm, threads, 2048 * 8>>>(d_out, iters, d_src, d_cyc);
  float ms = t.stop();
  long long cyc;
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  double loads = (double)iters * 32;  // per thread
  double warps = threads / 32.0 * ctas_per_sm;
  double lds_bytes_per_clk = loads * warps * 32 * VEC * 8 / cyc;  // delivered bytes
  double dfma_per_clk = loads * VEC * P * threads * ctas_per_sm / cyc;
  printf("lds_bcast vec=%dB P=%d threads=%d ctas/sm=%d : %.1f warp-LDS/kclk/SM  %.0f delivered B/clk/SM  %.1f DFMA/clk/SM  %.3f ms\n",
         VEC * 8, P, threads, ctas_per_sm, loads * warps / cyc * 1000, lds_bytes_per_clk,
         dfma_per_clk, ms);
}

template <int P>
void run_private(int threads, int ctas_per_sm) {
  const int iters = 1024;
  auto k = lds_private_kernel<P>;
  k<<<n_sm * ctas_per_sm, threads, 4096 * 8>>>(d_out, 8, d_src, d_cyc);
  CK(cudaDeviceSynchronize());
  Timer t;
  t.start();
  k<<<n_sm * ctas_per_sm, threads, 4096 * 8>>>(d_out, iters, d_src, d_cyc);
  float ms = t.stop();
  long long cyc;
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  double loads = (double)iters * 32;
  double warps = threads / 32.0 * ctas_per_sm;
  printf("lds_private P=%d threads=%d ctas/sm=%d : %.1f warp-LDS/kclk/SM  %.0f B/clk/SM  %.1f DFMA/clk/SM\n",
         P, threads, ctas_per_sm, loads * warps / cyc * 1000, loads * warps * 256 / cyc,
         loads * P * threads * ctas_per_sm / cyc);
}

template <int CHAINS, int DFMA_PER>
void run_dmma(int threads, int ctas_per_sm) {
  const int iters = 2048;
  auto k = dmma_kernel<CHAINS, DFMA_PER>;
  k<<<n_sm * ctas_per_sm, threads>>>(d_out, 8, 1.0000001, 1e-9, d_cyc);
  CK(cudaDeviceSynchronize());
  Timer t;
  t.start();
  k<<<n_sm * ctas_per_sm, threads>>>(d_out, iters, 1.0000001, 1e-9, d_cyc);
  float ms = t.stop();
  long long cyc;
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  double mma_per_warp = (double)iters * 8 * CHAINS;
  double warps = threads / 32.0 * ctas_per_sm;
  double mma_fma_clk = mma_per_warp * warps * 256 / cyc;
  double dfma_clk = mma_per_warp * DFMA_PER * threads * ctas_per_sm / cyc;
  printf("dmma chains=%d +dfma/mma=%d threads=%d ctas/sm=%d : DMMA %.1f FMA/clk/SM (%.2f TF)  DFMA %.1f /clk/SM  total %.2f TFLOP/s\n",
         CHAINS, DFMA_PER, threads, ctas_per_sm, mma_fma_clk,
         2 * mma_per_warp * warps * 256 * n_sm / ms * 1e-9, dfma_clk,
         2 * (mma_per_warp * warps * 256 + mma_per_warp * DFMA_PER * threads * ctas_per_sm) * n_sm / ms * 1e-9);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  n_sm = prop.multiProcessorCount;
  printf("device %s  SMs %d  clockRate %d kHz  smem/SM %zu\n", prop.name, n_sm, prop.clockRate,
         prop.sharedMemPerMultiprocessor);
  CK(cudaMalloc(&d_out, sizeof(double) * 1024 * 64 * n_sm));
  CK(cudaMalloc(&d_cyc, 8));
  std::vector<double> h(2048);
  for (int i = 0; i < 2048; i++) h[i] = 1.0 + 1e-9 * i;
  CK(cudaMalloc(&d_src, 2048 * 8));
  CK(cudaMemcpy(d_src, h.data(), 2048 * 8, cudaMemcpyHostToDevice));

  printf("--- 1. DFMA peak / latency\n");
  run_dfma<8>(1024, 2);
  printf("--- 1b. DFMA with fresh register operands (the contraction's pattern), no loads\n");
  run_dfma_pattern<2, 14, 14>(128, 2);
  run_dfma_pattern<2, 14, 14>(256, 1);
  run_dfma_pattern<2, 14, 14>(256, 4);
  run_dfma_pattern<4, 14, 14>(128, 2);
  run_dfma_pattern<4, 14, 14>(256, 4);
  run_dfma_pattern<4, 14, 2>(256, 4);
  run_dfma_pattern<8, 14, 14>(256, 2);
  run_dfma_pattern<8, 2, 2>(256, 4);
  printf("--- 1. (continued)\n");
  run_dfma<8>(256, 4);
  run_dfma<4>(256, 4);
  run_dfma<2>(256, 4);
  run_dfma<1>(256, 4);
  run_dfma<1>(128, 1);  // 1 warp per SMSP, 1 chain  => latency bound: DFMA/clk/SM = 128/lat
  run_dfma<2>(128, 1);
  run_dfma<4>(128, 1);
  run_dfma<8>(128, 1);
  run_dfma<6>(256, 1);
  run_dfma<6>(256, 2);

  printf("--- 2. LDS broadcast only (P=0: one DADD per loaded double)\n");
  run_mix<1, 0>(256, 4);
  run_mix<2, 0>(256, 4);
  printf("--- 3. LDS broadcast + P DFMA per loaded double\n");
  run_mix<1, 1>(256, 4);
  run_mix<1, 2>(256, 4);
  run_mix<1, 3>(256, 4);
  run_mix<1, 4>(256, 4);
  run_mix<1, 6>(256, 4);
  run_mix<1, 8>(256, 4);
  run_mix<2, 1>(256, 4);
  run_mix<2, 2>(256, 4);
  run_mix<2, 3>(256, 4);
  run_mix<2, 4>(256, 4);
  run_mix<2, 6>(256, 2);
  run_mix<2, 8>(256, 2);
  run_mix<1, 4>(256, 1);
  run_mix<2, 4>(256, 1);
  printf("--- 3b. lane-private LDS.64 + P DFMA\n");
  run_private<1>(256, 4);
  run_private<2>(256, 4);
  run_private<4>(256, 4);
  run_private<8>(256, 4);
  printf("--- 4. DMMA m8n8k4\n");
  run_dmma<1, 0>(256, 4);
  run_dmma<2, 0>(256, 4);
  run_dmma<4, 0>(256, 4);
  run_dmma<4, 0>(1024, 2);
  run_dmma<2, 4>(256, 4);
  run_dmma<2, 8>(256, 4);
  run_dmma<2, 16>(256, 4);
  run_dmma<4, 8>(256, 4);
  return 0;
}

// FP64 FMA peak of the device: the denominator of the evaluation kernel's roofline
// (MEASURED_PEAKS.json carries HBM and bf16 only).  Eight independent DFMA chains per
// thread, 256 threads, 4 CTAs per SM; best of `reps`.
#include "common.cuh"

namespace tb {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
  double acc[8];
#pragma unroll
  for (int c = 0; c < 8; c++) acc[c] = threadIdx.x * 1e-3 + c;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
#pragma unroll
      for (int c = 0; c < 8; c++) acc[c] = fma(acc[c], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) s += acc[c];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int run_fp64_peak(tbslas_ctx *ctx, int reps, double *tflops) {
  const int iters = 2048, ctas = ctx->n_sm * 4;
  void *buf;
  TB_TRY(ws_get(ctx, WS_MISC, sizeof(double) * 256 * (size_t)ctas, &buf));
  cudaEvent_t e0, e1;
  TB_CUDA(ctx, cudaEventCreate(&e0));
  TB_CUDA(ctx, cudaEventCreate(&e1));
  double best = 0;
  for (int r = 0; r < reps + 1; r++) {
    TB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    dfma_peak_kernel<<<ctas, 256, 0, ctx->stream>>>((double *)buf, iters, 1.0000001, 1e-9);
    ctx->launches++;
    TB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    TB_CUDA(ctx, cudaEventSynchronize(e1));
    float ms;
    TB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * iters * 16 * 8 * 256.0 * ctas / (ms * 1e-3) * 1e-12;
    if (r > 0 && tf > best) best = tf;  // r == 0 is the warm-up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return TBSLAS_OK;
}

}  // namespace tb

// Persistent, warp-pipelined variant of the Chebyshev evaluation kernel (degrees with a
// fully unrolled contraction, q <= 14, and coefficient blocks small enough for one
// buffer per warp).
//
// Same arithmetic as cheb_eval.cuh (ZLevel/RowPair, cheb_basis); what changes is the
// scheduling around the contraction, which the ncu profile of the one-tile-per-CTA kernel
// showed to cost ~8 % of the FP64 pipe (dependent gather perm -> pos -> bases at the head
// of every CTA, hidden only by the one other CTA on the SM):
//   * grid = 2 CTAs per SM, every WARP is an independent worker that walks a contiguous
//     range of warp-tiles (32*PPT points of one leaf each; consecutive tiles mostly share
//     the leaf);
//   * the leaf's coefficient block lives in a per-warp shared-memory buffer, refilled by
//     one bulk-TMA copy (mbarrier complete_tx) only when the leaf changes;
//   * the gathers run two tiles ahead through cp.async: while tile t is contracted, the
//     coordinates of tile t+1 (addresses from the already-landed slice of perm) and the
//     perm slice of tile t+2 stream into shared memory, so no global-load latency is ever
//     exposed in front of the FP64 work;
//   * no tile map: a worker finds its first leaf by binary search in tile_start and walks
//     forward.
#pragma once
#include "cheb_eval.cuh"

namespace tb {

__device__ __forceinline__ void cp_async_8(void *dst_smem, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16(void *dst_smem, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void *dst_smem, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// per-worker staging area, in doubles: [3][TP] coordinates (SoA), [4] geom, 2 x [3][TP]
// RK2 base positions (AXPY epilogue only), 3 x [TP] u32 perm slices (ring), 16 ints of
// tile bookkeeping (kept in shared memory so that nothing but the bases and accumulators
// is live in registers across the contraction)
template <int PPT, int EPI>
struct WarpStage {
  static constexpr int TP = 32 * PPT;
  // RK2 base: 2 x [3][TP] coordinates (EPI_AXPY), or 2 x [TP] leaf geometries of 4 doubles (EPI_AXPY_GRID)
  static constexpr int kBase = (EPI == EPI_AXPY) ? 6 * TP : (EPI == EPI_AXPY_GRID ? 8 * TP : 0);
  static constexpr int kPermDoubles = (3 * TP + 1) / 2;
  static constexpr int kDoubles = 3 * TP + 4 + kBase + kPermDoubles + (kPermDoubles & 1) + 8;
};
// bookkeeping words
enum {
  CT_WL = 0, CT_TS, CT_TE, CT_BS, CT_BE,  // walker: current leaf, its tile and slot ranges
  CT_RING = 5,                              // 3 x {leaf, slot0, cnt}; cnt == 0: end of stream
  CT_N = 14,                                // stream index of the tile being contracted
  CT_NEXT = 15,                             // next tile of the current chunk ...
};
// ... and its end live in the two ints that follow the 16 (the block is 8 doubles + 1)

constexpr int kWtThreads = 32;  // one warp per CTA: every address below is CTA-uniform
#ifndef TB_WT_PAIR
#define TB_WT_PAIR 0  // 1: rows (i,j) and (i,j+1) contracted together (4 instead of 2 DFMA chains per warp)
#endif
#ifndef TB_WT_LDS64
#define TB_WT_LDS64 0  // 1: coefficients by broadcast LDS.64 instead of LDS.128
#endif
#ifndef TB_WT_MINB
#define TB_WT_MINB 8
#endif

// Work distribution: tiles are handed out in chunks of at most `chunk` consecutive tiles from a
// global tile counter (zeroed by the locate launcher), so a warp the scheduler favours simply
// takes more chunks and all workers finish together.
template <int Q, int PPT, int EPI>
__global__ void __launch_bounds__(kWtThreads, TB_WT_MINB)
cheb_eval_wt_kernel(const EvalParams p, unsigned *__restrict__ chunk_counter, unsigned chunk, const GridBase gb) {
  constexpr int D = Q + 1;
  constexpr int TP = 32 * PPT;
  typedef WarpStage<PPT, EPI> Stage;
  extern __shared__ __align__(128) double s_all[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_cend;
  __shared__ double s_node[EPI == EPI_AXPY_GRID ? D : 1];  // the grid's 1-D nodes (grid-base epilogue)
  double *const s_coef = s_all;                       // [stride]
  double *const s_x = s_all + p.stride;               // [3][TP]
  double *const s_g = s_x + 3 * TP;                   // [4]
  double *const s_b = s_g + 4;                        // [2][3][TP] (AXPY only)
  uint32_t *const s_perm = reinterpret_cast<uint32_t *>(s_b + Stage::kBase);  // [3][TP]
  int *const s_ctl = reinterpret_cast<int *>(s_x + Stage::kDoubles - 8);       // [16]
  const int lane = threadIdx.x;
  if (lane == 0) {
    mbar_init(&s_bar, 1);
    s_ctl[CT_N] = 0;
    s_ctl[CT_NEXT] = 0;
    s_cend = 0;
  }
  if (EPI == EPI_AXPY_GRID && lane < D) s_node[lane] = gb.node[lane];
  __syncwarp();
  // the descriptor of the next tile of the stream goes into ring slot n % 3 (written by lane 0)
  auto describe_next = [&](unsigned n) {
    const unsigned n_tiles = __ldg(p.tile_start + p.n_bins);
    // lane 0 decides whether a new chunk is due and takes it; the whole warp then finds the chunk's
    // first leaf TOGETHER: a 32-ary search over tile_start (lane i probes the i-th of 32 evenly spaced
    // bins, one ballot per round) needs ceil(log32(n_bins)) dependent loads -- 4 for 81 k leaves --
    // where a one-lane binary search needs 17, with the other 31 lanes waiting for it.
    unsigned t = 0;
    int take = 0;  // 0: next tile of the current chunk, 1: new chunk at tile t, 2: end of stream
    if (lane == 0) {
      t = (unsigned)s_ctl[CT_NEXT];
      if (t >= (unsigned)s_cend) {
        // guided self-scheduling: a share of what is left, so the chunks shrink towards the
        // end of the launch and the workers finish within a couple of tiles of each other
        const unsigned done = *reinterpret_cast<volatile unsigned *>(chunk_counter);
        const unsigned rem = n_tiles > done ? n_tiles - done : 0u;
        unsigned sz = rem / (2u * gridDim.x);
        sz = sz < 2u ? 2u : (sz > chunk ? chunk : sz);
        t = atomicAdd(chunk_counter, sz);
        if (t >= n_tiles) {
          take = 2;
          s_ctl[CT_NEXT] = (int)n_tiles;
          s_cend = (int)n_tiles;
        } else {
          take = 1;
          s_cend = (int)min(t + sz, n_tiles);
        }
      }
    }
    take = __shfl_sync(0xffffffffu, take, 0);
    if (take == 1) {
      t = __shfl_sync(0xffffffffu, t, 0);
      // leaf of tile t = last bin j with tile_start[j] <= t  (tile_start[n_bins] = n_tiles > t)
      int lo = 0, hi = p.n_bins;
      while (hi - lo > 1) {
        const int step = (hi - lo + 31) >> 5;
        const int pos = lo + (lane + 1) * step;
        const bool le = pos < hi && __ldg(p.tile_start + pos) <= t;  // true on a prefix of the lanes
        const int cnt = __popc(__ballot_sync(0xffffffffu, le));
        const int nlo = lo + cnt * step;
        hi = min(nlo + step, hi);
        lo = nlo;
      }
      if (lane == 0) {
        s_ctl[CT_WL] = lo;
        s_ctl[CT_TS] = (int)__ldg(p.tile_start + lo);
        s_ctl[CT_TE] = (int)__ldg(p.tile_start + lo + 1);
        s_ctl[CT_BS] = (int)__ldg(p.bin_start + lo);
        s_ctl[CT_BE] = (int)__ldg(p.bin_start + lo + 1);
      }
    }
    if (lane == 0) {
      int *r = s_ctl + CT_RING + 3 * (n % 3);
      if (take != 2) {
        int wl = s_ctl[CT_WL];
        unsigned ts = (unsigned)s_ctl[CT_TS], te = (unsigned)s_ctl[CT_TE];
        unsigned bs = (unsigned)s_ctl[CT_BS], be = (unsigned)s_ctl[CT_BE];
        if (t >= te) {
          do {
            wl++;
            ts = te;
            te = __ldg(p.tile_start + wl + 1);
            bs = be;
            be = __ldg(p.bin_start + wl + 1);
          } while (t >= te);
          s_ctl[CT_WL] = wl;
          s_ctl[CT_TS] = (int)ts;
          s_ctl[CT_TE] = (int)te;
          s_ctl[CT_BS] = (int)bs;
          s_ctl[CT_BE] = (int)be;
        }
        const unsigned slot0 = bs + (t - ts) * TP;
        r[0] = wl;
        r[1] = (int)slot0;
        r[2] = (int)min((unsigned)TP, be - slot0);
        s_ctl[CT_NEXT] = (int)(t + 1);
      } else {
        r[0] = 0;
        r[1] = 0;
        r[2] = 0;
      }
    }
    __syncwarp();
  };
  auto ring_cnt = [&](unsigned n) { return (unsigned)s_ctl[CT_RING + 3 * (n % 3) + 2]; };
  auto fetch_perm = [&](unsigned n) {  // perm slice of stream tile n -> ring slot n % 3
    const int *r = s_ctl + CT_RING + 3 * (n % 3);
    const unsigned slot0 = (unsigned)r[1], cnt = (unsigned)r[2];
#pragma unroll
    for (int s = 0; s < PPT; s++) {
      const unsigned o = s * 32 + lane;
      cp_async_4(s_perm + (n % 3) * TP + o, p.perm + slot0 + (o < cnt ? o : 0u));
    }
  };
  auto fetch_points = [&](unsigned n) {  // coordinates (+ RK2 base) + geom of tile n -> staging area
    const int leaf = s_ctl[CT_RING + 3 * (n % 3)];
#pragma unroll
    for (int s = 0; s < PPT; s++) {
      const unsigned o = s * 32 + lane;
      const size_t i = s_perm[(n % 3) * TP + o];
      const double *x = p.pos + 3 * i;
      cp_async_8(s_x + o, x);
      cp_async_8(s_x + TP + o, x + 1);
      cp_async_8(s_x + 2 * TP + o, x + 2);
      if (EPI == EPI_AXPY) {
        const double *bp = p.base + 3 * i;
        double *d = s_b + (n & 1) * 3 * TP + o;
        cp_async_8(d, bp);
        cp_async_8(d + TP, bp + 1);
        cp_async_8(d + 2 * TP, bp + 2);
      }
      if (EPI == EPI_AXPY_GRID) {  // geometry of the grid leaf the point is a node of
        const double4 *gp = gb.ggeom + (unsigned)i / (unsigned)(D * D * D);  // (the launcher checked gb.D == D)
        double *d = s_b + ((n & 1) * TP + o) * 4;
        cp_async_16(d, gp);
        cp_async_16(d + 2, reinterpret_cast<const double *>(gp) + 2);
      }
    }
    if (lane < 4) cp_async_8(s_g + lane, reinterpret_cast<const double *>(p.geom + leaf) + lane);
  };

  // ---- prologue: perm(0) -> wait -> points(0) + perm(1) in flight -------------------------
  describe_next(0);
  if (ring_cnt(0) == 0) return;
  fetch_perm(0);
  cp_async_wait_all();
  __syncwarp();
  fetch_points(0);
  describe_next(1);
  if (ring_cnt(1)) fetch_perm(1);
  int cur_leaf = -1;
  uint32_t phase = 0;

#pragma unroll 1
  for (;;) {
    cp_async_wait_all();  // points(n), geom(n) and perm(n+1) have landed
    __syncwarp();
    bool inside[PPT];  // a point outside its leaf in any axis evaluates to 0 (all-zero bases: cheb_poly)
    double px[PPT][D], py[PPT][D], zc[PPT];
    {
      const unsigned n = (unsigned)s_ctl[CT_N];
      // ---- bases of tile n ------------------------------------------------------------
      const double gx = s_g[0], gy = s_g[1], gz = s_g[2], gw = s_g[3];
#pragma unroll
      for (int s = 0; s < PPT; s++) {
        const unsigned o = s * 32 + lane;
        // xi = (x - c) * 2 * 2^depth - 1, left to right (tree_functor.h:288-293)
        const double xi = __dadd_rn(__dmul_rn(__dsub_rn(s_x[o], gx), gw), -1.0);
        const double yi = __dadd_rn(__dmul_rn(__dsub_rn(s_x[TP + o], gy), gw), -1.0);
        const double zi = __dadd_rn(__dmul_rn(__dsub_rn(s_x[2 * TP + o], gz), gw), -1.0);
        const bool inx = cheb_basis<Q>(xi, px[s]);
        const bool iny = cheb_basis<Q>(yi, py[s]);
        const bool inz = fabs(zi) <= 1.0;
        zc[s] = inz ? zi : 0.0;
        inside[s] = inx && iny && inz;
      }
      __syncwarp();  // every lane has consumed the staging area
      // ---- keep two tiles in flight ----------------------------------------------------
      if (ring_cnt(n + 1)) {
        fetch_points(n + 1);
        describe_next(n + 2);
        if (ring_cnt(n + 2)) fetch_perm(n + 2);
      }
      // ---- coefficients of this leaf (only when the leaf changed) ------------------------
      const int leaf0 = s_ctl[CT_RING + 3 * (n % 3)];
      if (leaf0 != cur_leaf) {
        cur_leaf = leaf0;
        if (lane == 0) {
          fence_proxy_async();  // earlier generic-proxy reads of the buffer vs. the async write
          const uint32_t bytes = p.stride * 8u;
          mbar_expect_tx(&s_bar, bytes);
          tma_bulk_g2s(s_coef, p.coeff + (size_t)leaf0 * p.stride, bytes, &s_bar);
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1u;
      }
    }
    // ---- contraction ------------------------------------------------------------------
    bool last = false;
#pragma unroll 1
    for (int l = 0; l < p.dof; l++) {
#if TB_WT_LDS64
      const double *C2 = s_coef + l * p.ncoef_pad;  // 8-byte alignment only: broadcast LDS.64
#else
      const double2 *C2 = reinterpret_cast<const double2 *>(s_coef + l * p.ncoef_pad);
#endif
      double u[PPT], tz0[PPT], tz1[PPT];
#pragma unroll
      for (int s = 0; s < PPT; s++) u[s] = tz0[s] = tz1[s] = 0.0;
      ZLevel<Q, PPT, false, TB_WT_PAIR != 0, 0, 0>::run(C2, px, py, nullptr, zc, tz0, tz1, u);
      const unsigned n = (unsigned)s_ctl[CT_N];
      const unsigned cnt0 = ring_cnt(n);
#pragma unroll
      for (int s = 0; s < PPT; s++) {
        if (!inside[s]) u[s] = 0.0;
        const unsigned o = s * 32 + lane;
        if (o < cnt0) {
          const size_t i = s_perm[(n % 3) * TP + o];
          if (EPI == EPI_STORE) {
            p.out[i * p.dof + l] = u[s];
          } else if (EPI == EPI_AXPY) {  // x' = x0 + alpha * v, multiply then add as traj.inc:36,42
            p.out[3 * i + l] = __dadd_rn(s_b[(n & 1) * 3 * TP + l * TP + o], __dmul_rn(p.alpha, u[s]));
          } else {  // the same with x0 rebuilt from its leaf's geometry and its node index
            const double4 g = *reinterpret_cast<const double4 *>(s_b + ((n & 1) * TP + o) * 4);
            p.out[3 * i + l] = __dadd_rn(grid_base_coord_ct<D>(gb.periodic, s_node, g, (unsigned)i, l), __dmul_rn(p.alpha, u[s]));
          }
        }
      }
      if (l + 1 == p.dof) {  // advance the stream
        __syncwarp();        // all lanes are done with the coefficient buffer and the ring
        last = ring_cnt(n + 1) == 0;
        __syncwarp();
        if (lane == 0) s_ctl[CT_N] = (int)(n + 1);
      }
    }
    if (last) break;
  }
}

template <int Q, int PPT, int EPI>
size_t eval_wt_smem_bytes(const tbslas_tree *t) {
  return (t->stride + (size_t)WarpStage<PPT, EPI>::kDoubles) * sizeof(double);
}

template <int Q, int PPT>
int launch_cheb_eval_wt(tbslas_ctx *ctx, const EvalArgs &a) {
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
  p.tile_map = nullptr;
  p.out = a.out;
  p.base = a.base;
  p.alpha = a.alpha;
  GridBase gb;
  if (a.epilogue == EPI_AXPY_GRID) {
    if (!a.grid) return fail(ctx, TBSLAS_ERR_INVALID, "grid epilogue without a grid");
    if ((int)a.grid->D != Q + 1) return fail(ctx, TBSLAS_ERR_UNSUPPORTED, "grid epilogue: the grid's degree differs from the tree's");
    gb = *a.grid;
  }
  if (a.max_tiles == 0) return TBSLAS_OK;
  // every CTA (one warp) is a worker: no more workers than tiles, and as many per SM as are resident
  // together -- 8 up to 27 KB of shared memory per worker, fewer for larger coefficient blocks
  auto launch = [&](auto k, size_t smem) -> int {
    if (smem > 48 * 1024)
      TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 8;
    if (smem > 27 * 1024) {
      TB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kWtThreads, smem));
      if (per_sm < 1) return fail(ctx, TBSLAS_ERR_UNSUPPORTED, "evaluation kernel does not fit an SM (%zu bytes)", smem);
      if (per_sm > 8) per_sm = 8;
    }
    size_t grid = a.max_tiles;
    if (grid > (size_t)per_sm * ctx->n_sm) grid = (size_t)per_sm * ctx->n_sm;
    // largest chunk of consecutive tiles (leaf locality); the kernel shrinks them towards the end
    // (an eighth of a worker's tiles -- or half of them when it has fewer than 32: every refill is a global atomic
    // plus a search for the chunk's first leaf, which a small launch pays many times over.  C1, 20 tiles per
    // worker: evaluation launches 0.449 -> 0.421 ms per step; half everywhere cost C2 0.6 ms in its launch over
    // the tensor-grid exceptions, 57 tiles per worker)
    const size_t per_worker = a.max_tiles / grid;
    size_t chunk = per_worker / (per_worker < 32 ? 2 : 8);
    chunk = chunk < 2 ? 2 : (chunk > 16 ? 16 : chunk);
    k<<<(unsigned)grid, kWtThreads, smem, ctx->stream>>>(p, a.chunk_counter, (unsigned)chunk, gb);
    TB_CUDA(ctx, cudaGetLastError());
    return TBSLAS_OK;
  };
  if (a.epilogue == EPI_STORE) return launch(cheb_eval_wt_kernel<Q, PPT, EPI_STORE>, eval_wt_smem_bytes<Q, PPT, EPI_STORE>(t));
  if (a.epilogue == EPI_AXPY) return launch(cheb_eval_wt_kernel<Q, PPT, EPI_AXPY>, eval_wt_smem_bytes<Q, PPT, EPI_AXPY>(t));
  return launch(cheb_eval_wt_kernel<Q, PPT, EPI_AXPY_GRID>, eval_wt_smem_bytes<Q, PPT, EPI_AXPY_GRID>(t));
}

}  // namespace tb

// Point location in the Morton-ordered leaf list, and grouping of points by leaf.
//
// Replaces, for one rank (reference src/tree/tree_functor.h):
//   :442-449  periodic wrap (in place)
//   :464-483  Morton key of every point  [pvfmm::MortonId, restated: SURVEY.md App. A]
//   :487-489  omp_par::merge_sort of (key, index) pairs
//   :190-198  part_indx[j] = lower_bound(sorted keys, key(leaf j))
//   :491-513  owner split against the ranks' first-leaf keys
// The reference sorts all points by a 45-bit key three times per call only to group
// them by leaf; here each point finds its leaf directly and a counting sort on the
// leaf id (histogram by warp-aggregated atomics -> scan -> scatter) builds the
// grouping.  The assignment rule is unchanged: point p belongs to the last leaf j with
// key(leaf j) <= key(p); the last leaf takes every larger key.
#include "common.cuh"
#include "keys.cuh"

namespace tb {

constexpr int kLocateThreads = 256;
constexpr int kCoopIters = 4;  // cooperative searches per warp before per-lane fallback

// Cooperative 32-ary search by one warp: number of keys <= k (k warp-uniform).
// Each level, lane t probes keys[lo + t*step]; the ballot of "probe <= k" is a
// prefix of ones because the keys ascend, so its popcount selects the sub-range.
__device__ __forceinline__ int warp_count_le(const uint64_t *__restrict__ keys, int n, uint64_t k,
                                             int lane) {
  int lo = 0, hi = n;  // answer in [lo, hi]
  while (hi > lo) {
    const int step = (hi - lo + 31) >> 5;
    const int idx = lo + lane * step;
    const bool le = (idx < hi) && (__ldg(keys + idx) <= k);
    const int c = __popc(__ballot_sync(0xffffffffu, le));
    if (c == 0) return lo;
    const int nlo = lo + (c - 1) * step + 1;
    hi = min(hi, lo + c * step);
    lo = nlo;
  }
  return lo;
}

__device__ __forceinline__ int lane_count_le(const uint64_t *__restrict__ keys, int n, uint64_t k) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) <= k)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

template <bool MULTI>
__global__ void __launch_bounds__(kLocateThreads)
locate_kernel(const uint64_t *__restrict__ keys, int n_leaf, int periodic, double *__restrict__ pos,
              size_t n, int32_t *__restrict__ leaf_out, uint32_t *__restrict__ rank_out,
              uint32_t *__restrict__ count, const uint64_t *__restrict__ splitters, int nranks,
              int myrank, uint32_t *__restrict__ send_count) {
  const size_t i = (size_t)blockIdx.x * kLocateThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = i < n;
  uint64_t key = 0;
  if (valid) {
    double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
    if (periodic) {
      const double x0 = x, y0 = y, z0 = z;
      x = wrap_periodic(x);
      y = wrap_periodic(y);
      z = wrap_periodic(z);
      if (x != x0) pos[3 * i] = x;  // the reference rewrites the caller's buffer
      if (y != y0) pos[3 * i + 1] = y;
      if (z != z0) pos[3 * i + 2] = z;
    }
    key = point_key(x, y, z, periodic);
  }
  int my_leaf = 0;
  uint32_t my_rank = 0;
  bool todo = valid;

  if (MULTI && valid) {
    // owner = last rank whose first-leaf key is <= key (rank 0 below the first)
    uint64_t lo_key = __ldg(splitters + myrank);
    bool mine = key >= lo_key || myrank == 0;
    if (myrank + 1 < nranks) mine = mine && key < __ldg(splitters + myrank + 1);
    if (!mine) {
      int owner = 0;
      for (int r = 1; r < nranks; r++)
        if (__ldg(splitters + r) <= key) owner = r;
      todo = false;
      const unsigned peers = __match_any_sync(__activemask(), owner);
      const int leader = __ffs(peers) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(send_count + owner, (uint32_t)__popc(peers));
      base = __shfl_sync(peers, base, leader);
      my_rank = base + __popc(peers & ((1u << lane) - 1));
      my_leaf = -2 - owner;
    }
  }

  // Warp-cooperative phase: departure points are spatially coherent, so most lanes of
  // a warp share the leaf of the first unresolved lane.
  unsigned pending = __ballot_sync(0xffffffffu, todo);
  for (int it = 0; it < kCoopIters && pending; it++) {
    const int leader = __ffs(pending) - 1;
    const uint64_t lk = __shfl_sync(0xffffffffu, key, leader);
    const int j = warp_count_le(keys, n_leaf, lk, lane) - 1;  // warp-uniform
    const uint64_t klo = (j >= 0) ? __ldg(keys + j) : 0ull;
    const bool last = (j + 1 >= n_leaf);
    const uint64_t khi = last ? ~0ull : __ldg(keys + j + 1);
    const bool hit = ((pending >> lane) & 1u) && key >= klo && (last || key < khi);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    const int bin = (j >= 0) ? j : n_leaf;  // bin n_leaf = "no leaf" (evaluates to 0)
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(count + bin, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (hit) {
      my_leaf = bin;
      my_rank = base + __popc(m & ((1u << lane) - 1));
    }
    pending &= ~m;
    if (__popc(m) < 4) break;  // incoherent input: stop paying for warp-wide searches
  }
  if ((pending >> lane) & 1u) {  // per-lane fallback, atomics aggregated per leaf
    const int j = lane_count_le(keys, n_leaf, key) - 1;
    const int bin = (j >= 0) ? j : n_leaf;
    const unsigned peers = __match_any_sync(pending, bin);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(count + bin, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    my_leaf = bin;
    my_rank = base + __popc(peers & ((1u << lane) - 1));
  }
  if (valid) {
    leaf_out[i] = my_leaf;
    rank_out[i] = my_rank;
  }
}

int launch_locate(tbslas_ctx *ctx, const LocateArgs &a) {
  const tbslas_tree *t = a.tree;
  StageScope sc(ctx, ST_LOCATE, (double)a.n, 1);
  TB_CUDA(ctx, cudaMemsetAsync(a.count, 0, sizeof(uint32_t) * (t->n_leaf + 2), ctx->stream));
  if (a.n == 0) return TBSLAS_OK;
  const unsigned grid = (unsigned)((a.n + kLocateThreads - 1) / kLocateThreads);
  if (ctx->nranks > 1 && a.send_count) {
    TB_CUDA(ctx, cudaMemsetAsync(a.send_count, 0, sizeof(uint32_t) * ctx->nranks, ctx->stream));
    locate_kernel<true><<<grid, kLocateThreads, 0, ctx->stream>>>(
        t->d_key, (int)t->n_leaf, a.periodic, a.pos, a.n, a.leaf, a.rank, a.count, t->d_splitters,
        ctx->nranks, ctx->rank, a.send_count);
  } else {
    locate_kernel<false><<<grid, kLocateThreads, 0, ctx->stream>>>(
        t->d_key, (int)t->n_leaf, a.periodic, a.pos, a.n, a.leaf, a.rank, a.count, nullptr, 1, 0,
        nullptr);
  }
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

// ---------------------------------------------------------------------------
// Exclusive scan of the bin counts (points and evaluation tiles at once) by one CTA,
// and the tile -> (leaf, first slot) map the evaluation grid indexes.
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;

__global__ void __launch_bounds__(kScanThreads)
scan_bins_kernel(const uint32_t *__restrict__ count, int n_bins, int tile_pts,
                 uint32_t *__restrict__ bin_start, uint32_t *__restrict__ tile_start,
                 int2 *__restrict__ tile_map, unsigned max_tiles) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_bins; base += kScanThreads * kScanItems) {
    // packed value: low 32 bits = points, high 32 bits = tiles
    unsigned long long v[kScanItems], sum = 0;
    const int first = base + threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      const int j = first + k;
      const unsigned c = (j < n_bins) ? count[j] : 0u;
      v[k] = (unsigned long long)c | ((unsigned long long)((c + tile_pts - 1) / tile_pts) << 32);
      sum += v[k];
    }
    unsigned long long incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = warp_tot[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += o;
      }
      warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const unsigned long long carry = carry_s;
    unsigned long long run = carry + (warp ? warp_tot[warp - 1] : 0ull) + (incl - sum);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      const int j = first + k;
      if (j < n_bins) {
        const unsigned ps = (unsigned)(run & 0xffffffffu), ts = (unsigned)(run >> 32);
        bin_start[j] = ps;
        tile_start[j] = ts;
        const unsigned nt = (unsigned)(v[k] >> 32);
        for (unsigned c = 0; c < nt; c++)
          if (ts + c < max_tiles) tile_map[ts + c] = make_int2(j, (int)(ps + c * tile_pts));
      }
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == kScanThreads - 1) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    bin_start[n_bins] = (unsigned)(carry_s & 0xffffffffu);
    tile_start[n_bins] = (unsigned)(carry_s >> 32);
  }
}

__global__ void scatter_perm_kernel(const int32_t *__restrict__ leaf, const uint32_t *__restrict__ rank,
                                    const uint32_t *__restrict__ bin_start, size_t n,
                                    uint32_t *__restrict__ perm) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = leaf[i];
  if (j < 0) return;  // outsider: travels to its owner instead
  perm[__ldg(bin_start + j) + rank[i]] = (uint32_t)i;
}

int launch_bin(tbslas_ctx *ctx, const BinArgs &a) {
  StageScope sc(ctx, ST_BIN, (double)a.n, 2);
  const int n_bins = (int)a.n_leaf + 1;  // + null leaf
  scan_bins_kernel<<<1, kScanThreads, 0, ctx->stream>>>(a.count, n_bins, a.tile_pts, a.bin_start,
                                                        a.tile_start, a.tile_map,
                                                        (unsigned)a.max_tiles);
  TB_CUDA(ctx, cudaGetLastError());
  if (a.n) {
    const unsigned grid = (unsigned)((a.n + 255) / 256);
    scatter_perm_kernel<<<grid, 256, 0, ctx->stream>>>(a.leaf, a.rank, a.bin_start, a.n, a.perm);
    TB_CUDA(ctx, cudaGetLastError());
  }
  return TBSLAS_OK;
}

}  // namespace tb

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
#include <cstdlib>

#include "common.cuh"
#include "keys.cuh"

namespace tb {

constexpr int kLocateThreads = 256;

// One CTA walks kLocItems*kLocateThreads consecutive points.  Histogram updates go to
// a small shared-memory table (bin -> count) first, so the global bin counters see one
// atomic per (CTA, distinct bin) instead of one per warp: departure points of one source
// leaf land in a handful of leaves, and thousands of same-address L2 atomics serialise.
constexpr int kLocItems = 8;
#ifndef TB_LOCATE_GUESSES
#define TB_LOCATE_GUESSES 1  // leaves a warp remembers between batches (2 measured on C2: 4.81 ms against 4.77, the kernel is issue bound)
#endif
constexpr int kLocTable = 64;  // distinct bins a CTA can aggregate; overflow -> direct atomics

__device__ __forceinline__ int table_slot(int *s_bin, int bin) {
  unsigned h = ((unsigned)bin * 2654435761u) >> 26;  // log2(kLocTable) = 6 bits
  for (int probe = 0; probe < kLocTable; probe++) {
    const int old = atomicCAS(s_bin + h, -1, bin);
    if (old == -1 || old == bin) return (int)h;
    h = (h + 1) & (kLocTable - 1);
  }
  return -1;
}

// Bins: 0..n_leaf-1 leaves, n_leaf = "no leaf" (evaluates to 0), n_leaf+2+r = points owned
// by rank r (multi-rank only; count[] holds the send counts right after the leaf bins).
//
// Fast path (BOXES: the leaves are aligned, non-overlapping octants): a point is first tested
// against the integer box of the leaf this warp resolved last -- three XORs, a shift and a
// compare on the depth-15 anchors floor(c * 2^15); inside the box implies "last leaf with
// key <= key(point)", so the Morton key is only interleaved, and the leaf list only searched,
// for the lanes that miss.  Departure points arrive in leaf-major order, so nearly all hit.
// Slow path: per-lane lookup of the key in the tree's cell table (number of leaf keys <= the first
// key of every cell of a uniform depth-g grid) + a binary search over the few leaves of that cell.
template <bool MULTI, bool BOXES>
__global__ void __launch_bounds__(kLocateThreads)
locate_kernel(const uint64_t *__restrict__ keys, const uint4 *__restrict__ boxes,
              const uint32_t *__restrict__ cells, int cell_shift, int n_leaf, int periodic,
              double *__restrict__ pos, size_t n, const uint32_t *__restrict__ n_dev,
              int32_t *__restrict__ leaf_out,
              uint32_t *__restrict__ rank_out, uint32_t *__restrict__ count,
              const uint64_t *__restrict__ splitters, int nranks, int myrank) {
  if (n_dev) {  // point count known on the device only: the grid was sized for the capacity
    const size_t nd = *n_dev;
    n = nd < n ? nd : n;
    if ((size_t)blockIdx.x * (kLocateThreads * kLocItems) >= n) return;
  }
  __shared__ int s_bin[kLocTable];
  __shared__ unsigned s_cnt[kLocTable];
  __shared__ unsigned s_base[kLocTable];
  if (threadIdx.x < kLocTable) {
    s_bin[threadIdx.x] = -1;
    s_cnt[threadIdx.x] = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1;
  const size_t chunk0 = (size_t)blockIdx.x * (kLocateThreads * kLocItems);
  // per point: (row of the CTA table + 1) << 24 | rank inside that row, completed with the
  // row's global base once the whole CTA has counted (row 0: the rank is final already)
  __shared__ uint32_t s_rec[kLocItems][kLocateThreads];
  // the two leaves the previous batches of this warp resolved most lanes to (warp-uniform): a warp's 32
  // consecutive departure points usually straddle two leaves, and one remembered leaf sent the lanes
  // of the other through the key search in every batch
  int guess = -1, guess2 = -1;
  uint4 gbox = make_uint4(0, 0, 0, 0), gbox2 = make_uint4(0, 0, 0, 0);  // their integer boxes
  int gslot = -2, gslot2 = -2;  // their rows of the CTA table (-2: not looked up yet)

  auto claim = [&](int bin, unsigned m, int leader, int &slot, uint32_t &base) {
    // one table (or, on overflow, global) update for the lanes in m, all in `bin`
    if (lane == leader) {
      slot = table_slot(s_bin, bin);
      base = (slot >= 0) ? atomicAdd(s_cnt + slot, (unsigned)__popc(m))
                         : atomicAdd(count + bin, (unsigned)__popc(m));
    }
  };

  // coordinates are fetched one batch ahead so their latency overlaps the previous batch
  double nx = 0, ny = 0, nz = 0;
  if (chunk0 + threadIdx.x < n) {
    const size_t i0 = chunk0 + threadIdx.x;
    nx = pos[3 * i0];
    ny = pos[3 * i0 + 1];
    nz = pos[3 * i0 + 2];
  }
#pragma unroll 1
  for (int it = 0; it < kLocItems; it++) {
    const size_t i = chunk0 + (size_t)it * kLocateThreads + threadIdx.x;
    const bool valid = i < n;
    // depth-15 anchors (tree_functor.h:464-479); `in` is false when any anchor leaves the
    // 15-bit range or is NaN -- such a point has the saturated key (keys.cuh)
    unsigned ix = 0, iy = 0, iz = 0;
    bool in = false;
    double x = nx, y = ny, z = nz;
    if (it + 1 < kLocItems && i + kLocateThreads < n) {
      const size_t i1 = i + kLocateThreads;
      nx = pos[3 * i1];
      ny = pos[3 * i1 + 1];
      nz = pos[3 * i1 + 2];
    }
    if (valid) {
      if (periodic) {
        const double x0 = x, y0 = y, z0 = z;
        x = wrap_periodic(x);
        y = wrap_periodic(y);
        z = wrap_periodic(z);
        if (x != x0) pos[3 * i] = x;  // the reference rewrites the caller's buffer
        if (y != y0) pos[3 * i + 1] = y;
        if (z != z0) pos[3 * i + 2] = z;
      }
      const double xs = x * 32768.0, ys = y * 32768.0, zs = z * 32768.0;  // exact
      int jx = __double2int_rd(xs), jy = __double2int_rd(ys), jz = __double2int_rd(zs);
      if (!periodic) {  // a coordinate == 1.0 moves down by 2^-15
        if (xs == 32768.0) jx = 32767;
        if (ys == 32768.0) jy = 32767;
        if (zs == 32768.0) jz = 32767;
      }
      const double chk = xs + ys + zs;  // NaN if any is (the conversions above turn NaN into 0)
      ix = (unsigned)jx;
      iy = (unsigned)jy;
      iz = (unsigned)jz;
      in = ((ix | iy | iz) < 32768u) && (chk == chk);
    }
    int bin = 0, slot = -1;
    uint32_t rank = 0;
    bool todo = valid;

    // lanes inside the box of a remembered leaf g: one aggregated table update, no key, no search
    auto try_guess = [&](int g, const uint4 &gb, int &gs) -> int {
      const bool hit = todo && in && ((((ix ^ gb.x) | (iy ^ gb.y) | (iz ^ gb.z)) >> gb.w) == 0u);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) {
        const int claimer = __ffs(m) - 1;
        if (gs == -2) {  // first hit on this guess: find its row of the CTA table once
          int sl = -1;
          if (lane == claimer) sl = table_slot(s_bin, g);
          gs = __shfl_sync(0xffffffffu, sl, claimer);
        }
        uint32_t base = 0;
        if (lane == claimer)
          base = (gs >= 0) ? atomicAdd(s_cnt + gs, (unsigned)__popc(m)) : atomicAdd(count + g, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, claimer);
        if (hit) {
          bin = g;
          slot = gs;
          rank = base + __popc(m & lt);
          todo = false;
        }
      }
      return __popc(m);
    };
    int n_hit = 0, n_hit2 = 0;  // lanes the remembered leaves resolved
    if (BOXES && guess >= 0) n_hit = try_guess(guess, gbox, gslot);                    // warp-uniform conditions
    if (BOXES && TB_LOCATE_GUESSES > 1 && guess2 >= 0) n_hit2 = try_guess(guess2, gbox2, gslot2);

    if (__any_sync(0xffffffffu, todo)) {  // slow path: keys, owners, searches
      const uint64_t key = in ? anchor_key(ix, iy, iz) : ~0ull;
      if (MULTI) {
        // owner = last rank whose first-leaf key is <= key (rank 0 below the first)
        bool mine = true;
        int owner = myrank;
        if (todo) {
          mine = (key >= __ldg(splitters + myrank)) || myrank == 0;
          if (myrank + 1 < nranks) mine = mine && key < __ldg(splitters + myrank + 1);
          if (!mine) {
            owner = 0;
            for (int r = 1; r < nranks; r++)
              if (__ldg(splitters + r) <= key) owner = r;
          }
        }
        unsigned out = __ballot_sync(0xffffffffu, todo && !mine);
        while (out) {  // one aggregated update per destination rank present in the warp
          const int leader = __ffs(out) - 1;
          const int o = __shfl_sync(0xffffffffu, owner, leader);
          const unsigned m = __ballot_sync(0xffffffffu, ((out >> lane) & 1u) && owner == o);
          int sl = -1;
          uint32_t base = 0;
          claim(n_leaf + 2 + o, m, leader, sl, base);
          sl = __shfl_sync(0xffffffffu, sl, leader);
          base = __shfl_sync(0xffffffffu, base, leader);
          if ((m >> lane) & 1u) {
            bin = n_leaf + 2 + o;
            slot = sl;
            rank = base + __popc(m & lt);
            todo = false;
          }
          out &= ~m;
        }
      }

      // Per-lane lookup: the cell table gives the range of leaves that can hold the key (count of
      // leaf keys <= first key of the cell, and of the next cell), a short binary search finishes
      // (<= 3 steps when the leaves are at most one level finer than the table).
      int j = -1, n_new = 0;
      if (todo) {
        if (key == ~0ull) {
          j = n_leaf - 1;  // saturated key: after every leaf key
        } else {
          const unsigned c = (unsigned)(key >> cell_shift);
          int lo = (int)__ldg(cells + c), hi = (int)__ldg(cells + c + 1);
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(keys + mid) <= key)
              lo = mid + 1;
            else
              hi = mid;
          }
          j = lo - 1;
        }
        bin = (j >= 0) ? j : n_leaf;
      }
      const unsigned pending = __ballot_sync(0xffffffffu, todo);
      if (todo) {  // one table (or global) update per distinct leaf among the pending lanes
        const unsigned peers = __match_any_sync(pending, bin);
        const int leader = __ffs(peers) - 1;
        int sl = -1;
        uint32_t base = 0;
        claim(bin, peers, leader, sl, base);
        slot = __shfl_sync(peers, sl, leader);
        rank = __shfl_sync(peers, base, leader) + __popc(peers & lt);
        n_new = (j >= 0) ? __popc(peers) : 0;
      }
      // the leaf that took the most lanes replaces the remembered leaf that resolved fewer lanes of
      // this batch, unless that one still holds more of the warp (a lane a remembered leaf contains
      // never reaches the search, so the new leaf differs from both)
      const unsigned best = __reduce_max_sync(0xffffffffu, ((unsigned)n_new << 8) | (unsigned)(31 - lane));
      const int n_best = (int)(best >> 8);
      if (TB_LOCATE_GUESSES > 1 && (guess2 < 0 || n_hit2 < n_hit) && guess >= 0) {
        if (n_best > n_hit2) {
          guess2 = __shfl_sync(0xffffffffu, bin, 31 - (int)(best & 0xffu));
          gslot2 = -2;
          if (BOXES) gbox2 = __ldg(boxes + guess2);
        }
      } else if (n_best > n_hit) {
        guess = __shfl_sync(0xffffffffu, bin, 31 - (int)(best & 0xffu));
        gslot = -2;
        if (BOXES) gbox = __ldg(boxes + guess);
      }
    }
    if (valid) {
      leaf_out[i] = (bin <= n_leaf) ? bin : -2 - (bin - n_leaf - 2);
      if (slot >= 0) {  // rank inside a table row: < points per CTA (2^11)
        s_rec[it][threadIdx.x] = ((uint32_t)(slot + 1) << 24) | rank;
      } else {          // table overflow: counted straight into the global bin, rank is final
        s_rec[it][threadIdx.x] = 0u;
        rank_out[i] = rank;
      }
    }
  }

  __syncthreads();
  if (threadIdx.x < kLocTable && s_bin[threadIdx.x] >= 0)
    s_base[threadIdx.x] = atomicAdd(count + s_bin[threadIdx.x], s_cnt[threadIdx.x]);
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kLocItems; it++) {
    const size_t i = chunk0 + (size_t)it * kLocateThreads + threadIdx.x;
    if (i >= n) break;
    const uint32_t r = s_rec[it][threadIdx.x];
    if (r >> 24) rank_out[i] = (r & 0xffffffu) + s_base[(r >> 24) - 1];
  }
}

int launch_locate(tbslas_ctx *ctx, const LocateArgs &a) {
  const tbslas_tree *t = a.tree;
  StageScope sc(ctx, ST_LOCATE, (double)a.n, 1);
  const bool multi = ctx->nranks > 1 && a.send_count;
  TB_CUDA(ctx, cudaMemsetAsync(a.count, 0, sizeof(uint32_t) * (t->n_leaf + 2 + kMaxRanks + 1), ctx->stream));
  if (a.n == 0) return TBSLAS_OK;
  const size_t per_cta = (size_t)kLocateThreads * kLocItems;
  const unsigned grid = (unsigned)((a.n + per_cta - 1) / per_cta);
  const bool boxes = t->boxes_ok && !ctx->opt.locate_no_boxes;
  if (multi && a.send_count != a.count + t->n_leaf + 2)
    return fail(ctx, TBSLAS_ERR_INVALID, "send counts must follow the leaf bins");
#define TB_LOCATE(M, B)                                                                          \
  locate_kernel<M, B><<<grid, kLocateThreads, 0, ctx->stream>>>(                                 \
      t->d_key, t->d_box, t->d_cell, t->cell_shift, (int)t->n_leaf, a.periodic, a.pos, a.n,      \
      a.n_dev, a.leaf, a.rank, a.count,                                                                   \
      multi ? t->d_splitters : nullptr, multi ? ctx->nranks : 1, multi ? ctx->rank : 0)
  if (multi) {
    if (boxes) TB_LOCATE(true, true); else TB_LOCATE(true, false);
  } else {
    if (boxes) TB_LOCATE(false, true); else TB_LOCATE(false, false);
  }
#undef TB_LOCATE
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
                 uint32_t *__restrict__ bin_start, uint32_t *__restrict__ tile_start) {
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

// tile t -> (leaf, first slot): every tile finds its leaf by binary search in tile_start
__global__ void tile_map_kernel(const uint32_t *__restrict__ bin_start,
                                const uint32_t *__restrict__ tile_start, int n_bins, int tile_pts,
                                int2 *__restrict__ tile_map, unsigned max_tiles) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned n_tiles = min(__ldg(tile_start + n_bins), max_tiles);
  if (t >= n_tiles) return;
  int lo = 0, hi = n_bins;  // last j with tile_start[j] <= t (empty bins share a start: take the last)
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(tile_start + mid) <= t)
      lo = mid;
    else
      hi = mid;
  }
  tile_map[t] = make_int2(lo, (int)(__ldg(bin_start + lo) + (t - __ldg(tile_start + lo)) * tile_pts));
}

// Insiders: point id -> its slot in the leaf grouping.  Outsiders: coordinates and origin index
// -> the bucket of the owner rank (reference: the forward scatter of par::ScatterForward,
// tree_functor.h:574-575).  MODE 1 (NCCL exchange): a local send buffer in bucket order, bucket
// offsets = exclusive scan of the per-rank send counts, rebuilt per CTA in shared memory
// (nranks <= 64).  MODE 2 (peer exchange): the coordinates go straight into the owner's receive
// buffer over NVLink peer memory, at the place comm.cu's px_offsets_kernel derived from the count
// matrix; only the origin index stays here (bucket order, for the unpack).
constexpr int kScatterThreads = 256;
#ifndef TB_SCATTER_ITEMS
#define TB_SCATTER_ITEMS 4
#endif
constexpr int kScatterItems = TB_SCATTER_ITEMS;

template <int MODE>
__global__ void __launch_bounds__(kScatterThreads)
scatter_perm_kernel(const int32_t *__restrict__ leaf, const uint32_t *__restrict__ rank,
                                    const uint32_t *__restrict__ bin_start, size_t n,
                                    const uint32_t *__restrict__ n_dev,
                                    uint32_t *__restrict__ perm, const double *__restrict__ pos,
                                    const uint32_t *__restrict__ send_count, int nranks,
                                    double *__restrict__ send_pos, uint32_t *__restrict__ send_idx, PxPack px) {
  __shared__ unsigned s_off[kMaxRanks], s_dst[kMaxRanks];
  __shared__ unsigned s_skip;
  if (MODE == 1) {
    if (threadIdx.x == 0) {
      unsigned run = 0;
      for (int r = 0; r < nranks; r++) {
        s_off[r] = run;
        run += __ldg(send_count + r);
      }
    }
    __syncthreads();
  }
  if (MODE == 2) {
    if ((int)threadIdx.x < nranks) {
      s_off[threadIdx.x] = px.send_off[threadIdx.x];
      s_dst[threadIdx.x] = px.dst_off[threadIdx.x];
    }
    // a count matrix that does not fit the mailboxes (or a peer that never arrived): the offsets would
    // point past the owners' buffers, so nothing travels; the error is raised at the next host sync
    if (threadIdx.x == 0) s_skip = px.skip[0] | px.skip[1];
    __syncthreads();
  }
  if (n_dev) {
    const size_t nd = *n_dev;
    n = nd < n ? nd : n;
  }
  // kScatterItems points per thread, all (leaf, rank) loads issued before the first dependent
  // bin_start load: the pass is pure latency otherwise (one point per thread ran at 2.6 TB/s)
  const size_t i0 = (size_t)blockIdx.x * (kScatterThreads * kScatterItems) + threadIdx.x;
  int jv[kScatterItems];
  uint32_t rv[kScatterItems];
#pragma unroll
  for (int k = 0; k < kScatterItems; k++) {
    const size_t i = i0 + (size_t)k * kScatterThreads;
    jv[k] = i < n ? leaf[i] : 0;
    rv[k] = i < n ? rank[i] : 0u;
  }
  uint32_t bs[kScatterItems];
#pragma unroll
  for (int k = 0; k < kScatterItems; k++) bs[k] = jv[k] >= 0 ? __ldg(bin_start + jv[k]) : 0u;
#pragma unroll
  for (int k = 0; k < kScatterItems; k++) {
    const size_t i = i0 + (size_t)k * kScatterThreads;
    if (i >= n) break;
    const int j = jv[k];
    if (j >= 0) {
      perm[bs[k] + rv[k]] = (uint32_t)i;
    } else if (MODE == 1) {  // travels to its owner
      const unsigned slot = s_off[-2 - j] + rv[k];
      send_pos[3 * (size_t)slot] = pos[3 * i];
      send_pos[3 * (size_t)slot + 1] = pos[3 * i + 1];
      send_pos[3 * (size_t)slot + 2] = pos[3 * i + 2];
      send_idx[slot] = (uint32_t)i;
    } else if (MODE == 2 && !s_skip) {
      const int o = -2 - j;
      const unsigned r = rv[k];
      double *dst = reinterpret_cast<double *>(px.peer_base[o] + px.off_recv_pos) + 3 * ((size_t)s_dst[o] + r);
      dst[0] = pos[3 * i];
      dst[1] = pos[3 * i + 1];
      dst[2] = pos[3 * i + 2];
      send_idx[s_off[o] + r] = (uint32_t)i;
    }
  }
}

int launch_bin(tbslas_ctx *ctx, const BinArgs &a) {
  StageScope sc(ctx, ST_BIN, (double)a.n, a.tile_map ? 3 : 2);
  const int n_bins = (int)a.n_leaf + 1;  // + null leaf
  scan_bins_kernel<<<1, kScanThreads, 0, ctx->stream>>>(a.count, n_bins, a.tile_pts, a.bin_start,
                                                        a.tile_start);
  TB_CUDA(ctx, cudaGetLastError());
  if (a.tile_map) {
    tile_map_kernel<<<(unsigned)((a.max_tiles + 255) / 256), 256, 0, ctx->stream>>>(
        a.bin_start, a.tile_start, n_bins, a.tile_pts, a.tile_map, (unsigned)a.max_tiles);
    TB_CUDA(ctx, cudaGetLastError());
  }
  if (a.n) {
    const size_t per_cta = (size_t)kScatterThreads * kScatterItems;
    const unsigned grid = (unsigned)((a.n + per_cta - 1) / per_cta);
    if (a.px)
      scatter_perm_kernel<2><<<grid, 256, 0, ctx->stream>>>(a.leaf, a.rank, a.bin_start, a.n, a.n_dev, a.perm,
                                                          a.pos, a.send_count, a.nranks, nullptr, a.send_idx,
                                                          *a.px);
    else if (a.send_count)
      scatter_perm_kernel<1><<<grid, 256, 0, ctx->stream>>>(a.leaf, a.rank, a.bin_start, a.n, a.n_dev, a.perm,
                                                          a.pos, a.send_count, a.nranks, a.send_pos, a.send_idx,
                                                          PxPack());
    else
      scatter_perm_kernel<0><<<grid, 256, 0, ctx->stream>>>(a.leaf, a.rank, a.bin_start, a.n, a.n_dev, a.perm,
                                                          nullptr, nullptr, 1, nullptr, nullptr, PxPack());
    TB_CUDA(ctx, cudaGetLastError());
  }
  return TBSLAS_OK;
}

}  // namespace tb

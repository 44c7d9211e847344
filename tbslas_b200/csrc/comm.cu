// Multi-GPU point redistribution: one process per GPU, octree leaves sharded by Morton
// range, foreign departure points exchanged with NCCL all-to-all-v over NVLink and the
// evaluated values returned the same way.
//
// Replaces, per tree evaluation (reference src/tree/tree_functor.h):
//   :433-437  MPI_Allgather of every rank's first-leaf Morton id  -> once per tree, at
//             tree_create (comm_tree_splitters), kept on the device
//   :491-513  owner split of the local points                     -> locate_kernel<MULTI>
//   :569-570  par::SortScatterIndex (distributed sort + counts)   -> per-owner buckets built
//             by the locate/scatter kernels; the nranks x nranks count matrix travels in
//             one ncclAllGather
//   :574-575  par::ScatterForward   (3 doubles per outsider)      -> grouped ncclSend/ncclRecv
//   :587      EvalNodesLocal on the received points               -> the same locate/bin/eval
//             kernels on the receive buffer
//   :594-595  par::ScatterReverse   (dof doubles per outsider)    -> grouped ncclSend/ncclRecv
//   :602-619  scatter of the returned values to out[orig idx]     -> unpack_kernel (fused
//             with the RK2 position update when the caller asked for it)
// The reference evaluates outsiders BEFORE insiders, serially; here the forward exchange runs
// on a second stream while the insiders are being evaluated.
//
// NCCL is dlopen'ed at comm_init (libnccl.so.2: inside a torch process this resolves to the
// library torch already loaded), so single-GPU users and CPU-only symbol checks never
// need it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "common.cuh"
#include "gridbase.cuh"

namespace tb {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  int (*GetVersion)(int *) = nullptr;
};

static NcclApi g_nccl;
static const char *g_nccl_err = "";

static bool nccl_load() {
  if (g_nccl.handle) return true;
  void *h = nullptr;
  for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
    h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    g_nccl_err = "libnccl.so.2 not found (dlopen)";
    return false;
  }
#define TB_SYM(field, sym)                                     \
  *(void **)(&g_nccl.field) = dlsym(h, sym);                   \
  if (!g_nccl.field) {                                         \
    g_nccl_err = "libnccl lacks " sym;                         \
    return false;                                              \
  }
  TB_SYM(GetUniqueId, "ncclGetUniqueId")
  TB_SYM(CommInitRank, "ncclCommInitRank")
  TB_SYM(CommDestroy, "ncclCommDestroy")
  TB_SYM(AllGather, "ncclAllGather")
  TB_SYM(Send, "ncclSend")
  TB_SYM(Recv, "ncclRecv")
  TB_SYM(GroupStart, "ncclGroupStart")
  TB_SYM(GroupEnd, "ncclGroupEnd")
  TB_SYM(GetErrorString, "ncclGetErrorString")
#undef TB_SYM
  g_nccl.handle = h;
  return true;
}

#define TB_NCCL(ctx, call)                                                                  \
  do {                                                                                      \
    ncclResult_t r_ = (call);                                                               \
    if (r_ != ncclSuccess)                                                                  \
      return fail((ctx), TBSLAS_ERR_COMM, "%s: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), \
                  __FILE__, __LINE__);                                                      \
  } while (0)

static inline ncclComm_t comm_of(tbslas_ctx *ctx) { return (ncclComm_t)ctx->nccl_comm; }
int px_setup(tbslas_ctx *ctx, size_t cap);
int px_grow(tbslas_ctx *ctx, size_t want);

// `later` waits for everything enqueued so far on `earlier`
static int chain(tbslas_ctx *ctx, cudaStream_t earlier, cudaStream_t later) {
  TB_CUDA(ctx, cudaEventRecord(ctx->ev_comm, earlier));
  TB_CUDA(ctx, cudaStreamWaitEvent(later, ctx->ev_comm, 0));
  return TBSLAS_OK;
}

// All-to-all-v of `elem` bytes per item: send_cnt[r] items to rank r from `send` (bucket
// order), recv_cnt[r] items from rank r into `recv`.
static int alltoallv(tbslas_ctx *ctx, const void *send, const unsigned *send_cnt, void *recv,
                     const unsigned *recv_cnt, size_t elem, cudaStream_t s) {
  const int np = ctx->nranks;
  size_t so = 0, ro = 0;
  TB_NCCL(ctx, g_nccl.GroupStart());
  for (int r = 0; r < np; r++) {
    if (r != ctx->rank) {
      if (send_cnt[r])
        TB_NCCL(ctx, g_nccl.Send((const char *)send + so * elem, (size_t)send_cnt[r] * elem, ncclUint8,
                                 r, comm_of(ctx), s));
      if (recv_cnt[r])
        TB_NCCL(ctx, g_nccl.Recv((char *)recv + ro * elem, (size_t)recv_cnt[r] * elem, ncclUint8, r,
                                 comm_of(ctx), s));
    }
    so += send_cnt[r];
    ro += recv_cnt[r];
  }
  TB_NCCL(ctx, g_nccl.GroupEnd());
  return TBSLAS_OK;
}

// ---------------------------------------------------------------------------
// splitters: first-leaf key of every rank (tree_functor.h:425,433-437), once per tree
// ---------------------------------------------------------------------------
int comm_tree_splitters(tbslas_tree *t, uint64_t first_key) {
  tbslas_ctx *ctx = t->ctx;
  const int np = ctx->nranks;
  // {first key, leaf count, hash of the local leaf list, boxes usable} of every rank
  constexpr int W = 4;
  unsigned long long mine[W] = {first_key, (unsigned long long)t->n_leaf, t->struct_hash,
                                (unsigned long long)(t->boxes_ok ? 1 : 0)};
  void *buf;
  TB_TRY(ws_get(ctx, WS_MISC, sizeof(mine) * (np + 1), &buf));
  unsigned long long *d_mine = (unsigned long long *)buf, *d_all = d_mine + W;
  TB_CUDA(ctx, cudaMemcpyAsync(d_mine, mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
  TB_NCCL(ctx, g_nccl.AllGather(d_mine, d_all, W, ncclUint64, comm_of(ctx), ctx->stream));
  std::vector<unsigned long long> all(W * np);
  TB_CUDA(ctx, cudaMemcpyAsync(all.data(), d_all, sizeof(mine) * np, cudaMemcpyDeviceToHost, ctx->stream));
  TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  t->splitters.assign(np, 0);
  long long off = 0;
  uint64_t gh = 1469598103934665603ull;
  t->n_leaf_max = 0;
  t->boxes_all = true;
  for (int r = 0; r < np; r++) {
    t->splitters[r] = all[W * r];
    if (r < ctx->rank) off += (long long)all[W * r + 1];
    // what every rank must agree on: one hash over all the local leaf lists, the largest shard,
    // and whether every shard's leaves are aligned, disjoint octants
    gh = (gh ^ all[W * r + 1]) * 1099511628211ull;
    gh = (gh ^ all[W * r + 2]) * 1099511628211ull;
    if (all[W * r + 1] > t->n_leaf_max) t->n_leaf_max = (size_t)all[W * r + 1];
    if (all[W * r + 1] && !all[W * r + 3]) t->boxes_all = false;
  }
  t->global_hash = gh;
  t->rank_first.assign(np + 1, 0);
  for (int r = 0; r < np; r++) t->rank_first[r + 1] = t->rank_first[r] + (size_t)all[W * r + 1];
  t->n_leaf_global = t->rank_first[np];
  // a rank without leaves owns the empty range: give it the next owner's first key so the
  // "last rank whose splitter <= key" rule never selects it
  for (int r = np - 1; r >= 0; r--)
    if (all[W * r + 1] == 0) t->splitters[r] = (r + 1 < np) ? t->splitters[r + 1] : ~0ull;
  for (int r = 1; r < np; r++)
    if (t->splitters[r] < t->splitters[r - 1])
      return fail(ctx, TBSLAS_ERR_INVALID,
                  "ranks must own ascending Morton ranges (rank %d starts before rank %d)", r, r - 1);
  t->leaf_offset = off;
  {  // room for a quarter of the largest shard's arrival points each way (collective: same on all ranks)
    const size_t d = t->q + 1, want = t->n_leaf_max * d * d * d / 4;
    if (want <= 0x7fffffffu) TB_TRY(px_grow(ctx, want));
  }
  if (!t->d_splitters) TB_CUDA(ctx, cudaMalloc(&t->d_splitters, sizeof(uint64_t) * kMaxRanks));
  TB_CUDA(ctx, cudaMemcpyAsync(t->d_splitters, t->splitters.data(), sizeof(uint64_t) * np, cudaMemcpyHostToDevice,
                               ctx->stream));
  TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TBSLAS_OK;
}

// all-gather of `bytes` bytes per rank on the context's stream (no host synchronisation)
int comm_allgather_bytes(tbslas_ctx *ctx, const void *mine, void *all, size_t bytes) {
  TB_NCCL(ctx, g_nccl.AllGather(mine, all, bytes, ncclUint8, comm_of(ctx), ctx->stream));
  return TBSLAS_OK;
}

// ---------------------------------------------------------------------------
// co-partitioning: leaves move between ranks (tbslas::SemiMergeTree -> pvfmm RedistNodes,
// tree_utils.h:609-729: the reference recomputes break points and MOVES the nodes)
// ---------------------------------------------------------------------------
// Rank r ends up with global leaves [new_first[r], new_first[r+1]).  Old and new ranges are both
// contiguous in the global Morton order, so what rank a sends to rank b is one contiguous run:
// geometry (32 B), depth (1 B) and the coefficient block (stride doubles) of every leaf in
// [max(old_a, new_b), min(old_a', new_b')).  Grouped ncclSend/ncclRecv straight between the trees'
// device arrays; the derived structures (keys, boxes, cell table, hashes, split keys) are rebuilt
// from the received leaf list exactly as tree_create builds them.
int comm_reshard(tbslas_tree *t, const size_t *new_first) {
  tbslas_ctx *ctx = t->ctx;
  const int np = ctx->nranks, me = ctx->rank;
  if ((int)t->rank_first.size() != np + 1) return fail(ctx, TBSLAS_ERR_INVALID, "tree_reshard: not a sharded tree");
  const std::vector<size_t> &of = t->rank_first;
  if (new_first[0] != 0 || new_first[np] != of[np])
    return fail(ctx, TBSLAS_ERR_INVALID, "tree_reshard: new ranges must cover [0, %zu)", of[np]);
  for (int r = 0; r < np; r++)
    if (new_first[r] > new_first[r + 1]) return fail(ctx, TBSLAS_ERR_INVALID, "tree_reshard: ranges must ascend");
  const size_t n_new = new_first[me + 1] - new_first[me], stride = t->stride;
  double4 *g_new = nullptr;
  uint8_t *d_new = nullptr;
  double *c_new = nullptr;
  TB_CUDA(ctx, cudaMalloc(&g_new, sizeof(double4) * (n_new + 1)));
  TB_CUDA(ctx, cudaMalloc(&d_new, n_new + 1));
  TB_CUDA(ctx, cudaMalloc(&c_new, sizeof(double) * stride * (n_new + 1)));
  TB_CUDA(ctx, cudaMemsetAsync(c_new + stride * n_new, 0, sizeof(double) * stride, ctx->stream));  // null leaf
  auto overlap = [](size_t a0, size_t a1, size_t b0, size_t b1, size_t *lo, size_t *hi) {
    *lo = a0 > b0 ? a0 : b0;
    *hi = a1 < b1 ? a1 : b1;
    return *lo < *hi;
  };
  TB_NCCL(ctx, g_nccl.GroupStart());
  for (int r = 0; r < np; r++) {
    size_t lo, hi;
    if (overlap(of[me], of[me + 1], new_first[r], new_first[r + 1], &lo, &hi)) {  // mine -> r
      const size_t s = lo - of[me], m = hi - lo;
      if (r == me) {
        const size_t d = lo - new_first[me];
        TB_CUDA(ctx, cudaMemcpyAsync(g_new + d, t->d_geom + s, sizeof(double4) * m, cudaMemcpyDeviceToDevice, ctx->stream));
        TB_CUDA(ctx, cudaMemcpyAsync(d_new + d, t->d_depth + s, m, cudaMemcpyDeviceToDevice, ctx->stream));
        TB_CUDA(ctx, cudaMemcpyAsync(c_new + d * stride, t->d_coeff + s * stride, sizeof(double) * stride * m,
                                     cudaMemcpyDeviceToDevice, ctx->stream));
      } else {
        TB_NCCL(ctx, g_nccl.Send(t->d_geom + s, sizeof(double4) * m, ncclUint8, r, comm_of(ctx), ctx->stream));
        TB_NCCL(ctx, g_nccl.Send(t->d_depth + s, m, ncclUint8, r, comm_of(ctx), ctx->stream));
        TB_NCCL(ctx, g_nccl.Send(t->d_coeff + s * stride, sizeof(double) * stride * m, ncclUint8, r, comm_of(ctx),
                                 ctx->stream));
      }
    }
    if (r != me && overlap(of[r], of[r + 1], new_first[me], new_first[me + 1], &lo, &hi)) {  // r's -> me
      const size_t d = lo - new_first[me], m = hi - lo;
      TB_NCCL(ctx, g_nccl.Recv(g_new + d, sizeof(double4) * m, ncclUint8, r, comm_of(ctx), ctx->stream));
      TB_NCCL(ctx, g_nccl.Recv(d_new + d, m, ncclUint8, r, comm_of(ctx), ctx->stream));
      TB_NCCL(ctx, g_nccl.Recv(c_new + d * stride, sizeof(double) * stride * m, ncclUint8, r, comm_of(ctx),
                               ctx->stream));
    }
  }
  TB_NCCL(ctx, g_nccl.GroupEnd());
  // the new leaf list on the host -> derived structures
  std::vector<double4> hg(n_new);
  std::vector<uint8_t> hd(n_new);
  TB_CUDA(ctx, cudaMemcpyAsync(hg.data(), g_new, sizeof(double4) * n_new, cudaMemcpyDeviceToHost, ctx->stream));
  TB_CUDA(ctx, cudaMemcpyAsync(hd.data(), d_new, n_new, cudaMemcpyDeviceToHost, ctx->stream));
  TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(g_new);
  cudaFree(d_new);
  std::vector<double> hc(3 * n_new);
  for (size_t j = 0; j < n_new; j++) {
    hc[3 * j] = hg[j].x;
    hc[3 * j + 1] = hg[j].y;
    hc[3 * j + 2] = hg[j].z;
  }
  const int rc = tree_build_structure(t, hc, hd, false);
  if (rc != TBSLAS_OK) {
    cudaFree(c_new);
    return rc;
  }
  cudaFree(t->d_coeff);
  t->d_coeff = c_new;
  return comm_tree_splitters(t, t->first_key);
}

// ---------------------------------------------------------------------------
// return path: values (and leaf ids) of my outsiders -> the caller's arrays
// ---------------------------------------------------------------------------
// m_dev != nullptr: the number of outsiders is only known on the device (peer exchange); the grid
// is sized for the capacity and strides.
template <int EPI>
__global__ void unpack_kernel(const double *__restrict__ val, const int32_t *__restrict__ leaf_in,
                              const uint32_t *__restrict__ idx, size_t m, const uint32_t *__restrict__ m_dev,
                              int dof, double *__restrict__ out, const double *__restrict__ base, double alpha,
                              int32_t *__restrict__ leaf_out, const GridBase gb) {
  if (m_dev) m = *m_dev;
  const size_t total = m * dof, step = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {  // one scalar each
    const size_t s = e / dof;
    const int l = (int)(e - s * dof);
    const size_t i = idx[s];
    const double u = val[e];
    if (EPI == EPI_STORE) {
      out[i * dof + l] = u;
    } else if (EPI == EPI_AXPY) {  // same expression as the eval kernel's epilogue (traj.inc:36,42)
      out[3 * i + l] = __dadd_rn(base[3 * i + l], __dmul_rn(alpha, u));
    } else {  // base rebuilt from the grid (gridbase.cuh)
      const double4 g = gb.ggeom[(unsigned)i / gb.P];
      out[3 * i + l] = __dadd_rn(grid_base_coord(gb, g, (unsigned)i, l), __dmul_rn(alpha, u));
    }
    if (leaf_out && l == 0) leaf_out[i] = leaf_in[s];
  }
}

int eval_received_points(tbslas_tree *t, int bc, double *pos, size_t n, const uint32_t *n_dev, double *out,
                         int32_t *leaf_out);

__global__ void publish_words_kernel(const uint32_t *__restrict__ src, unsigned *host_dst, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) reinterpret_cast<volatile unsigned *>(host_dst)[i] = src[i];
  __threadfence_system();
}

// ===========================================================================
// Peer exchange ("mailbox"): the outsiders are written straight into their owner's receive
// buffer over NVLink peer memory and the values straight back into the origin's return buffer.
// ===========================================================================
// Every rank exports ONE device allocation (cudaIpc) laid out as
//     header | recv_pos [cap][3] f64 | ret_val [3*cap] f64 | ret_leaf [cap] i32
// and maps every peer's.  One evaluation = one EPOCH of three barriers, all on the device:
//   (1) counts   every rank stores its row of the send-count matrix into every peer's header and
//                raises flag_cnt[me] there; once all rows are in, each rank derives, from the SAME
//                matrix, where its buckets start in every owner's receive buffer, how much it
//                receives from whom, and where the values it computes go in each origin's return
//                buffer (px_offsets_kernel) -- par::SortScatterIndex without a sort;
//   (2) forward  the pack kernel (locate.cu, scatter_perm_kernel<2>) writes coordinates to
//                peer_recv_pos[owner][bucket start + rank in bucket]; flag_fwd[me] is raised on every
//                peer after it; the insiders are evaluated meanwhile; then the rank waits for the
//                flags of all peers and runs locate/bin/evaluate on what arrived, the point count
//                being read from device memory by every kernel;
//   (3) reverse  px_return_kernel writes the values (and leaf ids) into the origins' return buffers,
//                flag_ret[me] is raised everywhere, and after the peers' flags the unpack kernel
//                scatters the returned values to out[origin index] (fused with the RK2 update).
// The host enqueues a fixed sequence of kernels and never learns a count: no host synchronisation,
// no NCCL kernel resident next to the persistent evaluation kernel, no second stream.  A flag is the
// epoch number, so nothing is ever reset; ranks run their collective calls in the same order, which
// makes "flag >= epoch" the barrier.  A rank cannot run ahead by more than one phase: it needs every
// peer's flag of phase k before it writes anything of phase k+1 that a peer could still be reading.
// Capacity: cap points per rank each way (grown collectively at tree_create from the largest shard);
// a matrix that does not fit raises the sticky overflow error on every rank alike (all ranks evaluate
// the same predicate on the same matrix) and the exchange of that evaluation is skipped.
struct PxHeader {
  uint32_t cnt[kMaxRanks][kMaxRanks];  // cnt[src][dst], row src written by rank src
  uint32_t flag_cnt[kMaxRanks];
  uint32_t flag_fwd[kMaxRanks];
  uint32_t flag_ret[kMaxRanks];
};
constexpr size_t kPxHeaderBytes = 64 * 1024;
static_assert(sizeof(PxHeader) <= kPxHeaderBytes, "header region");

struct PxInfo {  // per evaluation, device resident, derived from the count matrix
  uint32_t send_cnt[kMaxRanks], send_off[kMaxRanks];  // my buckets in my own send order
  uint32_t dst_off[kMaxRanks];                        // where my bucket starts in owner d's recv_pos
  uint32_t recv_cnt[kMaxRanks], recv_off[kMaxRanks + 1];
  uint32_t ret_off[kMaxRanks];                        // where rank s's bucket for me starts in s's ret_val
  uint32_t n_send, n_recv, overflow, timeout;
};

struct PxLayout {
  size_t cap = 0;  // points
  size_t off_recv_pos() const { return kPxHeaderBytes; }
  size_t off_ret_val() const { return off_recv_pos() + 24 * cap; }
  size_t off_ret_leaf() const { return off_ret_val() + 24 * cap; }
  size_t bytes() const { return off_ret_leaf() + 4 * cap; }
};

struct PxPeers {  // device-resident table
  char *base[kMaxRanks];
};

struct ExchangeState {
  // NCCL all-to-all-v path
  unsigned send_cnt[kMaxRanks], recv_cnt[kMaxRanks];
  size_t n_send = 0, n_recv = 0;
  void *recv_pos = nullptr, *recv_val = nullptr, *ret_val = nullptr, *recv_leaf = nullptr, *ret_leaf = nullptr;
  // peer path
  bool px_ok = false;
  PxLayout lay;
  char *mailbox = nullptr;            // my exported allocation
  char *peer_base[kMaxRanks] = {};    // mapped peers (peer_base[me] == mailbox)
  PxPeers *d_peers = nullptr;
  PxInfo *d_info = nullptr;
  unsigned *h_err = nullptr;          // pinned, device-writable: sticky {overflow, timeout}
  uint32_t epoch = 0;
  bool info_valid = false;
};

__device__ __forceinline__ unsigned long long px_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kPxTimeoutNs = 30ull * 1000 * 1000 * 1000;

// wait until flag[r] >= epoch for all r < np (thread r waits for rank r); false on timeout
__device__ __forceinline__ bool px_wait_flags(const uint32_t *flags, int np, uint32_t epoch) {
  bool ok = true;
  if ((int)threadIdx.x < np) {
    const volatile uint32_t *f = flags + threadIdx.x;
    const unsigned long long t0 = px_now_ns();
    while ((int32_t)(*f - epoch) < 0) {
      __nanosleep(100);
      if (px_now_ns() - t0 > kPxTimeoutNs) {
        ok = false;
        break;
      }
    }
  }
  __threadfence_system();
  return __syncthreads_and(ok);
}

// phase (1a): my row of the count matrix -> every rank's header, then the flag
__global__ void px_post_counts_kernel(const uint32_t *__restrict__ send_count, const PxPeers *peers, int np, int me,
                                      uint32_t epoch) {
  const int r = threadIdx.x;
  if (r < np) {
    PxHeader *h = reinterpret_cast<PxHeader *>(peers->base[r]);
    for (int d = 0; d < np; d++) h->cnt[me][d] = send_count[d];
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&h->flag_cnt[me]) = epoch;
  }
}

// phase (1b): all rows are in -> offsets (one CTA of kMaxRanks threads)
__global__ void px_offsets_kernel(PxHeader *hdr, PxInfo *info, int np, int me, uint32_t epoch, unsigned cap,
                                  unsigned ret_cap_points, unsigned *h_err) {
  __shared__ uint32_t s_cnt[kMaxRanks][kMaxRanks + 1];
  __shared__ int s_over;
  const bool ok = px_wait_flags(hdr->flag_cnt, np, epoch);
  if (threadIdx.x == 0) s_over = 0;
  __syncthreads();
  const int r = threadIdx.x;
  if (r < np)
    for (int d = 0; d < np; d++) s_cnt[r][d] = *reinterpret_cast<volatile uint32_t *>(&hdr->cnt[r][d]);
  __syncthreads();
  if (r < np) {  // column sums (what rank r receives) and row sums (what rank r sends)
    unsigned long long col = 0, row = 0;
    for (int s = 0; s < np; s++) {
      col += s_cnt[s][r];
      row += s_cnt[r][s];
    }
    if (col > cap || row > ret_cap_points) atomicOr(&s_over, 1);
  }
  __syncthreads();
  const bool over = s_over != 0 || !ok;
  if (r < np) {
    unsigned so = 0, dof_ = 0, ro = 0, rt = 0;
    for (int d = 0; d < r; d++) so += s_cnt[me][d];   // my buckets before bucket r
    for (int s = 0; s < me; s++) dof_ += s_cnt[s][r];  // sources before me in owner r's buffer
    for (int s = 0; s < r; s++) ro += s_cnt[s][me];   // sources before r in my buffer
    for (int d = 0; d < me; d++) rt += s_cnt[r][d];   // origin r's buckets before its bucket for me
    info->send_cnt[r] = over ? 0u : s_cnt[me][r];
    info->send_off[r] = so;
    info->dst_off[r] = dof_;
    info->recv_cnt[r] = over ? 0u : s_cnt[r][me];
    info->recv_off[r] = over ? 0u : ro;
    info->ret_off[r] = rt;
  }
  if (r == 0) {
    unsigned ns = 0, nr = 0;
    for (int d = 0; d < np; d++) {
      ns += s_cnt[me][d];
      nr += s_cnt[d][me];
    }
    info->n_send = over ? 0u : ns;
    info->n_recv = over ? 0u : nr;
    info->recv_off[np] = over ? 0u : nr;
    info->overflow = s_over ? 1u : 0u;
    info->timeout = ok ? 0u : 1u;
    if (s_over) reinterpret_cast<volatile unsigned *>(h_err)[0] = 1u;
    if (!ok) reinterpret_cast<volatile unsigned *>(h_err)[1] = 1u;
    if (over) __threadfence_system();
  }
}

// raise flag `which` (0 fwd, 1 ret) of this epoch on every rank; all earlier writes of this
// stream (the kernel before this one has completed) are ordered before it
__global__ void px_signal_kernel(const PxPeers *peers, int np, int me, int which, uint32_t epoch) {
  const int r = threadIdx.x;
  if (r < np) {
    PxHeader *h = reinterpret_cast<PxHeader *>(peers->base[r]);
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(which == 0 ? &h->flag_fwd[me] : &h->flag_ret[me]) = epoch;
  }
}

__global__ void px_wait_kernel(PxHeader *hdr, PxInfo *info, int np, int which, uint32_t epoch, unsigned *h_err) {
  const bool ok = px_wait_flags(which == 0 ? hdr->flag_fwd : hdr->flag_ret, np, epoch);
  if (!ok && threadIdx.x == 0) {
    info->timeout = 1u;
    info->n_recv = 0u;
    info->n_send = 0u;
    reinterpret_cast<volatile unsigned *>(h_err)[1] = 1u;
    __threadfence_system();
  }
}

// phase (3): values (and leaf ids) of the received points -> the origins' return buffers
__global__ void px_return_kernel(const double *__restrict__ val, const int32_t *__restrict__ leaf,
                                 const PxInfo *__restrict__ info, const PxPeers *peers, int np, int dof,
                                 size_t off_ret_val, size_t off_ret_leaf) {
  __shared__ uint32_t s_off[kMaxRanks + 1], s_ret[kMaxRanks];
  if ((int)threadIdx.x <= np) s_off[threadIdx.x] = info->recv_off[threadIdx.x];
  if ((int)threadIdx.x < np) s_ret[threadIdx.x] = info->ret_off[threadIdx.x];
  __syncthreads();
  const size_t n = info->n_recv, total = n * dof, step = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const uint32_t k = (uint32_t)(e / dof);
    const int l = (int)(e - (size_t)k * dof);
    int s = 0;  // source rank of received point k: last s with recv_off[s] <= k
    int lo = 0, hi = np;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_off[mid] <= k) lo = mid; else hi = mid;
    }
    s = lo;
    const size_t slot = (size_t)s_ret[s] + (k - s_off[s]);
    reinterpret_cast<double *>(peers->base[s] + off_ret_val)[slot * dof + l] = val[e];
    if (leaf && l == 0) reinterpret_cast<int32_t *>(peers->base[s] + off_ret_leaf)[slot] = leaf[k];
  }
}

static ExchangeState &xs_of(tbslas_ctx *ctx) {
  if (!ctx->xs) ctx->xs = new ExchangeState();
  return *ctx->xs;
}

static void px_teardown(tbslas_ctx *ctx) {
  if (!ctx->xs) return;
  ExchangeState &x = *ctx->xs;
  for (int r = 0; r < ctx->nranks; r++) {
    if (r != ctx->rank && x.peer_base[r]) cudaIpcCloseMemHandle(x.peer_base[r]);
    x.peer_base[r] = nullptr;
  }
  x.px_ok = false;
}

// Collective: (re)allocate the mailboxes with room for `cap` points and map the peers'.
// Any failure on any rank leaves EVERY rank on the NCCL path (the outcome is all-gathered).
int px_setup(tbslas_ctx *ctx, size_t cap) {
  ExchangeState &x = xs_of(ctx);
  const int np = ctx->nranks, me = ctx->rank;
  const bool disabled = !ctx->opt.peer_exchange;
  TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  px_teardown(ctx);
  void *buf;
  TB_TRY(ws_get(ctx, WS_MISC, 128 * (np + 2), &buf));
  char *d_mine = (char *)buf, *d_all = d_mine + 128;
  // barrier: no rank frees a mailbox a peer still has mapped
  TB_NCCL(ctx, g_nccl.AllGather(d_mine, d_all, 8, ncclUint8, comm_of(ctx), ctx->stream));
  TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (x.mailbox) {
    cudaFree(x.mailbox);
    x.mailbox = nullptr;
  }
  x.lay.cap = cap;
  x.epoch = 0;
  x.info_valid = false;
  struct Msg {
    cudaIpcMemHandle_t h;
    int ok;
  } mine;
  static_assert(sizeof(Msg) <= 128, "message");
  memset(&mine, 0, sizeof(mine));
  bool ok = !disabled && cap > 0 && cudaMalloc(&x.mailbox, x.lay.bytes()) == cudaSuccess;
  if (ok) ok = cudaMemsetAsync(x.mailbox, 0, kPxHeaderBytes, ctx->stream) == cudaSuccess;  // ordered before the handle leaves
  if (ok) ok = cudaIpcGetMemHandle(&mine.h, x.mailbox) == cudaSuccess;
  if (ok && !x.d_peers) ok = cudaMalloc(&x.d_peers, sizeof(PxPeers)) == cudaSuccess;
  if (ok && !x.d_info) ok = cudaMalloc(&x.d_info, sizeof(PxInfo)) == cudaSuccess;
  if (ok && !x.h_err) {
    ok = cudaHostAlloc(&x.h_err, 16 * sizeof(unsigned), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess;
    if (ok) memset(x.h_err, 0, 16 * sizeof(unsigned));
  }
  cudaGetLastError();
  mine.ok = ok ? 1 : 0;
  std::vector<char> all(128 * np);
  for (int pass = 0; pass < 2; pass++) {  // pass 0: handles; pass 1: did every rank map every peer?
    TB_CUDA(ctx, cudaMemcpyAsync(d_mine, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
    TB_NCCL(ctx, g_nccl.AllGather(d_mine, d_all, 128, ncclUint8, comm_of(ctx), ctx->stream));
    TB_CUDA(ctx, cudaMemcpyAsync(all.data(), d_all, 128 * np, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bool all_ok = true;
    for (int r = 0; r < np; r++) all_ok = all_ok && reinterpret_cast<Msg *>(all.data() + 128 * r)->ok;
    if (!all_ok) {
      px_teardown(ctx);
      if (x.mailbox) cudaFree(x.mailbox);
      x.mailbox = nullptr;
      return TBSLAS_OK;  // NCCL path
    }
    if (pass == 1) break;
    for (int r = 0; r < np && ok; r++) {
      if (r == me) {
        x.peer_base[r] = x.mailbox;
        continue;
      }
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, reinterpret_cast<Msg *>(all.data() + 128 * r)->h,
                               cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
        break;
      }
      x.peer_base[r] = (char *)ptr;
    }
    mine.ok = ok ? 1 : 0;
  }
  PxPeers peers;
  memset(&peers, 0, sizeof(peers));
  for (int r = 0; r < np; r++) peers.base[r] = x.peer_base[r];
  TB_CUDA(ctx, cudaMemcpyAsync(x.d_peers, &peers, sizeof(peers), cudaMemcpyHostToDevice, ctx->stream));
  TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  x.px_ok = true;
  return TBSLAS_OK;
}

int px_grow(tbslas_ctx *ctx, size_t want) {
  if (ctx->xs && ctx->xs->px_ok && want > ctx->xs->lay.cap) return px_setup(ctx, want);
  return TBSLAS_OK;
}

bool comm_peer_exchange(tbslas_ctx *ctx) { return ctx->xs && ctx->xs->px_ok && ctx->exchange_mode == 1; }

// sticky errors of the peer exchange, reported at the next host synchronisation point
int comm_check(tbslas_ctx *ctx) {
  if (!ctx->xs || !ctx->xs->h_err) return TBSLAS_OK;
  volatile unsigned *e = ctx->xs->h_err;
  if (e[1]) return fail(ctx, TBSLAS_ERR_COMM, "peer exchange: a rank did not reach the barrier within 30 s");
  if (e[0]) {
    e[0] = 0;
    return fail(ctx, TBSLAS_ERR_COMM,
                "peer exchange: more outsider points than the mailbox holds (%zu per rank); raise it with "
                "tbslas_b200_comm_set_mailbox or select the NCCL exchange", ctx->xs->lay.cap);
  }
  return TBSLAS_OK;
}

// ---- peer path, steps as called by api.cu ------------------------------------------------------
// after locate: counts to everyone, offsets back
int px_begin(tbslas_tree *t, const uint32_t *send_count_dev, PxPack *pack) {
  tbslas_ctx *ctx = t->ctx;
  ExchangeState &x = xs_of(ctx);
  const int np = ctx->nranks, me = ctx->rank;
  x.epoch++;
  StageScope sc(ctx, ST_EXCHANGE, 0.0, 2);
  px_post_counts_kernel<<<1, kMaxRanks, 0, ctx->stream>>>(send_count_dev, x.d_peers, np, me, x.epoch);
  const size_t ret_cap_points = 3 * x.lay.cap / (size_t)t->dof;
  px_offsets_kernel<<<1, kMaxRanks, 0, ctx->stream>>>(reinterpret_cast<PxHeader *>(x.mailbox), x.d_info, np, me,
                                                      x.epoch, (unsigned)x.lay.cap, (unsigned)ret_cap_points,
                                                      x.h_err);
  TB_CUDA(ctx, cudaGetLastError());
  x.info_valid = true;
  pack->send_off = x.d_info->send_off;
  pack->dst_off = x.d_info->dst_off;
  pack->peer_base = x.d_peers->base;
  pack->off_recv_pos = x.lay.off_recv_pos();
  pack->skip = &x.d_info->overflow;  // followed by `timeout`
  return TBSLAS_OK;
}

// after the pack kernel: tell the owners their points are there
int px_packed(tbslas_ctx *ctx) {
  ExchangeState &x = xs_of(ctx);
  StageScope sc(ctx, ST_EXCHANGE, 0.0, 1);
  px_signal_kernel<<<1, kMaxRanks, 0, ctx->stream>>>(x.d_peers, ctx->nranks, ctx->rank, 0, x.epoch);
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

// after the insiders: evaluate what arrived, return the values, unpack mine
int px_finish(tbslas_tree *t, int bc, const uint32_t *send_idx, size_t n_local, int epilogue, double *out,
              const double *base, double alpha, int32_t *leaf_out, const GridBase *gb) {
  tbslas_ctx *ctx = t->ctx;
  ExchangeState &x = xs_of(ctx);
  const int np = ctx->nranks, me = ctx->rank, dof = t->dof;
  PxHeader *hdr = reinterpret_cast<PxHeader *>(x.mailbox);
  const size_t cap = x.lay.cap;
  void *recv_val, *recv_leaf = nullptr;
  TB_TRY(ws_get(ctx, WS_SENDVAL, sizeof(double) * dof * (cap + 1), &recv_val));
  if (leaf_out) TB_TRY(ws_get(ctx, WS_RECVLEAF, sizeof(int32_t) * (cap + 1), &recv_leaf));
  {
    StageScope sc(ctx, ST_EXCHANGE, 0.0, 1);
    px_wait_kernel<<<1, kMaxRanks, 0, ctx->stream>>>(hdr, x.d_info, np, 0, x.epoch, x.h_err);
    TB_CUDA(ctx, cudaGetLastError());
  }
  // OutEvaluation: every received point lies in this rank's Morton range
  TB_TRY(eval_received_points(t, bc, reinterpret_cast<double *>(x.mailbox + x.lay.off_recv_pos()), cap,
                              &x.d_info->n_recv, (double *)recv_val, (int32_t *)recv_leaf));
  const unsigned g = (unsigned)ctx->n_sm * 4;
  {
    StageScope sc(ctx, ST_EXCHANGE, 0.0, 3);
    px_return_kernel<<<g, 256, 0, ctx->stream>>>((const double *)recv_val, (const int32_t *)recv_leaf, x.d_info,
                                                 x.d_peers, np, dof, x.lay.off_ret_val(), x.lay.off_ret_leaf());
    px_signal_kernel<<<1, kMaxRanks, 0, ctx->stream>>>(x.d_peers, np, me, 1, x.epoch);
    px_wait_kernel<<<1, kMaxRanks, 0, ctx->stream>>>(hdr, x.d_info, np, 1, x.epoch, x.h_err);
    TB_CUDA(ctx, cudaGetLastError());
  }
  if (n_local) {
    StageScope sc(ctx, ST_UNPACK, 0.0, 1);
    const double *rv = reinterpret_cast<const double *>(x.mailbox + x.lay.off_ret_val());
    const int32_t *rl = reinterpret_cast<const int32_t *>(x.mailbox + x.lay.off_ret_leaf());
    const GridBase gbv = gb ? *gb : GridBase();
    if (epilogue == EPI_STORE)
      unpack_kernel<EPI_STORE><<<g, 256, 0, ctx->stream>>>(rv, rl, send_idx, 0, &x.d_info->n_send, dof, out, base,
                                                          alpha, leaf_out, gbv);
    else if (epilogue == EPI_AXPY)
      unpack_kernel<EPI_AXPY><<<g, 256, 0, ctx->stream>>>(rv, rl, send_idx, 0, &x.d_info->n_send, dof, out, base,
                                                         alpha, leaf_out, gbv);
    else
      unpack_kernel<EPI_AXPY_GRID><<<g, 256, 0, ctx->stream>>>(rv, rl, send_idx, 0, &x.d_info->n_send, dof, out,
                                                              base, alpha, leaf_out, gbv);
    TB_CUDA(ctx, cudaGetLastError());
  }
  return TBSLAS_OK;
}

// ---- NCCL path ---------------------------------------------------------------------------------
// Step 1 (right after locate): the nranks x nranks matrix of send counts, to pinned host
// memory; runs on the comm stream so the caller keeps enqueueing insider work.
int comm_begin_exchange(tbslas_ctx *ctx, const uint32_t *send_count_dev) {
  const int np = ctx->nranks;
  void *buf;
  TB_TRY(ws_get(ctx, WS_COUNTMAT, sizeof(uint32_t) * np * np, &buf));
  StageScope sc(ctx, ST_EXCHANGE, 0.0, 0);
  TB_TRY(chain(ctx, ctx->stream, ctx->comm_stream));
  TB_NCCL(ctx, g_nccl.AllGather(send_count_dev, buf, np, ncclUint32, comm_of(ctx), ctx->comm_stream));
  // stored by a kernel into the device-accessible pinned matrix: a cudaMemcpy would queue on the
  // device-to-host copy engine behind whatever bulk copy-out is in flight (pipelined host calls)
  publish_words_kernel<<<1, 256, 0, ctx->comm_stream>>>((const uint32_t *)buf, ctx->h_counts, np * np);
  TB_CUDA(ctx, cudaGetLastError());
  TB_CUDA(ctx, cudaEventRecord(ctx->ev_counts, ctx->comm_stream));
  return TBSLAS_OK;
}

// Step 2: the host learns the count matrix (one sync), sizes the receive buffers and posts the
// forward exchange (3 doubles per outsider, OutScatterForward) on the comm stream behind the
// pack (`ev_packed` was recorded on the main stream after it).
int comm_forward_exchange(tbslas_tree *t, const double *send_pos, bool want_leaf) {
  tbslas_ctx *ctx = t->ctx;
  const int np = ctx->nranks, me = ctx->rank;
  ExchangeState &x = xs_of(ctx);
  TB_CUDA(ctx, cudaEventSynchronize(ctx->ev_counts));
  x.n_send = x.n_recv = 0;
  for (int r = 0; r < np; r++) {
    x.send_cnt[r] = ctx->h_counts[me * np + r];
    x.recv_cnt[r] = ctx->h_counts[r * np + me];
    x.n_send += x.send_cnt[r];
    x.n_recv += x.recv_cnt[r];
  }
  if (x.send_cnt[me] || x.recv_cnt[me]) return fail(ctx, TBSLAS_ERR_COMM, "self-send in the count matrix");
  ctx->last_sent = x.n_send;
  ctx->last_recv = x.n_recv;
  x.info_valid = false;
  const int dof = t->dof;
  TB_TRY(ws_get(ctx, WS_RECV, sizeof(double) * 3 * (x.n_recv + 1), &x.recv_pos));
  TB_TRY(ws_get(ctx, WS_SENDVAL, sizeof(double) * dof * (x.n_recv + 1), &x.recv_val));
  TB_TRY(ws_get(ctx, WS_RECVVAL, sizeof(double) * dof * (x.n_send + 1), &x.ret_val));
  x.recv_leaf = x.ret_leaf = nullptr;
  if (want_leaf) {
    TB_TRY(ws_get(ctx, WS_RECVLEAF, sizeof(int32_t) * (x.n_recv + 1), &x.recv_leaf));
    TB_TRY(ws_get(ctx, WS_RETLEAF, sizeof(int32_t) * (x.n_send + 1), &x.ret_leaf));
  }
  StageScope sc(ctx, ST_EXCHANGE, (double)(24 * (x.n_send + x.n_recv)), 0);
  TB_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_packed, 0));
  TB_TRY(alltoallv(ctx, send_pos, x.send_cnt, x.recv_pos, x.recv_cnt, 24, ctx->comm_stream));
  TB_CUDA(ctx, cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
  return TBSLAS_OK;
}

// Step 3: evaluation of what arrived, reverse exchange, unpack.
int comm_finish_exchange(tbslas_tree *t, int bc, const uint32_t *send_idx, int epilogue, double *out,
                         const double *base, double alpha, int32_t *leaf_out, const GridBase *gb) {
  tbslas_ctx *ctx = t->ctx;
  ExchangeState &x = xs_of(ctx);
  const int dof = t->dof;
  TB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
  // OutEvaluation: every received point lies in this rank's Morton range
  TB_TRY(eval_received_points(t, bc, (double *)x.recv_pos, x.n_recv, nullptr, (double *)x.recv_val,
                              (int32_t *)x.recv_leaf));
  {  // reverse: dof doubles per outsider (OutScatterReverse)
    StageScope sc(ctx, ST_EXCHANGE, (double)(8 * dof * (x.n_send + x.n_recv)), 0);
    TB_TRY(alltoallv(ctx, x.recv_val, x.recv_cnt, x.ret_val, x.send_cnt, 8 * (size_t)dof, ctx->stream));
    if (leaf_out) TB_TRY(alltoallv(ctx, x.recv_leaf, x.recv_cnt, x.ret_leaf, x.send_cnt, 4, ctx->stream));
  }
  if (x.n_send) {
    StageScope sc(ctx, ST_UNPACK, (double)x.n_send, 1);
    const size_t m = x.n_send * dof;
    const unsigned grid = (unsigned)((m + 255) / 256);
    const GridBase gbv = gb ? *gb : GridBase();
    if (epilogue == EPI_STORE)
      unpack_kernel<EPI_STORE><<<grid, 256, 0, ctx->stream>>>((const double *)x.ret_val, (const int32_t *)x.ret_leaf,
                                                            send_idx, x.n_send, nullptr, dof, out, base, alpha, leaf_out, gbv);
    else if (epilogue == EPI_AXPY)
      unpack_kernel<EPI_AXPY><<<grid, 256, 0, ctx->stream>>>((const double *)x.ret_val, (const int32_t *)x.ret_leaf,
                                                           send_idx, x.n_send, nullptr, dof, out, base, alpha, leaf_out, gbv);
    else
      unpack_kernel<EPI_AXPY_GRID><<<grid, 256, 0, ctx->stream>>>((const double *)x.ret_val, (const int32_t *)x.ret_leaf,
                                                                send_idx, x.n_send, nullptr, dof, out, base, alpha,
                                                                leaf_out, gbv);
    TB_CUDA(ctx, cudaGetLastError());
  }
  return TBSLAS_OK;
}

void comm_destroy(tbslas_ctx *ctx) {
  if (ctx->xs) {
    px_teardown(ctx);
    ExchangeState &x = *ctx->xs;
    if (x.mailbox) cudaFree(x.mailbox);
    if (x.d_peers) cudaFree(x.d_peers);
    if (x.d_info) cudaFree(x.d_info);
    if (x.h_err) cudaFreeHost(x.h_err);
    delete ctx->xs;
    ctx->xs = nullptr;
  }
  if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm_of(ctx));
  ctx->nccl_comm = nullptr;
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  ctx->comm_stream = nullptr;
  for (cudaEvent_t *e : {&ctx->ev_comm, &ctx->ev_counts, &ctx->ev_packed})
    if (*e) {
      cudaEventDestroy(*e);
      *e = nullptr;
    }
}

// outsiders of the most recent evaluation, for tbslas_b200_comm_last_exchange
static int comm_last_counts(tbslas_ctx *ctx, size_t *sent, size_t *recv) {
  if (ctx->xs && ctx->xs->info_valid) {
    PxInfo h;
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    TB_CUDA(ctx, cudaMemcpy(&h, ctx->xs->d_info, sizeof(h), cudaMemcpyDeviceToHost));
    ctx->last_sent = h.n_send;
    ctx->last_recv = h.n_recv;
  }
  if (sent) *sent = ctx->last_sent;
  if (recv) *recv = ctx->last_recv;
  return TBSLAS_OK;
}

}  // namespace tb

using namespace tb;

extern "C" {

int tbslas_b200_comm_unique_id(void *id128) {
  if (!id128) return TBSLAS_ERR_INVALID;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (!nccl_load()) return TBSLAS_ERR_COMM;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return TBSLAS_ERR_COMM;
  memcpy(id128, &id, sizeof(id));
  return TBSLAS_OK;
}

int tbslas_b200_comm_init(tbslas_ctx *ctx, int nranks, int rank, const void *id128) {
  if (!ctx || !id128) return TBSLAS_ERR_INVALID;
  if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks)
    return fail(ctx, TBSLAS_ERR_INVALID, "comm_init: rank %d of %d (max %d ranks)", rank, nranks, kMaxRanks);
  if (ctx->nccl_comm) return fail(ctx, TBSLAS_ERR_INVALID, "comm_init: already initialised");
  if (!nccl_load()) return fail(ctx, TBSLAS_ERR_COMM, "%s", g_nccl_err);
  TB_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  TB_NCCL(ctx, g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm;
  ctx->nranks = nranks;
  ctx->rank = rank;
  TB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  for (cudaEvent_t *e : {&ctx->ev_comm, &ctx->ev_counts, &ctx->ev_packed})
    TB_CUDA(ctx, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  // peer-memory mailboxes (the default exchange where every rank can map every peer; otherwise,
  // on every rank alike, the NCCL all-to-all-v)
  if (nranks > 1) TB_TRY(px_setup(ctx, ctx->opt.mailbox_points));
  return TBSLAS_OK;
}

int tbslas_b200_comm_set_mailbox(tbslas_ctx *ctx, size_t points) {
  if (!ctx || !points || points > 0x7fffffffu) return TBSLAS_ERR_INVALID;
  if (ctx->nranks < 2 || !ctx->nccl_comm) return TBSLAS_OK;
  return px_setup(ctx, points);
}

int tbslas_b200_comm_set_exchange(tbslas_ctx *ctx, int mode) {
  if (!ctx || (mode != 0 && mode != 1)) return TBSLAS_ERR_INVALID;
  ctx->exchange_mode = mode;
  return TBSLAS_OK;
}

int tbslas_b200_comm_exchange_mode(tbslas_ctx *ctx, int *mode, size_t *mailbox_points) {
  if (!ctx) return TBSLAS_ERR_INVALID;
  if (mode) *mode = (ctx->nranks > 1 && comm_peer_exchange(ctx)) ? 1 : 0;
  if (mailbox_points) *mailbox_points = (ctx->xs && ctx->xs->px_ok) ? ctx->xs->lay.cap : 0;
  return TBSLAS_OK;
}

int tbslas_b200_comm_rank(tbslas_ctx *ctx, int *rank, int *nranks) {
  if (!ctx) return TBSLAS_ERR_INVALID;
  if (rank) *rank = ctx->rank;
  if (nranks) *nranks = ctx->nranks;
  return TBSLAS_OK;
}

int tbslas_b200_comm_last_exchange(tbslas_ctx *ctx, size_t *sent, size_t *received) {
  if (!ctx) return TBSLAS_ERR_INVALID;
  return comm_last_counts(ctx, sent, received);
}

}  // extern "C"

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
  // {first key, leaf count} of every rank
  unsigned long long mine[2] = {first_key, (unsigned long long)t->n_leaf};
  void *buf;
  TB_TRY(ws_get(ctx, WS_MISC, sizeof(mine) * (np + 1), &buf));
  unsigned long long *d_mine = (unsigned long long *)buf, *d_all = d_mine + 2;
  TB_CUDA(ctx, cudaMemcpyAsync(d_mine, mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
  TB_NCCL(ctx, g_nccl.AllGather(d_mine, d_all, 2, ncclUint64, comm_of(ctx), ctx->stream));
  std::vector<unsigned long long> all(2 * np);
  TB_CUDA(ctx, cudaMemcpyAsync(all.data(), d_all, sizeof(mine) * np, cudaMemcpyDeviceToHost, ctx->stream));
  TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  t->splitters.assign(np, 0);
  long long off = 0;
  for (int r = 0; r < np; r++) {
    t->splitters[r] = all[2 * r];
    if (r < ctx->rank) off += (long long)all[2 * r + 1];
  }
  // a rank without leaves owns the empty range: give it the next owner's first key so the
  // "last rank whose splitter <= key" rule never selects it
  for (int r = np - 1; r >= 0; r--)
    if (all[2 * r + 1] == 0) t->splitters[r] = (r + 1 < np) ? t->splitters[r + 1] : ~0ull;
  for (int r = 1; r < np; r++)
    if (t->splitters[r] < t->splitters[r - 1])
      return fail(ctx, TBSLAS_ERR_INVALID,
                  "ranks must own ascending Morton ranges (rank %d starts before rank %d)", r, r - 1);
  t->leaf_offset = off;
  if (!t->d_splitters) TB_CUDA(ctx, cudaMalloc(&t->d_splitters, sizeof(uint64_t) * kMaxRanks));
  TB_CUDA(ctx, cudaMemcpy(t->d_splitters, t->splitters.data(), sizeof(uint64_t) * np,
                          cudaMemcpyHostToDevice));
  return TBSLAS_OK;
}

// ---------------------------------------------------------------------------
// return path: values (and leaf ids) of my outsiders -> the caller's arrays
// ---------------------------------------------------------------------------
template <int EPI>
__global__ void unpack_kernel(const double *__restrict__ val, const int32_t *__restrict__ leaf_in,
                              const uint32_t *__restrict__ idx, size_t m, int dof,
                              double *__restrict__ out, const double *__restrict__ base, double alpha,
                              int32_t *__restrict__ leaf_out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per scalar
  if (e >= m * dof) return;
  const size_t s = e / dof;
  const int l = (int)(e - s * dof);
  const size_t i = idx[s];
  const double u = val[e];
  if (EPI == EPI_STORE) {
    out[i * dof + l] = u;
  } else {  // same expression as the eval kernel's epilogue (traj.inc:36,42)
    out[3 * i + l] = __dadd_rn(base[3 * i + l], __dmul_rn(alpha, u));
  }
  if (leaf_out && l == 0) leaf_out[i] = leaf_in[s];
}

int eval_received_points(tbslas_tree *t, int bc, double *pos, size_t n, double *out, int32_t *leaf_out);

__global__ void publish_words_kernel(const uint32_t *__restrict__ src, unsigned *host_dst, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) reinterpret_cast<volatile unsigned *>(host_dst)[i] = src[i];
  __threadfence_system();
}

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
struct ExchangeState {
  unsigned send_cnt[kMaxRanks], recv_cnt[kMaxRanks];
  size_t n_send = 0, n_recv = 0;
  void *recv_pos = nullptr, *recv_val = nullptr, *ret_val = nullptr, *recv_leaf = nullptr, *ret_leaf = nullptr;
};
static ExchangeState g_xs;  // one exchange in flight per process (one context per GPU per process)

int comm_forward_exchange(tbslas_tree *t, const double *send_pos, bool want_leaf) {
  tbslas_ctx *ctx = t->ctx;
  const int np = ctx->nranks, me = ctx->rank;
  ExchangeState &x = g_xs;
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
                         const double *base, double alpha, int32_t *leaf_out) {
  tbslas_ctx *ctx = t->ctx;
  ExchangeState &x = g_xs;
  const int dof = t->dof;
  TB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
  // OutEvaluation: every received point lies in this rank's Morton range
  TB_TRY(eval_received_points(t, bc, (double *)x.recv_pos, x.n_recv, (double *)x.recv_val, (int32_t *)x.recv_leaf));
  {  // reverse: dof doubles per outsider (OutScatterReverse)
    StageScope sc(ctx, ST_EXCHANGE, (double)(8 * dof * (x.n_send + x.n_recv)), 0);
    TB_TRY(alltoallv(ctx, x.recv_val, x.recv_cnt, x.ret_val, x.send_cnt, 8 * (size_t)dof, ctx->stream));
    if (leaf_out) TB_TRY(alltoallv(ctx, x.recv_leaf, x.recv_cnt, x.ret_leaf, x.send_cnt, 4, ctx->stream));
  }
  if (x.n_send) {
    StageScope sc(ctx, ST_UNPACK, (double)x.n_send, 1);
    const size_t m = x.n_send * dof;
    const unsigned grid = (unsigned)((m + 255) / 256);
    if (epilogue == EPI_STORE)
      unpack_kernel<EPI_STORE><<<grid, 256, 0, ctx->stream>>>((const double *)x.ret_val, (const int32_t *)x.ret_leaf,
                                                            send_idx, x.n_send, dof, out, base, alpha, leaf_out);
    else
      unpack_kernel<EPI_AXPY><<<grid, 256, 0, ctx->stream>>>((const double *)x.ret_val, (const int32_t *)x.ret_leaf,
                                                           send_idx, x.n_send, dof, out, base, alpha, leaf_out);
    TB_CUDA(ctx, cudaGetLastError());
  }
  return TBSLAS_OK;
}

void comm_destroy(tbslas_ctx *ctx) {
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
  if (sent) *sent = ctx->last_sent;
  if (received) *received = ctx->last_recv;
  return TBSLAS_OK;
}

}  // extern "C"

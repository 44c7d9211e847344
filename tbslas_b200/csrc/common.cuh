// Internal declarations shared by the translation units of libtbslas_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "tbslas_b200.h"

namespace tb {

constexpr int kMaxDepth = 15;  // pvfmm MAX_DEPTH (reference sim_config.h:40)
constexpr int kMaxRanks = 64;

// Profiling stages; names mirror the reference's pvfmm::Profile tags where one exists
// (tree_functor.h:463 LclHQSort, :585/:674 Out/InEvaluation, :568-593 OutScatter*).
enum Stage {
  ST_H2D = 0,
  ST_D2H,
  ST_LOCATE,    // wrap + key + leaf search + histogram   (reference: LclHQSort + part_indx)
  ST_BIN,       // scan + tile map + scatter of point ids  (reference: the sort's permutation)
  ST_CHEB_EVAL, // tensor-Chebyshev evaluation             (reference: In/OutEvaluation)
  ST_COMBINE,   // cubic-in-time / extrapolation / axpy    (tree_set_functor.h:66-72)
  ST_CUBIC,     // uniform-grid cubic interpolation        (fast_interp)
  ST_PACK,      // bucket outsider points by owner         (OutScatterIndex)
  ST_EXCHANGE,  // NCCL all-to-all-v                       (OutScatterForward/Reverse)
  ST_UNPACK,
  ST_GRIDPTS,   // arrival point generation                (CollectChebTreeGridPoints)
  ST_REFIT,     // values -> coefficients GEMM              (SetTreeGridValues)
  ST_TENSOR,    // velocity at the arrival grids by sum factorisation (tensor_eval.cu)
  ST_COUNT
};

struct Buf {
  void *p = nullptr;
  size_t cap = 0;
};

// Workspace slots (grown on demand, freed at finalize; the reference keeps
// function-static vectors for the same purpose, tree_functor.h:166,519-520).
enum Slot {
  WS_LEAF = 0,   // int32  [n]
  WS_RANK,       // uint32 [n]
  WS_PERM,       // uint32 [n]
  WS_COUNT,      // uint32 [L+2]
  WS_BINSTART,   // uint32 [L+2]
  WS_TILESTART,  // uint32 [L+2]
  WS_TILELEAF,   // int2   [ntile_max]
  WS_POS_A,      // double [3n]   host-call staging / RK2 state
  WS_POS_B,
  WS_POS_C,
  WS_VAL_A,      // double [n*dof] (x4 for set4)
  WS_VAL_B,
  WS_VAL_C,      // second value buffer of the host-call pipeline
  WS_LEAFOUT,    // int32 [n] staging for host leaf_idx
  WS_LEAFOUT2,
  WS_GRID,       // cubic grid staging
  WS_SEND,       // exchange buffers
  WS_RECV,
  WS_SENDVAL,
  WS_RECVVAL,
  WS_SENDIDX,
  WS_COUNTMAT,   // uint32 [nranks][nranks] send-count matrix
  WS_RECVLEAF,   // int32 leaf ids of received points / returned to the origin
  WS_RETLEAF,
  WS_MISC,
  WS_COEF,       // coefficients of a field combined in time (FieldSetFunctor / FieldExtrapFunctor)
  WS_GRIDMAP,    // int32 [n_leaf] velocity leaf that contains each leaf of the advected tree
  WS_EXC_IDX,    // uint32 [n] arrival points that need the generic evaluation
  WS_EXC_POS,
  WS_EXC_VAL,
  WS_GHOST,      // last leaf of every rank (tensor-grid path over a Morton-sharded velocity tree)
  WS_COUNT_SLOTS
};

struct Pt2Coeff {  // point-to-coefficient matrix of one degree, zero-padded to [Kp][Np]
  double *d_M = nullptr;
  int Kp = 0, Np = 0;
};

struct ProfRec {
  int stage;
  cudaEvent_t a, b;
  double units;
};

struct ExcCount {  // see tbslas_ctx::exc_cache
  uint64_t vel_hash, grid_hash;
  size_t vel_leaves, grid_leaves, leaf0, n_leaf;
  int q, periodic;
  size_t count;
};

struct ExchangeState;  // comm.cu
struct PxPack {  // where the pack kernel puts outsiders in peer-exchange mode (comm.cu)
  const uint32_t *send_off = nullptr;  // [nranks] my buckets in my own send order (device)
  const uint32_t *dst_off = nullptr;   // [nranks] where my bucket starts in the owner's receive buffer
  char *const *peer_base = nullptr;    // [nranks] mapped mailboxes (device table)
  size_t off_recv_pos = 0;
  const uint32_t *skip = nullptr;      // [2] {overflow, timeout}: non-zero = nothing may be written to a peer
};

}  // namespace tb

struct tbslas_ctx {
  int device = 0;
  int n_sm = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  tb::Buf ws[tb::WS_COUNT_SLOTS];
  // profiling
  bool prof = false;
  std::vector<tb::ProfRec> recs;
  std::vector<cudaEvent_t> ev_pool;
  double acc_ms[tb::ST_COUNT] = {0};
  double acc_units[tb::ST_COUNT] = {0};
  long long acc_launch[tb::ST_COUNT] = {0};
  long long launches = 0;
  // communicator (single rank unless comm_init was called)
  int rank = 0, nranks = 1;
  void *nccl_comm = nullptr;
  cudaStream_t comm_stream = nullptr;  // NCCL traffic that overlaps the insider evaluation
  cudaEvent_t ev_comm = nullptr, ev_counts = nullptr, ev_packed = nullptr;
  size_t last_sent = 0, last_recv = 0;  // outsiders of the most recent tree evaluation
  // host-buffer calls: H2D / compute / D2H of consecutive chunks overlap on three streams
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaStream_t copy_aux = nullptr;  // asynchronous coefficient uploads
  cudaEvent_t ev_pipe[5][2] = {};  // [in, phaseA, pos_out, phaseB, out][buffer parity]
  tb::Pt2Coeff pt2coeff[TBSLAS_MAX_CHEB_DEG + 1];
  // FieldSetFunctor / FieldExtrapFunctor over trees with one leaf list: 1 = combine the trees'
  // coefficients in time and evaluate once (default), 0 = evaluate every tree and combine the
  // values per point, in the reference's order (tree_set_functor.h:55-72)
  int time_combine = 1;
  // tree-level calls (semilag_insitu*): 1 = the velocity at the arrival grids is evaluated by sum
  // factorisation (tensor_eval.cu), 0 = point by point like any other point set
  int tensor_grid = 1;
  size_t tensor_grid_min_points = (size_t)4 << 20;  // tbslas_b200_set_tensor_grid(ctx, 2): no minimum
  size_t last_exceptions = 0;  // arrival points of the last such call that took the generic path
  // tree-level calls: 1 = the arrival points are never written to HBM when the step can rebuild them
  // from (leaf geometry, node index) where it needs them again (gridbase.cuh); 0 (default) = always
  // materialised.  Measured on C2: the tensor stage drops from 7.6 to 6.6 ms, but the index decode in
  // the second evaluation's epilogue costs 5.2 ms -- a net loss, so it is an option, not the default.
  int virtual_x = 0;
  // How many arrival points of a (grid leaf range, velocity tree) pair take the generic path is a
  // property of the two leaf lists and the boundary condition alone, so it is read back from the
  // device ONCE per pair and remembered: steps on unchanged trees never wait for the device.
  std::vector<tb::ExcCount> exc_cache;
  // pinned host scratch for small device->host reads (exchange counts: [nranks][nranks]);
  // h_exc is the tensor-grid exception count's own word
  unsigned *h_counts = nullptr;
  unsigned *h_exc = nullptr;
  // host-buffer calls: chunks per call (0 = chosen from the bytes that cross PCIe)
  int host_chunks = 0;
  // multi-rank: state of the exchange in flight (one per context; comm.cu)
  tb::ExchangeState *xs = nullptr;
  int exchange_mode = 1;  // 1: peer-memory mailboxes where available, 0: NCCL all-to-all-v
  // A/B switches of the kernels, read from the environment ONCE, at tbslas_b200_init, into the context
  // (no process-global state in the hot functions): TBSLAS_EXCHANGE_FIRST=0, TBSLAS_LOCATE_NO_BOXES=1,
  // TBSLAS_TENSOR_GENERIC=1, TBSLAS_TENSOR_DMMA=0|2, TBSLAS_TENSOR_CTAS=<n per SM>, TBSLAS_EVAL_VARIANT=1|2, TBSLAS_EXCHANGE=nccl,
  // TBSLAS_MAILBOX_POINTS=<n>
  struct Options {
    bool exchange_first = true, locate_no_boxes = false, tensor_generic = false;
    int tensor_dmma = 3;          // 0: scalar tensor-grid kernels, 2: second DMMA kernel, 3: third (the default)
    int tensor_ctas_per_sm = 0;   // grid cap of the DMMA kernels per SM (0: the kernel's own measured best)
    bool peer_exchange = true;
    int eval_variant = 0;
    size_t mailbox_points = (size_t)4 << 20;
  } opt;
};

struct tbslas_tree {
  tbslas_ctx *ctx = nullptr;
  int q = 0, dof = 0;
  size_t n_leaf = 0;       // local leaves
  size_t ncoef = 0;        // (q+1)(q+2)(q+3)/6
  size_t stride = 0;       // doubles per leaf coefficient block (dof*ncoef padded to even)
  uint64_t *d_key = nullptr;   // [n_leaf]   interleaved anchor keys, ascending
  double4 *d_geom = nullptr;   // [n_leaf+1] {cx, cy, cz, 2*2^depth}; [n_leaf] = null leaf
  uint8_t *d_depth = nullptr;  // [n_leaf]
  uint4 *d_box = nullptr;      // [n_leaf+1] {ax, ay, az, 15-depth}: integer anchor at depth 15
  uint32_t *d_cell = nullptr;  // [8^g + 2] cell table of the locate kernel: number of leaf keys <= the
                               // first key of every depth-g cell (g = (45 - cell_shift) / 3)
  int cell_shift = 45;
  uint32_t *d_pt_count = nullptr;  // [n_leaf] points located in every leaf by the last evaluation
                                   // (insiders + points received from other ranks)
  bool pt_count_valid = false;
  uint64_t struct_hash = 0;    // hash of (keys, depths): trees with equal hashes and leaf counts
                               // share their leaf list, so one locate/bin pass serves them all
  uint64_t global_hash = 0;    // the same over the leaf lists of ALL ranks (sharded trees): decisions
                               // every rank must take alike (one evaluation or one per tree) use it
  bool boxes_all = false;      // boxes_ok on every rank that holds leaves of this tree
  bool boxes_ok = false;       // leaves are aligned, non-overlapping octants: "point inside the
                               // box of leaf j" implies "j is the last leaf with key <= key(point)"
  double *d_coeff = nullptr;   // [(n_leaf+1)*stride]; block n_leaf is all zero (null leaf)
  // tbslas_b200_tree_update_coeff_async: the upload runs on the context's copy-in stream; the next
  // reader (or writer) of d_coeff on the context's stream waits for ev_coeff first (tree_coeff_ready)
  cudaEvent_t ev_coeff = nullptr;
  bool coeff_pending = false;
  size_t n_leaf_max = 0;       // most local leaves any rank holds (== n_leaf in a single-rank context)
  size_t n_leaf_global = 0;    // leaves of all ranks
  std::vector<size_t> rank_first;  // [nranks+1] global index of every rank's first leaf (sharded trees)
  uint64_t first_key = ~0ull;  // key of local leaf 0 (~0: no leaves)
  bool replicated = false;  // multi-rank context, but every rank holds the WHOLE tree: no exchange
  // Morton-range sharding (nranks > 1)
  long long leaf_offset = 0;                 // global index of local leaf 0
  std::vector<uint64_t> splitters;           // first leaf key of every rank
  uint64_t *d_splitters = nullptr;
};

namespace tb {

int fail(tbslas_ctx *ctx, int code, const char *fmt, ...);
int ws_get(tbslas_ctx *ctx, Slot s, size_t bytes, void **out);
int tree_coeff_ready(const tbslas_tree *t);
int tree_build_structure(tbslas_tree *t, const std::vector<double> &hc, const std::vector<uint8_t> &hd,
                         bool alloc_coeff);  // orders the context's stream after a pending async upload

struct StageScope {  // CUDA-event bracket of one stage on the context's stream
  tbslas_ctx *ctx;
  int idx = -1;
  StageScope(tbslas_ctx *c, int stage, double units, int n_launch);
  ~StageScope();
};

#define TB_CUDA(ctx, call)                                                          \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess)                                                          \
      return tb::fail((ctx), TBSLAS_ERR_CUDA, "%s: %s (%s:%d)", #call,              \
                      cudaGetErrorString(e_), __FILE__, __LINE__);                  \
  } while (0)

#define TB_TRY(expr)                \
  do {                              \
    int rc_ = (expr);               \
    if (rc_ != TBSLAS_OK) return rc_; \
  } while (0)

// ---- stage launchers (each enqueues on ctx->stream) ---------------------------
// locate.cu
struct LocateArgs {
  const tbslas_tree *tree;
  int periodic;
  double *pos;            // [n][3], wrapped in place when periodic
  size_t n;
  int32_t *leaf;          // [n] local leaf index, n_leaf = null leaf, <= -2: outsider of rank -2-v
  uint32_t *rank;         // [n] position inside its leaf bin (or send bucket)
  uint32_t *count;        // [n_leaf+2 (+kMaxRanks) +1] zeroed by the launcher; the last word is
                          // the evaluation kernel's chunk counter
  uint32_t *send_count;   // [nranks] (multi-rank only, zeroed by the launcher) or nullptr
  const uint32_t *n_dev = nullptr;  // != nullptr: the point count lives on the device (<= n, which
                                    // then only sizes the grid): points received through the mailbox
};
int launch_locate(tbslas_ctx *ctx, const LocateArgs &a);
// scan of bin counts, tile map, scatter of point ids
struct BinArgs {
  size_t n_leaf;          // bins 0..n_leaf (n_leaf = null leaf)
  size_t n;
  int tile_pts;           // points per evaluation tile
  const int32_t *leaf;
  const uint32_t *rank;
  const uint32_t *count;
  uint32_t *bin_start;    // [n_leaf+2]
  uint32_t *tile_start;   // [n_leaf+2]; tile_start[n_leaf+1] = number of tiles
  int2 *tile_map;         // [max_tiles] {leaf, first slot}; nullptr: not needed
  uint32_t *perm;         // [n] point ids grouped by leaf
  size_t max_tiles;
  // multi-rank only (send_count != nullptr): outsiders are packed into per-owner buckets
  const double *pos = nullptr;
  const uint32_t *send_count = nullptr;  // [nranks]
  int nranks = 1;
  double *send_pos = nullptr;            // [n_out][3]
  uint32_t *send_idx = nullptr;          // [n_out] origin index of each packed point
  const PxPack *px = nullptr;            // peer exchange: coordinates go straight to the owners
  const uint32_t *n_dev = nullptr;       // see LocateArgs
};
int launch_bin(tbslas_ctx *ctx, const BinArgs &a);

// cheb_eval_*.cu
// EPI_AXPY_GRID: as EPI_AXPY, with base[i] rebuilt from (leaf geometry, node index) instead of read
// (gridbase.cuh): tree-level steps never materialise their arrival points
enum Epilogue { EPI_STORE = 0, EPI_AXPY = 1, EPI_AXPY_GRID = 2 };
struct GridBase {
  const double4 *ggeom = nullptr;  // geometry of the grid tree's leaves, offset to the call's first leaf
  unsigned P = 1, D = 1;           // points per leaf (D^3), nodes per axis
  int periodic = 0;
  double node[TBSLAS_MAX_CHEB_DEG + 1];  // tbslas::new_nodes, 1-D
};
struct EvalArgs {
  const tbslas_tree *tree;
  const double *pos;       // [n][3]
  size_t n;                // only used for accounting
  const uint32_t *perm;
  const uint32_t *bin_start;
  const uint32_t *tile_start;  // [n_leaf+2], last entry = number of tiles
  const int2 *tile_map;
  size_t max_tiles;
  int epilogue;
  unsigned *chunk_counter; // zeroed by launch_locate; work distribution of the persistent kernel
  double *out;             // STORE: [n][dof]; AXPY: [n][3] = base + alpha*value (dof must be 3)
  const double *base;
  double alpha;
  const GridBase *grid = nullptr;  // EPI_AXPY_GRID
};
int eval_tile_points(const tbslas_tree *t);   // points per tile of the kernel that will run
bool eval_needs_tile_map(const tbslas_tree *t);  // only the one-tile-per-CTA kernels index a tile map
bool eval_supports_grid_base(const tbslas_tree *t);  // EPI_AXPY_GRID: the persistent kernel only
int launch_cheb_eval(tbslas_ctx *ctx, const EvalArgs &a);

// combine.cu
int launch_cubic_time(tbslas_ctx *ctx, const double *v4 /*[4][m]*/, size_t m, const double times[4],
                      double t, double *out, const double *base, double alpha, int axpy);
int launch_extrap(tbslas_ctx *ctx, const double *vc, const double *vp, size_t m, double *out,
                  const double *base, double alpha, int axpy);
void cubic_time_weights(const double times[4], double t, double w[4]);
int launch_combine_coeff(tbslas_ctx *ctx, const double *const c[4], const double w[4], int n_tree, size_t m,
                         double *out);
int launch_axpy(tbslas_ctx *ctx, const double *base, const double *v, double alpha, size_t m,
                double *out);
int launch_keep_counts(tbslas_ctx *ctx, const uint32_t *count, uint32_t *keep, size_t n_leaf, bool add);
int launch_leaf_fixup(tbslas_ctx *ctx, int32_t *leaf, size_t n, size_t n_leaf, long long offset);
// cubic_grid.cu
int launch_cubic_grid(tbslas_ctx *ctx, const double *grid, int n_reg, int dof, const double *pos,
                      size_t n, double *out);
// gridpts.cu
int launch_grid_points(tbslas_ctx *ctx, const tbslas_tree *t, double *out, size_t leaf0 = 0,
                       size_t n_leaf = (size_t)-1);
void new_nodes_host(int q, double *x);  // tbslas::new_nodes 1-D table (host libm)
// refit.cu
int set_pt2coeff(tbslas_ctx *ctx, int q, const double *M_host);
int launch_refit(tbslas_ctx *ctx, tbslas_tree *t, const double *vals, int point_major);
// tensor_eval.cu
int launch_tensor_grid_eval(tbslas_ctx *ctx, tbslas_tree *vel, const tbslas_tree *grid, size_t leaf0,
                            size_t n_leaf, int bc, double *x, double *out, double alpha, bool gen_points,
                            bool virtual_x);
bool tensor_grid_supports_virtual_x(const tbslas_tree *vel);
void make_grid_base(const tbslas_tree *grid, size_t leaf0, int bc, GridBase *gb);
// tailnorm.cu
int launch_tail_norm(tbslas_ctx *ctx, const tbslas_tree *t, double *out);
// peak.cu
int run_fp64_peak(tbslas_ctx *ctx, int reps, double *tflops);

}  // namespace tb

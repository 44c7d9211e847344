// C ABI of libtbslas_b200.so: context, trees and the host-side orchestration of the
// semi-Lagrangian hot path (see include/tbslas_b200.h for the contract and the
// reference file:line each entry point replaces).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost nothing unless a tool is attached

#include "common.cuh"
#include "keys.cuh"

namespace tb {

int fail(tbslas_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

int ws_get(tbslas_ctx *ctx, Slot s, size_t bytes, void **out) {
  Buf &b = ctx->ws[s];
  if (bytes > b.cap) {
    if (b.p) {
      // stream-ordered users of the old block must finish before it is released
      TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      TB_CUDA(ctx, cudaFree(b.p));
      b.p = nullptr;
      b.cap = 0;
    }
    size_t cap = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, cap);
    if (e != cudaSuccess) {
      b.p = nullptr;
      return fail(ctx, TBSLAS_ERR_NOMEM, "cudaMalloc(%zu bytes) for workspace %d: %s", cap, (int)s,
                  cudaGetErrorString(e));
    }
    b.cap = cap;
  }
  *out = b.p;
  return TBSLAS_OK;
}

// The active pvfmm::Profile::Tic tags of the reference's EvalTree (tree_functor.h) that each stage
// stands for; NVTX ranges carry "<stage> (<tag>)" so a timeline reads like the reference's profile.
static const char *kStageRefTags[ST_COUNT] = {
    "",                                   // H2D
    "",                                   // D2H
    "LclHQSort",                          // Locate     :463 (sort of the (key,index) pairs + part_indx)
    "LclHQSort",                          // Bin        :463 (the sort's permutation)
    "InEvaluation/OutEvaluation",         // ChebEval   :674 / :585
    "",                                   // Combine    (tree_set_functor.h:66-72, no tag)
    "",                                   // CubicGrid  (fast_interp, no tag)
    "OutScatterIndex",                    // Pack       :568
    "OutScatterForward/OutScatterReverse",// Exchange   :573 / :593
    "OutScatterReverse",                  // Unpack     :593-619
    "",                                   // GridPoints (CollectChebTreeGridPoints, no tag)
    "",                                   // Refit      (SetTreeGridValues, no tag)
    "InEvaluation",                       // TensorGrid (first velocity evaluation of a tree-level step)
};
static const char *kStageNvtx[ST_COUNT] = {
    "H2D", "D2H", "Locate (LclHQSort)", "Bin (LclHQSort)", "ChebEval (In/OutEvaluation)", "Combine", "CubicGrid",
    "Pack (OutScatterIndex)", "Exchange (OutScatterForward/Reverse)", "Unpack (OutScatterReverse)", "GridPoints",
    "Refit (SetTreeGridValues)", "TensorGrid (InEvaluation)"};

StageScope::StageScope(tbslas_ctx *c, int stage, double units, int n_launch) : ctx(c) {
  nvtxRangePushA(kStageNvtx[stage]);
  c->launches += n_launch;
  c->acc_launch[stage] += n_launch;
  c->acc_units[stage] += units;
  if (!c->prof) return;
  ProfRec r;
  r.stage = stage;
  r.units = units;
  for (cudaEvent_t *e : {&r.a, &r.b}) {
    if (!c->ev_pool.empty()) {
      *e = c->ev_pool.back();
      c->ev_pool.pop_back();
    } else if (cudaEventCreate(e) != cudaSuccess) {
      return;
    }
  }
  cudaEventRecord(r.a, c->stream);
  c->recs.push_back(r);
  idx = (int)c->recs.size() - 1;
}
StageScope::~StageScope() {
  if (idx >= 0) cudaEventRecord(ctx->recs[idx].b, ctx->stream);
  nvtxRangePop();
}

static int prof_collect(tbslas_ctx *ctx) {
  if (ctx->recs.empty()) return TBSLAS_OK;
  TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (ProfRec &r : ctx->recs) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) ctx->acc_ms[r.stage] += ms;
    ctx->ev_pool.push_back(r.a);
    ctx->ev_pool.push_back(r.b);
  }
  ctx->recs.clear();
  return TBSLAS_OK;
}

static const char *kStageNames[ST_COUNT] = {"H2D",     "D2H",       "Locate", "Bin",
                                            "ChebEval", "Combine",   "CubicGrid", "Pack",
                                            "Exchange", "Unpack",    "GridPoints", "Refit",
                                            "TensorGrid"};

// multi-rank pieces (comm.cu)
int comm_tree_splitters(tbslas_tree *t, uint64_t first_key);
int comm_begin_exchange(tbslas_ctx *ctx, const uint32_t *send_count_dev);
int comm_forward_exchange(tbslas_tree *t, const double *send_pos, bool want_leaf);
int comm_finish_exchange(tbslas_tree *t, int bc, const uint32_t *send_idx, int epilogue, double *out,
                         const double *base, double alpha, int32_t *leaf_out, const GridBase *gb);
bool comm_peer_exchange(tbslas_ctx *ctx);
int comm_check(tbslas_ctx *ctx);
int px_begin(tbslas_tree *t, const uint32_t *send_count_dev, PxPack *pack);
int px_packed(tbslas_ctx *ctx);
int px_finish(tbslas_tree *t, int bc, const uint32_t *send_idx, size_t n_local, int epilogue, double *out,
              const double *base, double alpha, int32_t *leaf_out, const GridBase *gb);
void comm_destroy(tbslas_ctx *ctx);
int comm_reshard(tbslas_tree *t, const size_t *new_first);

// ---------------------------------------------------------------------------
// one tree evaluation on device buffers (tbslas::EvalTree, tree_functor.h:397-690)
// ---------------------------------------------------------------------------
// `same_as`: a tree that was evaluated at these very points by the previous call (nothing in
// between): when it has the same leaf list the grouping left in the workspace is reused and
// only the evaluation kernel runs (the four snapshots of a FieldSetFunctor, the two trees of
// a FieldExtrapFunctor -- the reference sorts and searches once per tree).
static bool same_leaves(const tbslas_tree *a, const tbslas_tree *b) {
  // equal degree and block stride too: the combined-coefficient route reads every tree's blocks
  // with tree[0]'s layout (trees of different degree take the per-tree route, like the reference)
  return a && b && a->ctx == b->ctx && a->n_leaf == b->n_leaf && a->struct_hash == b->struct_hash &&
         a->global_hash == b->global_hash && a->q == b->q && a->stride == b->stride &&
         eval_tile_points(a) == eval_tile_points(b) && eval_needs_tile_map(a) == eval_needs_tile_map(b);
}

int tree_coeff_ready(const tbslas_tree *ct) {
  tbslas_tree *t = const_cast<tbslas_tree *>(ct);
  if (!t->coeff_pending) return TBSLAS_OK;
  TB_CUDA(t->ctx, cudaStreamWaitEvent(t->ctx->stream, t->ev_coeff, 0));
  t->coeff_pending = false;
  return TBSLAS_OK;
}

// `n_dev` != nullptr: the number of points is known on the device only (points that arrived through
// the peer-exchange mailbox); n is then the capacity that sizes workspaces and grids.
static int eval_local_points(tbslas_tree *t, int bc, double *pos, size_t n, int epilogue,
                             double *out, const double *base, double alpha, int32_t *leaf_out,
                             bool allow_exchange, const tbslas_tree *same_as = nullptr,
                             const uint32_t *n_dev = nullptr, const GridBase *gb = nullptr) {
  tbslas_ctx *ctx = t->ctx;
  if (n >= (size_t)0xfffffff0u)
    return fail(ctx, TBSLAS_ERR_INVALID, "n = %zu exceeds the 32-bit point index range", n);
  if (epilogue != EPI_STORE && t->dof != 3)
    return fail(ctx, TBSLAS_ERR_INVALID, "position update needs a dof-3 field (dof = %d)", t->dof);
  if (epilogue == EPI_AXPY_GRID && (!gb || !eval_supports_grid_base(t)))
    return fail(ctx, TBSLAS_ERR_INVALID, "grid-base epilogue not available for this tree");
  const int tile_pts = eval_tile_points(t);
  if (tile_pts <= 0)
    return fail(ctx, TBSLAS_ERR_UNSUPPORTED, "Chebyshev degree %d not supported (1..%d)", t->q,
                TBSLAS_MAX_CHEB_DEG);
  TB_TRY(tree_coeff_ready(t));
  const size_t max_tiles = n / tile_pts + t->n_leaf + 2;
  void *leaf, *rank, *count, *bin_start, *tile_start, *tile_map = nullptr, *perm, *send_count = nullptr;
  void *send_pos = nullptr, *send_idx = nullptr;
  TB_TRY(ws_get(ctx, WS_LEAF, sizeof(int32_t) * (n + 1), &leaf));
  TB_TRY(ws_get(ctx, WS_RANK, sizeof(uint32_t) * (n + 1), &rank));
  TB_TRY(ws_get(ctx, WS_PERM, sizeof(uint32_t) * (n + 1), &perm));
  TB_TRY(ws_get(ctx, WS_COUNT, sizeof(uint32_t) * (t->n_leaf + 2 + kMaxRanks + 1), &count));
  TB_TRY(ws_get(ctx, WS_BINSTART, sizeof(uint32_t) * (t->n_leaf + 2), &bin_start));
  TB_TRY(ws_get(ctx, WS_TILESTART, sizeof(uint32_t) * (t->n_leaf + 2), &tile_start));
  if (eval_needs_tile_map(t)) TB_TRY(ws_get(ctx, WS_TILELEAF, sizeof(int2) * max_tiles, &tile_map));
  const bool multi = allow_exchange && ctx->nranks > 1 && !t->replicated;
  const bool peer = multi && comm_peer_exchange(ctx);
  if (multi) {
    send_count = (uint32_t *)count + t->n_leaf + 2;
    // worst case: every point is an outsider (sizes must be known before the counts are)
    if (!peer) TB_TRY(ws_get(ctx, WS_SEND, sizeof(double) * 3 * (n + 1), &send_pos));
    TB_TRY(ws_get(ctx, WS_SENDIDX, sizeof(uint32_t) * (n + 1), &send_idx));
  }

  const bool exchange_first = ctx->opt.exchange_first;
  const bool reuse = !multi && !leaf_out && !n_dev && same_leaves(t, same_as) &&
                     !(same_as->ctx->nranks > 1 && !same_as->replicated);
  if (reuse) {  // the persistent evaluation kernel's work counter is the one thing to reset
    TB_CUDA(ctx, cudaMemsetAsync((uint32_t *)count + t->n_leaf + 2 + kMaxRanks, 0, sizeof(uint32_t), ctx->stream));
  } else {
    LocateArgs la;
    la.tree = t;
    la.periodic = (bc == TBSLAS_PERIODIC);
    la.pos = pos;
    la.n = n;
    la.n_dev = n_dev;
    la.leaf = (int32_t *)leaf;
    la.rank = (uint32_t *)rank;
    la.count = (uint32_t *)count;
    la.send_count = (uint32_t *)send_count;
    TB_TRY(launch_locate(ctx, la));
    PxPack pack;
    if (peer)
      TB_TRY(px_begin(t, (const uint32_t *)send_count, &pack));
    else if (multi)
      TB_TRY(comm_begin_exchange(ctx, (const uint32_t *)send_count));

    BinArgs ba;
    ba.n_leaf = t->n_leaf;
    ba.n = n;
    ba.n_dev = n_dev;
    ba.tile_pts = tile_pts;
    ba.leaf = (const int32_t *)leaf;
    ba.rank = (const uint32_t *)rank;
    ba.count = (const uint32_t *)count;
    ba.bin_start = (uint32_t *)bin_start;
    ba.tile_start = (uint32_t *)tile_start;
    ba.tile_map = (int2 *)tile_map;
    ba.perm = (uint32_t *)perm;
    ba.max_tiles = max_tiles;
    if (multi) {
      ba.pos = pos;
      ba.send_count = (const uint32_t *)send_count;
      ba.nranks = ctx->nranks;
      ba.send_pos = (double *)send_pos;
      ba.send_idx = (uint32_t *)send_idx;
      if (peer) ba.px = &pack;
    }
    TB_TRY(launch_bin(ctx, ba));
    if (peer) {
      // the outsiders are in their owners' mailboxes: say so, then evaluate the insiders while the
      // peers' points arrive
      TB_TRY(px_packed(ctx));
    } else if (multi) {
      // The persistent evaluation kernel fills every SM, so an NCCL kernel enqueued behind it on
      // another stream could not start before it drains; posting the forward exchange FIRST lets the
      // outsiders travel while the insiders are evaluated (TBSLAS_EXCHANGE_FIRST=0: old order).
      TB_CUDA(ctx, cudaEventRecord(ctx->ev_packed, ctx->stream));
      if (exchange_first) TB_TRY(comm_forward_exchange(t, (const double *)send_pos, leaf_out != nullptr));
    }
  }
  // remember the points per leaf (insiders now; the pass over received points adds its own)
  if (t->n_leaf) {
    if (!t->d_pt_count) TB_CUDA(ctx, cudaMalloc(&t->d_pt_count, sizeof(uint32_t) * t->n_leaf));
    TB_TRY(launch_keep_counts(ctx, (const uint32_t *)count, t->d_pt_count, t->n_leaf, !allow_exchange));
    t->pt_count_valid = true;
  }

  EvalArgs ea;
  ea.tree = t;
  ea.pos = pos;
  ea.n = n_dev ? 0 : n;
  ea.perm = (const uint32_t *)perm;
  ea.bin_start = (const uint32_t *)bin_start;
  ea.tile_start = (const uint32_t *)tile_start;
  ea.tile_map = (const int2 *)tile_map;
  ea.max_tiles = max_tiles;
  ea.epilogue = epilogue;
  ea.chunk_counter = (unsigned *)count + t->n_leaf + 2 + kMaxRanks;
  ea.out = out;
  ea.base = base;
  ea.alpha = alpha;
  ea.grid = gb;
  TB_TRY(launch_cheb_eval(ctx, ea));

  if (leaf_out) {
    TB_CUDA(ctx, cudaMemcpyAsync(leaf_out, leaf, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice,
                                 ctx->stream));
    TB_TRY(launch_leaf_fixup(ctx, leaf_out, n, t->n_leaf, t->leaf_offset));
  }
  if (peer) {
    TB_TRY(px_finish(t, bc, (const uint32_t *)send_idx, n, epilogue, out, base, alpha, leaf_out, gb));
  } else if (multi) {
    if (!exchange_first) TB_TRY(comm_forward_exchange(t, (const double *)send_pos, leaf_out != nullptr));
    TB_TRY(comm_finish_exchange(t, bc, (const uint32_t *)send_idx, epilogue, out, base, alpha, leaf_out, gb));
  }
  return TBSLAS_OK;
}

// used by tensor_eval.cu: the generic path for the arrival points its shortcut does not cover
int eval_tree_dev_points(tbslas_tree *t, int bc, double *pos, size_t n, double *out) {
  return eval_local_points(t, bc, pos, n, EPI_STORE, out, nullptr, 0.0, nullptr, true);
}

// used by comm.cu to evaluate points received from other ranks (all of them local)
int eval_received_points(tbslas_tree *t, int bc, double *pos, size_t n, const uint32_t *n_dev, double *out,
                         int32_t *leaf_out) {
  return eval_local_points(t, bc, pos, n, EPI_STORE, out, nullptr, 0.0, leaf_out, false, nullptr, n_dev);
}

static int eval_tree_dev(tbslas_tree *t, int bc, double *pos, size_t n, int epilogue, double *out,
                         const double *base, double alpha, int32_t *leaf_out,
                         const tbslas_tree *same_as = nullptr, const GridBase *gb = nullptr) {
  return eval_local_points(t, bc, pos, n, epilogue, out, base, alpha, leaf_out, true, same_as, nullptr, gb);
}

static int check_field(tbslas_ctx **ctx_out, const tbslas_field *f, int *dof) {
  if (!f) return TBSLAS_ERR_INVALID;
  const int nt = f->kind == TBSLAS_FIELD_STEADY ? 1 : f->kind == TBSLAS_FIELD_SET4 ? 4
                 : f->kind == TBSLAS_FIELD_EXTRAP ? 2 : 0;
  if (nt == 0 || !f->tree[0]) return TBSLAS_ERR_INVALID;
  tbslas_ctx *ctx = f->tree[0]->ctx;
  for (int i = 0; i < nt; i++) {
    if (!f->tree[i] || f->tree[i]->ctx != ctx)
      return fail(ctx, TBSLAS_ERR_INVALID, "field tree %d missing or from another context", i);
    if (f->tree[i]->dof != f->tree[0]->dof)
      return fail(ctx, TBSLAS_ERR_INVALID, "field trees disagree on dof");
  }
  *ctx_out = ctx;
  *dof = f->tree[0]->dof;
  return TBSLAS_OK;
}

// The ONE tree whose evaluation gives the field, if there is one: tree[0] of a steady field, or --
// for SET4 / EXTRAP over trees with one leaf list -- a view of tree[0] with the coefficients
// combined in time:  sum_k w_k eval(tree_k)(x) == eval(sum_k w_k coeff_k)(x), the evaluation being
// linear in the coefficients and the leaf of x the same in every tree.  4 (resp. 2) evaluations,
// their locate passes and, across ranks, their point exchanges become one; values agree with the
// reference's order of operations to rounding (~1e-15 of the field scale).  *out = nullptr when
// the field needs one evaluation per tree.
static int field_single_tree(tbslas_ctx *ctx, const tbslas_field *f, double tq, tbslas_tree *view,
                             tbslas_tree **out) {
  *out = nullptr;
  if (f->kind == TBSLAS_FIELD_STEADY) {  // time argument ignored, tree_functor.h:808-811
    *out = f->tree[0];
    return TBSLAS_OK;
  }
  const int nt = f->kind == TBSLAS_FIELD_SET4 ? 4 : 2;
  bool one_list = ctx->time_combine != 0;
  for (int i = 1; i < nt && one_list; i++)
    one_list = same_leaves(f->tree[0], f->tree[i]) && f->tree[i]->replicated == f->tree[0]->replicated;
  if (!one_list) return TBSLAS_OK;
  tbslas_tree *t0 = f->tree[0];
  double w[4] = {0, 0, 0, 0};
  if (f->kind == TBSLAS_FIELD_SET4) {
    cubic_time_weights(f->times, tq, w);
  } else {  // tree[0] = previous, tree[1] = current: 1.5*current - 0.5*previous
    w[0] = -0.5;
    w[1] = 1.5;
  }
  const size_t mc = (t0->n_leaf + 1) * t0->stride;
  void *cc;
  TB_TRY(ws_get(ctx, WS_COEF, sizeof(double) * mc, &cc));
  for (int i = 0; i < nt; i++) TB_TRY(tree_coeff_ready(f->tree[i]));
  const double *src[4] = {f->tree[0]->d_coeff, f->tree[1]->d_coeff, nt == 4 ? f->tree[2]->d_coeff : nullptr,
                          nt == 4 ? f->tree[3]->d_coeff : nullptr};
  TB_TRY(launch_combine_coeff(ctx, src, w, nt, mc, (double *)cc));
  if (t0->n_leaf && !t0->d_pt_count) TB_CUDA(ctx, cudaMalloc(&t0->d_pt_count, sizeof(uint32_t) * t0->n_leaf));
  *view = *t0;  // same leaves, keys, boxes, splitters; combined coefficients
  view->d_coeff = (double *)cc;
  *out = view;
  return TBSLAS_OK;
}

// whether field_single_tree will find one tree (no side effects: nothing is combined yet)
static bool field_is_single(tbslas_ctx *ctx, const tbslas_field *f) {
  if (f->kind == TBSLAS_FIELD_STEADY) return true;
  if (!ctx->time_combine) return false;
  const int nt = f->kind == TBSLAS_FIELD_SET4 ? 4 : 2;
  for (int i = 1; i < nt; i++)
    if (!same_leaves(f->tree[0], f->tree[i]) || f->tree[i]->replicated != f->tree[0]->replicated) return false;
  return true;
}

// out = field(pos)                       (axpy == 0)
// out = base + alpha * field(pos)        (axpy == 1; the RK2 position update; `gb`: base is not an
//                                         array but the grid points gb describes, rebuilt on the fly)
static int eval_field_dev(const tbslas_field *f, double tq, int bc, double *pos, size_t n,
                          double *out, int axpy, const double *base, double alpha,
                          const GridBase *gb = nullptr) {
  tbslas_ctx *ctx;
  int dof;
  TB_TRY(check_field(&ctx, f, &dof));
  tbslas_tree view, *one = nullptr;
  TB_TRY(field_single_tree(ctx, f, tq, &view, &one));
  if (gb && !one) return fail(ctx, TBSLAS_ERR_INVALID, "grid-base update needs a single-tree field");
  if (one) {
    TB_TRY(eval_tree_dev(one, bc, pos, n, axpy ? (gb ? EPI_AXPY_GRID : EPI_AXPY) : EPI_STORE, out, base, alpha,
                         nullptr, nullptr, gb));
    if (one == &view) f->tree[0]->pt_count_valid = view.pt_count_valid;
    return TBSLAS_OK;
  }
  const size_t m = n * dof;
  void *va;
  if (f->kind == TBSLAS_FIELD_SET4) {  // tree_set_functor.h:55-72
    TB_TRY(ws_get(ctx, WS_VAL_A, sizeof(double) * 4 * m, &va));
    double *v4 = (double *)va;
    for (int i = 0; i < 4; i++)
      TB_TRY(eval_tree_dev(f->tree[i], bc, pos, n, EPI_STORE, v4 + i * m, nullptr, 0.0, nullptr,
                           i ? f->tree[i - 1] : nullptr));
    return launch_cubic_time(ctx, v4, m, f->times, tq, out, base, alpha, axpy);
  }
  // EXTRAP: tree[0] = previous, tree[1] = current; current first (tree_extrap_functor.h:59-66)
  TB_TRY(ws_get(ctx, WS_VAL_A, sizeof(double) * 2 * m, &va));
  double *vc = (double *)va, *vp = vc + m;
  TB_TRY(eval_tree_dev(f->tree[1], bc, pos, n, EPI_STORE, vc, nullptr, 0.0, nullptr));
  TB_TRY(eval_tree_dev(f->tree[0], bc, pos, n, EPI_STORE, vp, nullptr, 0.0, nullptr, f->tree[1]));
  return launch_extrap(ctx, vc, vp, m, out, base, alpha, axpy);
}

// tbslas::ComputeTrajRK2 on device state: xsol [n][3] receives the end points, xtmp [n][3] is
// scratch.  x0 = starting points: either xsol itself (in place, as the reference works on its
// copy, traj.inc:57-64) or a separate read-only array, which saves the 48 B/point copy
// xinit -> xsol -- only when the boundary is not periodic, because the periodic wrap rewrites the
// evaluated positions in place (tree_functor.h:442-449) and the start array may be the caller's.
// `grid` != nullptr: the start points are exactly the Chebyshev grid points of leaves [leaf0, ...)
// of that tree (tree-level calls); the first velocity evaluation then runs by sum factorisation
// (tensor_eval.cu) wherever that applies.
static int traj_rk2_dev(const tbslas_field *f1, const tbslas_field *f2, int bc, double *xsol,
                        double *xtmp, size_t n, double tinit, double tfinal, int nrk,
                        const double *x0 = nullptr, const tbslas_tree *grid = nullptr, size_t leaf0 = 0,
                        bool gen_points = false, size_t n_call = 0) {
  const double tau = (tfinal - tinit) / nrk;  // traj.inc:55
  double tcur = tinit;
  tbslas_ctx *ctx = f1->tree[0]->ctx;
  if (x0 && x0 != xsol && bc == TBSLAS_PERIODIC) {
    StageScope sc(ctx, ST_COMBINE, (double)(24 * n), 0);
    TB_CUDA(ctx, cudaMemcpyAsync(xsol, x0, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice, ctx->stream));
    x0 = xsol;
  }
  for (int s = 0; s < nrk; s++) {
    double *x = (s == 0 && x0) ? const_cast<double *>(x0) : xsol;  // not written unless periodic
    // v1 = V(x, t);  xtmp = x + 0.5*tau*v1      traj.inc:33-36 (x wrapped in place if periodic)
    bool done = false, virtual_x = false;
    if (s == 0 && grid && x == xsol) {
      const size_t P = (size_t)(grid->q + 1) * (grid->q + 1) * (grid->q + 1);
      // small point sets are latency bound either way; a Morton-sharded velocity tree makes the
      // shortcut's generic pass a collective, so there the choice must not depend on this rank's n
      const bool sharded = ctx->nranks > 1 && !f1->tree[0]->replicated;
      // (`n_call`: points of the whole call when this is one chunk of it -- every chunk takes the path
      // the un-chunked call would, so the chunked host flavour returns the same bits)
      if (ctx->tensor_grid && ((n_call ? n_call : n) >= ctx->tensor_grid_min_points || sharded) && n % P == 0) {
        tbslas_tree view, *one = nullptr;
        TB_TRY(field_single_tree(ctx, f1, tcur, &view, &one));
        if (one) {
          // Virtual x: when the points are generated here and the second stage can rebuild them
          // in its epilogue (one tree, persistent kernel), the arrival points are never written.
          const tbslas_field *fs2 = f2 ? f2 : f1;
          const bool virt = gen_points && ctx->virtual_x && tensor_grid_supports_virtual_x(one) &&
                            field_is_single(ctx, fs2) && eval_supports_grid_base(fs2->tree[0]) &&
                            fs2->tree[0]->q == grid->q;  // (the epilogue decodes node indices at the kernel's degree)
          const int rc = launch_tensor_grid_eval(ctx, one, grid, leaf0, n / P, bc, x, xtmp, 0.5 * tau, gen_points, virt);
          if (rc == TBSLAS_OK) {
            done = true;
            virtual_x = virt;
          } else if (rc != TBSLAS_ERR_UNSUPPORTED) {
            return rc;
          }
        }
      }
      // `gen_points`: the start points are still to be written (CollectChebTreeGridPoints)
      if (!done && gen_points && n) TB_TRY(launch_grid_points(ctx, grid, x, leaf0, n / P));
    }
    if (!done) TB_TRY(eval_field_dev(f1, tcur, bc, x, n, xtmp, 1, x, 0.5 * tau));
    // v2 = V(xtmp, t + tau/2);  x = x + tau*v2  traj.inc:40-42
    if (virtual_x) {
      GridBase gb;
      make_grid_base(grid, leaf0, bc, &gb);
      TB_TRY(eval_field_dev(f2 ? f2 : f1, tcur + 0.5 * tau, bc, xtmp, n, xsol, 1, nullptr, tau, &gb));
    } else {
      TB_TRY(eval_field_dev(f2 ? f2 : f1, tcur + 0.5 * tau, bc, xtmp, n, xsol, 1, x, tau));
    }
    tcur = tcur + tau;
  }
  return TBSLAS_OK;
}

struct HostIO {  // staging of caller buffers that live in host memory
  tbslas_ctx *ctx;
  int mem;
  int h2d(Slot s, const void *src, size_t bytes, void **dev) {
    if (mem == TBSLAS_MEM_DEVICE) {
      *dev = const_cast<void *>(src);
      return TBSLAS_OK;
    }
    TB_TRY(ws_get(ctx, s, bytes, dev));
    StageScope sc(ctx, ST_H2D, (double)bytes, 0);
    TB_CUDA(ctx, cudaMemcpyAsync(*dev, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return TBSLAS_OK;
  }
  int out_buf(Slot s, void *dst, size_t bytes, void **dev) {
    if (mem == TBSLAS_MEM_DEVICE) {
      *dev = dst;
      return TBSLAS_OK;
    }
    return ws_get(ctx, s, bytes, dev);
  }
  int d2h(void *dst, const void *dev, size_t bytes) {
    if (mem == TBSLAS_MEM_DEVICE || !dst) return TBSLAS_OK;
    StageScope sc(ctx, ST_D2H, (double)bytes, 0);
    TB_CUDA(ctx, cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return TBSLAS_OK;
  }
  int finish() {
    if (mem == TBSLAS_MEM_DEVICE) return TBSLAS_OK;
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return comm_check(ctx);
  }
};


// ---------------------------------------------------------------------------
// Host-buffer calls, pipelined: the points are cut into chunks and the H2D copy of chunk
// c+1, the kernels of chunk c and the D2H copy of chunk c-1 run concurrently (copy-in
// stream, the context's stream, copy-out stream; two buffer sets).  PCIe is full duplex and
// a B200 has separate copy engines per direction, so a host call costs about
// max(H2D, compute, D2H) instead of their sum.  Every chunk is a complete, independent
// evaluation (the result of a point depends only on that point), so results are identical
// to the one-shot path bit for bit.  In a multi-rank context every chunk is collective:
// the chunk count must not depend on this rank's n.
// ---------------------------------------------------------------------------
struct PipeBufs {
  double *pos, *tmp, *val;
  int32_t *leaf;
};
struct PipeSpec {
  const double *h_pos = nullptr;  // [n][3] input
  double *h_pos_a = nullptr;      // positions after phase A (trajectory end points), optional
  double *h_pos_b = nullptr;      // positions after phase B (periodic wrap written back), optional
  double *h_val = nullptr;        // [n][val_dof], optional
  int val_dof = 0;
  int32_t *h_leaf = nullptr;      // [n], optional
  bool need_tmp = false;
  size_t unit = 1;                // chunks are whole multiples of `unit` points (a leaf's grid)
  size_t n_collective = 0;        // multi-rank: a point count every rank knows (0: none)
};
enum { EV_IN = 0, EV_A, EV_POSOUT, EV_B, EV_OUT };

// Chunks of a host-buffer call.  What is exposed is the copy-in of the first chunk (calls with host
// input) and the copy-out of the last one; what a chunk costs is the ramp-up and tail of every
// kernel it launches (measured on C2: about 1 ms per chunk, profiles/r02_e2e_chunk_sweep.json).
//  * calls with host input are bound by the copy-in: uniform chunks of about 4 Mi points, <= 16;
//  * tree-level calls (arrival points generated in HBM, only values leave) use chunks of
//    geometrically DEcreasing size, ratio kChunkRatio.  The copy-out of chunk c then runs under the
//    kernels of chunk c+1, which is kChunkRatio times as large: that fits as long as copying a point's
//    values out takes less than kChunkRatio of computing it -- 8 B at ~50 GB/s = 0.16 ns against
//    0.32 ns per point of the q = 8..14 step, i.e. 0.5 (measured, also with every GPU of the box
//    copying at once: 47 GB/s) -- so only the last, small copy is exposed and about log(n / 8 Mi)
//    chunks do what dozens of uniform ones would.
// `n_collective`: in a multi-rank context every chunk is a collective evaluation, so the count must be
// the same on every rank: it is derived from a size all ranks know (the largest shard's arrival points
// for tree-level calls) or is the fixed 8.
constexpr double kChunkRatio = 0.62;
static double chunk_cum(int K, int c) {  // fraction of the units in chunks 0..c-1
  return (1.0 - pow(kChunkRatio, c)) / (1.0 - pow(kChunkRatio, K));
}
// (Four or more GPUs of one host share its PCIe root complexes: measured 21.8-26 GB/s per GPU with
// four copying at once against 53 GB/s for one alone, bench line e2e.pcie_d2h_GBps_slowest_rank.  The
// copy-out is then the bottleneck whatever the schedule; 16 uniform chunks were measured there and
// lost to the geometric schedule, 43.1 against 37.2 ms per step, because every chunk of a multi-rank
// call is a collective evaluation.)
static int pipe_chunks(tbslas_ctx *ctx, size_t n, bool has_input, size_t n_collective, bool *geometric) {
  *geometric = false;
  if (ctx->host_chunks > 0) return ctx->host_chunks;
  if (ctx->nranks > 1) {
    if (!n_collective) return 8;
    n = n_collective;
  }
  if (has_input) {
    const size_t per = (size_t)4 << 20, k = (n + per / 2) / per;
    return (int)(k < 1 ? 1 : (k > 16 ? 16 : k));
  }
  *geometric = true;
  int k = 1;  // fewest chunks whose last one holds at most 8 Mi points
  while (k < 12 && (double)n * (1.0 - chunk_cum(k, k - 1)) > (double)((size_t)8 << 20)) k++;
  return k;
}

// first unit of chunk c of K over U units
static size_t chunk_first_unit(size_t U, int K, int c, bool geometric) {
  if (c <= 0) return 0;
  if (c >= K) return U;
  if (!geometric) return (size_t)(((unsigned __int128)U * (unsigned)c) / (unsigned)K);
  const size_t f = (size_t)((double)U * chunk_cum(K, c));
  return f > U ? U : f;
}

template <class FA, class FB>
static int run_host_pipeline(tbslas_ctx *ctx, const PipeSpec &sp, size_t n, FA phase_a, FB phase_b) {
  const bool has_input = sp.h_pos != nullptr;
  bool halving = false;  // (chunks of geometrically decreasing size)
  const int K = pipe_chunks(ctx, n, has_input, sp.n_collective, &halving);
  const size_t unit = sp.unit ? sp.unit : 1;
  const size_t U = (n + unit - 1) / unit;
  size_t chunk = 0;  // largest chunk, in points
  for (int c = 0; c < K; c++) {
    const size_t u = chunk_first_unit(U, K, c + 1, halving) - chunk_first_unit(U, K, c, halving);
    if (u * unit > chunk) chunk = u * unit;
  }
  PipeBufs bufs[2] = {};
  const Slot pos_slot[2] = {WS_POS_A, WS_POS_C}, val_slot[2] = {WS_VAL_B, WS_VAL_C},
             leaf_slot[2] = {WS_LEAFOUT, WS_LEAFOUT2};
  void *tmp = nullptr;
  if (sp.need_tmp) TB_TRY(ws_get(ctx, WS_POS_B, sizeof(double) * 3 * (chunk + 1), &tmp));
  for (int b = 0; b < (K > 1 ? 2 : 1); b++) {
    void *p, *v = nullptr, *l = nullptr;
    TB_TRY(ws_get(ctx, pos_slot[b], sizeof(double) * 3 * (chunk + 1), &p));
    if (sp.h_val) TB_TRY(ws_get(ctx, val_slot[b], sizeof(double) * sp.val_dof * (chunk + 1), &v));
    if (sp.h_leaf) TB_TRY(ws_get(ctx, leaf_slot[b], sizeof(int32_t) * (chunk + 1), &l));
    bufs[b] = PipeBufs{(double *)p, (double *)tmp, (double *)v, (int32_t *)l};
  }
  cudaStream_t s_in = ctx->copy_in, s_out = ctx->copy_out, s_run = ctx->stream;
  auto ev = [&](int what, int b) { return ctx->ev_pipe[what][b]; };
  // order the side streams after whatever the caller enqueued before this call
  TB_CUDA(ctx, cudaEventRecord(ev(EV_OUT, 0), s_run));
  TB_CUDA(ctx, cudaEventRecord(ev(EV_OUT, 1), s_run));
  TB_CUDA(ctx, cudaEventRecord(ev(EV_B, 0), s_run));
  TB_CUDA(ctx, cudaEventRecord(ev(EV_B, 1), s_run));
  for (int c = 0; c < K; c++) {
    const int b = c & 1;
    size_t off = chunk_first_unit(U, K, c, halving) * unit, end = chunk_first_unit(U, K, c + 1, halving) * unit;
    if (off > n) off = n;
    if (end > n) end = n;
    const size_t m = end - off;
    PipeBufs &B = bufs[b];
    if (has_input) {
      // ---- copy in (buffer b is free once chunk c-2 has been computed and copied out)
      TB_CUDA(ctx, cudaStreamWaitEvent(s_in, ev(EV_B, b), 0));
      TB_CUDA(ctx, cudaStreamWaitEvent(s_in, ev(EV_OUT, b), 0));
      if (m) {
        TB_CUDA(ctx, cudaMemcpyAsync(B.pos, sp.h_pos + 3 * off, sizeof(double) * 3 * m,
                                     cudaMemcpyHostToDevice, s_in));
        ctx->acc_units[ST_H2D] += (double)(24 * m);
      }
      TB_CUDA(ctx, cudaEventRecord(ev(EV_IN, b), s_in));
      TB_CUDA(ctx, cudaStreamWaitEvent(s_run, ev(EV_IN, b), 0));
    } else {
      // nothing to copy in: the copy-in stream stays out of it (an asynchronous coefficient upload may
      // be running there); buffer b is free once chunk c-2 has been copied out
      TB_CUDA(ctx, cudaStreamWaitEvent(s_run, ev(EV_OUT, b), 0));
    }
    // ---- compute
    TB_TRY(phase_a(B, m, off));
    if (sp.h_pos_a) {
      TB_CUDA(ctx, cudaEventRecord(ev(EV_A, b), s_run));
      TB_CUDA(ctx, cudaStreamWaitEvent(s_out, ev(EV_A, b), 0));
      if (m) TB_CUDA(ctx, cudaMemcpyAsync(sp.h_pos_a + 3 * off, B.pos, sizeof(double) * 3 * m,
                                          cudaMemcpyDeviceToHost, s_out));
      TB_CUDA(ctx, cudaEventRecord(ev(EV_POSOUT, b), s_out));
      TB_CUDA(ctx, cudaStreamWaitEvent(s_run, ev(EV_POSOUT, b), 0));  // phase B may wrap B.pos
      ctx->acc_units[ST_D2H] += (double)(24 * m);
    }
    TB_TRY(phase_b(B, m, off));
    TB_CUDA(ctx, cudaEventRecord(ev(EV_B, b), s_run));
    // ---- copy out
    TB_CUDA(ctx, cudaStreamWaitEvent(s_out, ev(EV_B, b), 0));
    if (m && sp.h_val) {
      TB_CUDA(ctx, cudaMemcpyAsync(sp.h_val + sp.val_dof * off, B.val, sizeof(double) * sp.val_dof * m,
                                   cudaMemcpyDeviceToHost, s_out));
      ctx->acc_units[ST_D2H] += (double)(8 * sp.val_dof * m);
    }
    if (m && sp.h_leaf)
      TB_CUDA(ctx, cudaMemcpyAsync(sp.h_leaf + off, B.leaf, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, s_out));
    if (m && sp.h_pos_b)
      TB_CUDA(ctx, cudaMemcpyAsync(sp.h_pos_b + 3 * off, B.pos, sizeof(double) * 3 * m,
                                   cudaMemcpyDeviceToHost, s_out));
    TB_CUDA(ctx, cudaEventRecord(ev(EV_OUT, b), s_out));
  }
  TB_CUDA(ctx, cudaStreamSynchronize(s_out));
  TB_CUDA(ctx, cudaStreamSynchronize(s_run));
  return comm_check(ctx);
}

// Everything a tree derives from its leaf list (host coordinates + depths, Morton order): keys,
// geometry, integer boxes, the locate kernel's cell table, hashes.  Replaces the arrays the tree
// holds (used by tree_create and, after leaves moved between ranks, by tree_reshard); the
// coefficient block is (re)allocated and zeroed only when `alloc_coeff`.
int tree_build_structure(tbslas_tree *t, const std::vector<double> &hc, const std::vector<uint8_t> &hd,
                         bool alloc_coeff) {
  tbslas_ctx *ctx = t->ctx;
  const size_t n_leaf = hd.size();
  std::vector<uint64_t> hk(n_leaf);
  std::vector<double4> hg(n_leaf + 1);
  std::vector<uint4> hb(n_leaf + 1);
  bool boxes_ok = true;
  for (size_t j = 0; j < n_leaf; j++) {
    if (hd[j] > kMaxDepth) return fail(ctx, TBSLAS_ERR_INVALID, "leaf %zu: depth %d > 15", j, hd[j]);
    hk[j] = leaf_key(hc[3 * j], hc[3 * j + 1], hc[3 * j + 2]);
    if (hk[j] == ~0ull) return fail(ctx, TBSLAS_ERR_INVALID, "leaf %zu: corner outside [0,1)^3", j);
    if (j && !(hk[j - 1] < hk[j]))
      return fail(ctx, TBSLAS_ERR_INVALID, "leaves must be in strictly ascending Morton order (leaf %zu)", j);
    // (x - c) * 2.0 * s, s = 2^depth (tree_functor.h:285-293): 2*s is exact
    hg[j] = make_double4(hc[3 * j], hc[3 * j + 1], hc[3 * j + 2], 2.0 * (double)(1ull << hd[j]));
    // integer box of the octant (locate fast path): usable when every leaf is aligned to its
    // own depth and ends before the next leaf begins
    const unsigned sh = (unsigned)(kMaxDepth - hd[j]);
    hb[j] = make_uint4((unsigned)floor(hc[3 * j] * 32768.0), (unsigned)floor(hc[3 * j + 1] * 32768.0),
                       (unsigned)floor(hc[3 * j + 2] * 32768.0), sh);
    const unsigned low = (1u << sh) - 1u;
    if ((hb[j].x | hb[j].y | hb[j].z) & low) boxes_ok = false;
    if (j && hk[j] - hk[j - 1] < (1ull << (3 * hb[j - 1].w))) boxes_ok = false;
  }
  hg[n_leaf] = make_double4(0, 0, 0, 2.0);  // null leaf: zero coefficients
  hb[n_leaf] = make_uint4(0, 0, 0, 0);
  TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(t->d_key);
  cudaFree(t->d_geom);
  cudaFree(t->d_depth);
  cudaFree(t->d_box);
  cudaFree(t->d_cell);
  cudaFree(t->d_pt_count);
  t->d_key = nullptr, t->d_geom = nullptr, t->d_depth = nullptr, t->d_box = nullptr, t->d_cell = nullptr;
  t->d_pt_count = nullptr;
  t->pt_count_valid = false;
  t->n_leaf = n_leaf;
  TB_CUDA(ctx, cudaMalloc(&t->d_key, sizeof(uint64_t) * (n_leaf + 1)));
  TB_CUDA(ctx, cudaMalloc(&t->d_geom, sizeof(double4) * (n_leaf + 1)));
  TB_CUDA(ctx, cudaMalloc(&t->d_depth, n_leaf + 1));
  TB_CUDA(ctx, cudaMalloc(&t->d_box, sizeof(uint4) * (n_leaf + 1)));
  TB_CUDA(ctx, cudaMemcpyAsync(t->d_box, hb.data(), sizeof(uint4) * (n_leaf + 1), cudaMemcpyHostToDevice, ctx->stream));
  t->boxes_ok = boxes_ok;
  t->boxes_all = boxes_ok;
  std::vector<uint32_t> cell;
  {  // cell table: depth g with about two cells per leaf, 1 <= g <= 6 (1 MiB)
    int g = 1;
    while (g < 6 && ((size_t)1 << (3 * g)) < 2 * n_leaf) g++;
    const size_t n_cell = (size_t)1 << (3 * g);
    t->cell_shift = 3 * (kMaxDepth - g);
    cell.assign(n_cell + 2, 0u);
    size_t j = 0;
    for (size_t c = 0; c < n_cell; c++) {
      const uint64_t first = (uint64_t)c << t->cell_shift;
      while (j < n_leaf && hk[j] <= first) j++;
      cell[c] = (uint32_t)j;
    }
    cell[n_cell] = cell[n_cell + 1] = (uint32_t)n_leaf;
    TB_CUDA(ctx, cudaMalloc(&t->d_cell, sizeof(uint32_t) * cell.size()));
    TB_CUDA(ctx, cudaMemcpyAsync(t->d_cell, cell.data(), sizeof(uint32_t) * cell.size(), cudaMemcpyHostToDevice, ctx->stream));
  }
  {
    uint64_t h = 1469598103934665603ull;  // FNV-1a over the leaf keys and depths
    for (size_t j = 0; j < n_leaf; j++) {
      h = (h ^ hk[j]) * 1099511628211ull;
      h = (h ^ hd[j]) * 1099511628211ull;
    }
    t->struct_hash = h;
    t->global_hash = h;
  }
  t->n_leaf_max = n_leaf;
  t->n_leaf_global = n_leaf;
  t->leaf_offset = 0;
  t->rank_first.assign({(size_t)0, n_leaf});
  TB_CUDA(ctx, cudaMemcpyAsync(t->d_key, hk.data(), sizeof(uint64_t) * n_leaf, cudaMemcpyHostToDevice, ctx->stream));
  TB_CUDA(ctx, cudaMemcpyAsync(t->d_geom, hg.data(), sizeof(double4) * (n_leaf + 1), cudaMemcpyHostToDevice, ctx->stream));
  TB_CUDA(ctx, cudaMemcpyAsync(t->d_depth, hd.data(), n_leaf, cudaMemcpyHostToDevice, ctx->stream));
  if (alloc_coeff) {
    cudaFree(t->d_coeff);
    t->d_coeff = nullptr;
    TB_CUDA(ctx, cudaMalloc(&t->d_coeff, sizeof(double) * t->stride * (n_leaf + 1)));
    // on the context's stream: the coefficient upload that follows must not overtake it
    TB_CUDA(ctx, cudaMemsetAsync(t->d_coeff, 0, sizeof(double) * t->stride * (n_leaf + 1), ctx->stream));
  }
  t->first_key = n_leaf ? hk[0] : ~0ull;
  t->splitters.assign(1, t->first_key);
  // everything above went through the context's stream (a copy on the legacy default stream is not
  // ordered with a non-blocking stream); the host vectors must outlive the copies
  TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TBSLAS_OK;
}

}  // namespace tb

using namespace tb;

// ===========================================================================
extern "C" {

const char *tbslas_b200_version(void) { return "tbslas_b200 0.1 (sm_100a)"; }

int tbslas_b200_init(int device, tbslas_ctx **out) {
  if (!out) return TBSLAS_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return TBSLAS_ERR_CUDA;
  if (device < 0 || device >= ndev) return TBSLAS_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return TBSLAS_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return TBSLAS_ERR_CUDA;
  if (prop.major != 10) return TBSLAS_ERR_CUDA;  // sm_100a kernels only; no fallback
  tbslas_ctx *ctx = new tbslas_ctx();
  {  // A/B switches: the environment is read here and nowhere else
    auto flag = [](const char *name, bool dflt) {
      const char *e = getenv(name);
      return e ? atoi(e) != 0 : dflt;
    };
    ctx->opt.exchange_first = flag("TBSLAS_EXCHANGE_FIRST", true);
    ctx->opt.locate_no_boxes = flag("TBSLAS_LOCATE_NO_BOXES", false);
    ctx->opt.tensor_generic = flag("TBSLAS_TENSOR_GENERIC", false);
    if (const char *e = getenv("TBSLAS_TENSOR_DMMA")) ctx->opt.tensor_dmma = atoi(e);
    if (const char *e = getenv("TBSLAS_VIRTUAL_X")) ctx->virtual_x = atoi(e) != 0;  // = tbslas_b200_set_virtual_arrival_points
    if (const char *e = getenv("TBSLAS_HOST_CHUNKS")) ctx->host_chunks = atoi(e);     // = tbslas_b200_set_host_chunks
    if (const char *e = getenv("TBSLAS_TENSOR_CTAS")) ctx->opt.tensor_ctas_per_sm = atoi(e) > 0 ? atoi(e) : 0;
    if (const char *e = getenv("TBSLAS_EVAL_VARIANT")) ctx->opt.eval_variant = atoi(e);
    if (const char *e = getenv("TBSLAS_EXCHANGE")) ctx->opt.peer_exchange = strcmp(e, "nccl") != 0;
    if (const char *e = getenv("TBSLAS_MAILBOX_POINTS")) ctx->opt.mailbox_points = (size_t)strtoull(e, nullptr, 10);
  }
  ctx->device = device;
  ctx->n_sm = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return TBSLAS_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  bool ok = cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->copy_aux, cudaStreamNonBlocking) == cudaSuccess;
  for (auto &pair : ctx->ev_pipe)
    for (cudaEvent_t &e : pair) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
  // device-writable pinned words: the exchange's count matrix (comm.cu) and, separately, the
  // tensor-grid exception count (tensor_eval.cu)
  const bool pinned = ok &&
      cudaHostAlloc(&ctx->h_counts, sizeof(unsigned) * kMaxRanks * kMaxRanks,
                    cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
      cudaHostAlloc(&ctx->h_exc, sizeof(unsigned) * 16, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess;
  if (!ok || !pinned) {
    const int rc = ok ? TBSLAS_ERR_NOMEM : TBSLAS_ERR_CUDA;
    cudaGetLastError();
    tbslas_b200_finalize(ctx);
    return rc;
  }
  *out = ctx;
  return TBSLAS_OK;
}

int tbslas_b200_finalize(tbslas_ctx *ctx) {
  if (!ctx) return TBSLAS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  comm_destroy(ctx);
  for (Buf &b : ctx->ws)
    if (b.p) cudaFree(b.p);
  for (ProfRec &r : ctx->recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
  if (ctx->h_exc) cudaFreeHost(ctx->h_exc);
  for (Pt2Coeff &m : ctx->pt2coeff)
    if (m.d_M) cudaFree(m.d_M);
  for (auto &pair : ctx->ev_pipe)
    for (cudaEvent_t &e : pair)
      if (e) cudaEventDestroy(e);
  if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
  if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
  if (ctx->copy_aux) cudaStreamDestroy(ctx->copy_aux);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return TBSLAS_OK;
}

int tbslas_b200_set_stream(tbslas_ctx *ctx, void *s) {
  if (!ctx) return TBSLAS_ERR_INVALID;
  TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  return TBSLAS_OK;
}

int tbslas_b200_set_time_combine(tbslas_ctx *ctx, int mode) {
  if (!ctx || (mode != 0 && mode != 1)) return TBSLAS_ERR_INVALID;
  ctx->time_combine = mode;
  return TBSLAS_OK;
}

int tbslas_b200_last_grid_exceptions(tbslas_ctx *ctx, size_t *n) {
  if (!ctx || !n) return TBSLAS_ERR_INVALID;
  *n = ctx->last_exceptions;
  return TBSLAS_OK;
}

int tbslas_b200_set_virtual_arrival_points(tbslas_ctx *ctx, int on) {
  if (!ctx || (on != 0 && on != 1)) return TBSLAS_ERR_INVALID;
  ctx->virtual_x = on;
  return TBSLAS_OK;
}

int tbslas_b200_set_host_chunks(tbslas_ctx *ctx, int chunks) {
  if (!ctx || chunks < 0 || chunks > 1024) return TBSLAS_ERR_INVALID;
  ctx->host_chunks = chunks;
  return TBSLAS_OK;
}

int tbslas_b200_set_tensor_grid(tbslas_ctx *ctx, int mode) {
  if (!ctx || mode < 0 || mode > 2) return TBSLAS_ERR_INVALID;
  ctx->tensor_grid = mode != 0;
  ctx->tensor_grid_min_points = mode == 2 ? 1 : (size_t)4 << 20;
  return TBSLAS_OK;
}

int tbslas_b200_synchronize(tbslas_ctx *ctx) {
  if (!ctx) return TBSLAS_ERR_INVALID;
  TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return comm_check(ctx);
}

const char *tbslas_b200_last_error(tbslas_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// ---------------------------------------------------------------- trees
static int tree_create_impl(tbslas_ctx *ctx, int q, int dof, size_t n_leaf, const double *coord,
                            const uint8_t *depth, const double *coeff, int mem, bool replicated,
                            tbslas_tree **out);

int tbslas_b200_tree_create(tbslas_ctx *ctx, int q, int dof, size_t n_leaf, const double *coord,
                            const uint8_t *depth, const double *coeff, int mem, tbslas_tree **out) {
  return tree_create_impl(ctx, q, dof, n_leaf, coord, depth, coeff, mem, false, out);
}

int tbslas_b200_tree_create_replicated(tbslas_ctx *ctx, int q, int dof, size_t n_leaf, const double *coord,
                                       const uint8_t *depth, const double *coeff, int mem,
                                       tbslas_tree **out) {
  return tree_create_impl(ctx, q, dof, n_leaf, coord, depth, coeff, mem, true, out);
}

static int tree_create_impl(tbslas_ctx *ctx, int q, int dof, size_t n_leaf, const double *coord,
                            const uint8_t *depth, const double *coeff, int mem, bool replicated,
                            tbslas_tree **out) {
  if (!ctx || !out) return TBSLAS_ERR_INVALID;
  *out = nullptr;
  if (q < 1 || q > TBSLAS_MAX_CHEB_DEG)
    return fail(ctx, TBSLAS_ERR_UNSUPPORTED, "Chebyshev degree %d not supported (1..%d)", q,
                TBSLAS_MAX_CHEB_DEG);
  // a rank of a multi-rank context may own no leaf of this tree (empty Morton range)
  const size_t min_leaf = (ctx->nranks > 1 && !replicated) ? 0 : 1;
  if (dof < 1 || dof > 16 || n_leaf < min_leaf || n_leaf > 0x7ffffff0u ||
      (n_leaf && (!coord || !depth || !coeff)))
    return fail(ctx, TBSLAS_ERR_INVALID, "tree_create: bad argument (dof=%d, n_leaf=%zu)", dof, n_leaf);
  TB_CUDA(ctx, cudaSetDevice(ctx->device));
  // geometry and keys are built on the host (n_leaf is small next to the points)
  std::vector<double> hc(3 * n_leaf);
  std::vector<uint8_t> hd(n_leaf);
  if (mem == TBSLAS_MEM_DEVICE) {
    TB_CUDA(ctx, cudaMemcpy(hc.data(), coord, sizeof(double) * 3 * n_leaf, cudaMemcpyDeviceToHost));
    TB_CUDA(ctx, cudaMemcpy(hd.data(), depth, n_leaf, cudaMemcpyDeviceToHost));
  } else {
    memcpy(hc.data(), coord, sizeof(double) * 3 * n_leaf);
    memcpy(hd.data(), depth, n_leaf);
  }
  tbslas_tree *t = new tbslas_tree();
  t->ctx = ctx;
  t->q = q;
  t->dof = dof;
  t->replicated = replicated;
  t->ncoef = (size_t)(q + 1) * (q + 2) * (q + 3) / 6;
  const size_t ncoef_pad = t->ncoef + (t->ncoef & 1);  // 16-byte rows for TMA / LDS.128
  t->stride = ncoef_pad * dof;
  auto bail = [&](int rc) {
    tbslas_b200_tree_destroy(t);
    return rc;
  };
  int rc = tree_build_structure(t, hc, hd, true);
  if (rc != TBSLAS_OK) return bail(rc);
  rc = n_leaf ? tbslas_b200_tree_update_coeff(t, coeff, mem) : TBSLAS_OK;
  if (rc != TBSLAS_OK) return bail(rc);
  if (ctx->nranks > 1 && !replicated) {
    rc = comm_tree_splitters(t, t->first_key);
    if (rc != TBSLAS_OK) return bail(rc);
  }
  *out = t;
  return TBSLAS_OK;
}

// [leaf][dof][Ncoef] -> rows padded to an even number of doubles (one flat copy when Ncoef is
// even already: a pitched copy is one DMA descriptor per 5 KB row)
static int copy_coeff_in(tbslas_tree *t, const double *coeff, int mem, cudaStream_t s) {
  tbslas_ctx *ctx = t->ctx;
  const size_t ncoef_pad = t->stride / t->dof;
  const cudaMemcpyKind kind = mem == TBSLAS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (ncoef_pad == t->ncoef)
    TB_CUDA(ctx, cudaMemcpyAsync(t->d_coeff, coeff, sizeof(double) * t->ncoef * t->n_leaf * t->dof, kind, s));
  else
    TB_CUDA(ctx, cudaMemcpy2DAsync(t->d_coeff, ncoef_pad * sizeof(double), coeff, t->ncoef * sizeof(double),
                                   t->ncoef * sizeof(double), t->n_leaf * t->dof, kind, s));
  return TBSLAS_OK;
}

int tbslas_b200_tree_update_coeff(tbslas_tree *t, const double *coeff, int mem) {
  if (!t || (t->n_leaf && !coeff)) return TBSLAS_ERR_INVALID;
  tbslas_ctx *ctx = t->ctx;
  if (!t->n_leaf) return TBSLAS_OK;
  TB_TRY(tree_coeff_ready(t));  // an earlier asynchronous upload lands first
  StageScope sc(ctx, ST_H2D, (double)(t->n_leaf * t->dof * t->ncoef * 8), 0);
  TB_TRY(copy_coeff_in(t, coeff, mem, ctx->stream));
  if (mem == TBSLAS_MEM_HOST) TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TBSLAS_OK;
}

int tbslas_b200_tree_update_coeff_async(tbslas_tree *t, const double *coeff, int mem) {
  if (!t || (t->n_leaf && !coeff)) return TBSLAS_ERR_INVALID;
  tbslas_ctx *ctx = t->ctx;
  if (!t->n_leaf) return TBSLAS_OK;
  if (!t->ev_coeff) TB_CUDA(ctx, cudaEventCreateWithFlags(&t->ev_coeff, cudaEventDisableTiming));
  // readers of the old coefficients already enqueued on the context's stream finish first
  TB_CUDA(ctx, cudaEventRecord(t->ev_coeff, ctx->stream));
  TB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_aux, t->ev_coeff, 0));
  ctx->acc_units[ST_H2D] += (double)(t->n_leaf * t->dof * t->ncoef * 8);
  TB_TRY(copy_coeff_in(t, coeff, mem, ctx->copy_aux));
  TB_CUDA(ctx, cudaEventRecord(t->ev_coeff, ctx->copy_aux));
  t->coeff_pending = true;
  return TBSLAS_OK;
}

int tbslas_b200_tree_get_coeff(tbslas_tree *t, double *coeff, int mem) {
  if (!t || (t->n_leaf && !coeff)) return TBSLAS_ERR_INVALID;
  tbslas_ctx *ctx = t->ctx;
  if (!t->n_leaf) return TBSLAS_OK;
  TB_TRY(tree_coeff_ready(t));
  const size_t ncoef_pad = t->stride / t->dof;
  StageScope sc(ctx, ST_D2H, (double)(t->n_leaf * t->dof * t->ncoef * 8), 0);
  const cudaMemcpyKind kind = mem == TBSLAS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (ncoef_pad == t->ncoef)
    TB_CUDA(ctx, cudaMemcpyAsync(coeff, t->d_coeff, sizeof(double) * t->ncoef * t->n_leaf * t->dof, kind,
                                 ctx->stream));
  else
    TB_CUDA(ctx, cudaMemcpy2DAsync(coeff, t->ncoef * sizeof(double), t->d_coeff, ncoef_pad * sizeof(double),
                                   t->ncoef * sizeof(double), t->n_leaf * t->dof, kind, ctx->stream));
  if (mem == TBSLAS_MEM_HOST) TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TBSLAS_OK;
}

// Co-partitioning (tbslas::SemiMergeTree / pvfmm RedistNodes, tree_utils.h:609-729): the leaves of a
// Morton-sharded tree move between ranks so that rank r owns global leaves [new_first[r],
// new_first[r+1]).  Collective.
int tbslas_b200_tree_reshard(tbslas_tree *t, const size_t *new_first) {
  if (!t || !new_first) return TBSLAS_ERR_INVALID;
  tbslas_ctx *ctx = t->ctx;
  if (t->replicated) return fail(ctx, TBSLAS_ERR_INVALID, "tree_reshard: a replicated tree has no shards");
  TB_TRY(tree_coeff_ready(t));
  if (ctx->nranks < 2) {
    if (new_first[0] != 0 || new_first[1] != t->n_leaf)
      return fail(ctx, TBSLAS_ERR_INVALID, "tree_reshard: single rank owns [0, %zu)", t->n_leaf);
    return TBSLAS_OK;
  }
  return comm_reshard(t, new_first);
}

int tbslas_b200_tree_global_range(const tbslas_tree *t, size_t *first, size_t *total) {
  if (!t) return TBSLAS_ERR_INVALID;
  if (first) *first = (size_t)t->leaf_offset;
  if (total) *total = t->n_leaf_global;
  return TBSLAS_OK;
}

int tbslas_b200_tree_destroy(tbslas_tree *t) {
  if (!t) return TBSLAS_ERR_INVALID;
  cudaStreamSynchronize(t->ctx->stream);
  if (t->coeff_pending) cudaStreamSynchronize(t->ctx->copy_aux);
  if (t->ev_coeff) cudaEventDestroy(t->ev_coeff);
  cudaFree(t->d_key);
  cudaFree(t->d_geom);
  cudaFree(t->d_depth);
  cudaFree(t->d_box);
  cudaFree(t->d_cell);
  cudaFree(t->d_pt_count);
  cudaFree(t->d_coeff);
  cudaFree(t->d_splitters);
  delete t;
  return TBSLAS_OK;
}

int tbslas_b200_tree_info(const tbslas_tree *t, int *q, int *dof, size_t *n_leaf) {
  if (!t) return TBSLAS_ERR_INVALID;
  if (q) *q = t->q;
  if (dof) *dof = t->dof;
  if (n_leaf) *n_leaf = t->n_leaf;
  return TBSLAS_OK;
}

// ---------------------------------------------------------------- evaluation
int tbslas_b200_eval_field(const tbslas_field *f, double tq, int bc, double *pos, size_t n,
                           double *out, int mem) {
  tbslas_ctx *ctx;
  int dof;
  TB_TRY(check_field(&ctx, f, &dof));
  if (n && (!pos || !out)) return fail(ctx, TBSLAS_ERR_INVALID, "null buffer");
  if (mem == TBSLAS_MEM_DEVICE) return eval_field_dev(f, tq, bc, pos, n, out, 0, nullptr, 0.0);
  PipeSpec sp;
  sp.h_pos = pos;
  sp.h_pos_b = (bc == TBSLAS_PERIODIC) ? pos : nullptr;  // wrapped in place
  sp.h_val = out;
  sp.val_dof = dof;
  return run_host_pipeline(
      ctx, sp, n, [](PipeBufs &, size_t, size_t) { return (int)TBSLAS_OK; },
      [&](PipeBufs &B, size_t m, size_t) { return eval_field_dev(f, tq, bc, B.pos, m, B.val, 0, nullptr, 0.0); });
}

int tbslas_b200_eval(tbslas_tree *t, int bc, double *pos, size_t n, double *out, int32_t *leaf_idx,
                     int mem) {
  if (!t) return TBSLAS_ERR_INVALID;
  tbslas_ctx *ctx = t->ctx;
  if (n && (!pos || !out)) return fail(ctx, TBSLAS_ERR_INVALID, "null buffer");
  if (mem == TBSLAS_MEM_DEVICE) return eval_tree_dev(t, bc, pos, n, EPI_STORE, out, nullptr, 0.0, leaf_idx);
  PipeSpec sp;
  sp.h_pos = pos;
  sp.h_pos_b = (bc == TBSLAS_PERIODIC) ? pos : nullptr;
  sp.h_val = out;
  sp.val_dof = t->dof;
  sp.h_leaf = leaf_idx;
  return run_host_pipeline(
      ctx, sp, n, [](PipeBufs &, size_t, size_t) { return (int)TBSLAS_OK; },
      [&](PipeBufs &B, size_t m, size_t) { return eval_tree_dev(t, bc, B.pos, m, EPI_STORE, B.val, nullptr, 0.0, B.leaf); });
}

int tbslas_b200_eval_set4(tbslas_tree *const trees[4], const double times[4], double tq, int bc,
                          double *pos, size_t n, double *out, int mem) {
  if (!trees || !times) return TBSLAS_ERR_INVALID;
  tbslas_field f;
  f.kind = TBSLAS_FIELD_SET4;
  for (int i = 0; i < 4; i++) {
    f.tree[i] = trees[i];
    f.times[i] = times[i];
  }
  return tbslas_b200_eval_field(&f, tq, bc, pos, n, out, mem);
}

int tbslas_b200_eval_extrap(tbslas_tree *tp, tbslas_tree *tc, int bc, double *pos, size_t n,
                            double *out, int mem) {
  tbslas_field f;
  memset(&f, 0, sizeof(f));
  f.kind = TBSLAS_FIELD_EXTRAP;
  f.tree[0] = tp;
  f.tree[1] = tc;
  return tbslas_b200_eval_field(&f, 0.0, bc, pos, n, out, mem);
}

int tbslas_b200_traj_rk2(const tbslas_field *f1, const tbslas_field *f2, int bc, const double *pos,
                         size_t n, double tinit, double tfinal, int nrk, double *out_pos, int mem) {
  tbslas_ctx *ctx, *ctx2;
  int dof, dof2;
  TB_TRY(check_field(&ctx, f1, &dof));
  if (f2) {
    TB_TRY(check_field(&ctx2, f2, &dof2));
    if (ctx2 != ctx || dof2 != dof) return fail(ctx, TBSLAS_ERR_INVALID, "f1/f2 mismatch");
  }
  if (dof != 3) return fail(ctx, TBSLAS_ERR_INVALID, "velocity field must have dof 3 (got %d)", dof);
  if (nrk < 1 || (n && (!pos || !out_pos))) return fail(ctx, TBSLAS_ERR_INVALID, "bad argument");
  if (mem == TBSLAS_MEM_HOST) {
    PipeSpec sp;
    sp.h_pos = pos;
    sp.h_pos_a = out_pos;
    sp.need_tmp = true;
    return run_host_pipeline(
        ctx, sp, n,
        [&](PipeBufs &B, size_t m, size_t) { return traj_rk2_dev(f1, f2, bc, B.pos, B.tmp, m, tinit, tfinal, nrk); },
        [](PipeBufs &, size_t, size_t) { return (int)TBSLAS_OK; });
  }
  void *xtmp;
  TB_TRY(ws_get(ctx, WS_POS_B, sizeof(double) * 3 * n, &xtmp));
  return traj_rk2_dev(f1, f2, bc, out_pos, (double *)xtmp, n, tinit, tfinal, nrk, pos);
}

// pos == nullptr: the arrival points are generated on the device from `con`'s own leaves
// (tbslas::CollectChebTreeGridPoints), i.e. steps (1)+(2) of tbslas::SolveSemilagInSitu.
static int semilag_impl(const tbslas_field *f1, const tbslas_field *f2, tbslas_tree *con, int bc,
                        const double *pos, size_t n, int timestep, double dt, int nrk,
                        double *out_vals, double *out_dep, int mem) {
  tbslas_ctx *ctx, *ctx2;
  int dof, dof2;
  TB_TRY(check_field(&ctx, f1, &dof));
  if (f2) {
    TB_TRY(check_field(&ctx2, f2, &dof2));
    if (ctx2 != ctx || dof2 != dof) return fail(ctx, TBSLAS_ERR_INVALID, "f1/f2 mismatch");
  }
  if (dof != 3) return fail(ctx, TBSLAS_ERR_INVALID, "velocity field must have dof 3 (got %d)", dof);
  if (!con || con->ctx != ctx) return fail(ctx, TBSLAS_ERR_INVALID, "advected tree missing");
  const bool insitu = (pos == nullptr);
  if (insitu) {
    const size_t d = con->q + 1;
    n = con->n_leaf * d * d * d;
  }
  if (nrk < 1 || (n && !out_vals)) return fail(ctx, TBSLAS_ERR_INVALID, "bad argument");
  const double tinit = timestep * dt;   // semilag.inc:34-35
  const double tfinal = tinit - dt;
  if (mem == TBSLAS_MEM_HOST && !insitu) {
    PipeSpec sp;
    sp.h_pos = pos;
    sp.h_pos_a = out_dep;  // as ComputeTrajRK2 returns them, before the scalar evaluation wraps them
    sp.h_val = out_vals;
    sp.val_dof = con->dof;
    sp.need_tmp = true;
    return run_host_pipeline(
        ctx, sp, n,
        [&](PipeBufs &B, size_t m, size_t) { return traj_rk2_dev(f1, f2, bc, B.pos, B.tmp, m, tinit, tfinal, nrk); },
        [&](PipeBufs &B, size_t m, size_t) { return eval_tree_dev(con, bc, B.pos, m, EPI_STORE, B.val, nullptr, 0.0, nullptr); });
  }
  if (mem == TBSLAS_MEM_HOST && insitu && (n || (ctx->nranks > 1 && !con->replicated))) {
    // arrival points generated per chunk of leaves in HBM; the values of chunk c-1 travel to the
    // host while chunk c is computed
    const size_t P = (size_t)(con->q + 1) * (con->q + 1) * (con->q + 1);
    PipeSpec sp;
    sp.h_pos_a = out_dep;
    sp.h_val = out_vals;
    sp.val_dof = con->dof;
    sp.need_tmp = true;
    sp.unit = P;
    sp.n_collective = con->n_leaf_max * P;
    return run_host_pipeline(
        ctx, sp, n,
        [&](PipeBufs &B, size_t m, size_t off) {
          return traj_rk2_dev(f1, f2, bc, B.pos, B.tmp, m, tinit, tfinal, nrk, B.pos, con, off / P, true, n);
        },
        [&](PipeBufs &B, size_t m, size_t) { return eval_tree_dev(con, bc, B.pos, m, EPI_STORE, B.val, nullptr, 0.0, nullptr); });
  }
  HostIO io{ctx, mem};
  void *xsol, *xtmp, *dval;
  if (mem == TBSLAS_MEM_DEVICE && out_dep)
    xsol = out_dep;
  else
    TB_TRY(ws_get(ctx, WS_POS_A, sizeof(double) * 3 * n, &xsol));
  TB_TRY(ws_get(ctx, WS_POS_B, sizeof(double) * 3 * n, &xtmp));
  TB_TRY(io.out_buf(WS_VAL_B, out_vals, sizeof(double) * con->dof * n, &dval));
  TB_TRY(traj_rk2_dev(f1, f2, bc, (double *)xsol, (double *)xtmp, n, tinit, tfinal, nrk,
                      insitu ? (const double *)xsol : pos, insitu ? con : nullptr, 0, insitu));
  if (mem == TBSLAS_MEM_DEVICE && out_dep && bc == TBSLAS_PERIODIC) {
    // keep the caller's departure points un-wrapped: evaluate on a copy
    TB_CUDA(ctx, cudaMemcpyAsync(xtmp, xsol, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice,
                                 ctx->stream));
    xsol = xtmp;
  }
  TB_TRY(eval_tree_dev(con, bc, (double *)xsol, n, EPI_STORE, (double *)dval, nullptr, 0.0, nullptr));
  TB_TRY(io.d2h(out_vals, dval, sizeof(double) * con->dof * n));
  return io.finish();
}

int tbslas_b200_semilag_rk2(const tbslas_field *f1, const tbslas_field *f2, tbslas_tree *con, int bc,
                            const double *pos, size_t n, int timestep, double dt, int nrk,
                            double *out_vals, double *out_dep, int mem) {
  if (n && !pos) {
    tbslas_ctx *ctx = (f1 && f1->tree[0]) ? f1->tree[0]->ctx : nullptr;
    return fail(ctx, TBSLAS_ERR_INVALID, "null point buffer");
  }
  static const double none[3] = {0, 0, 0};
  return semilag_impl(f1, f2, con, bc, pos ? pos : none, n, timestep, dt, nrk, out_vals, out_dep, mem);
}

int tbslas_b200_semilag_insitu(const tbslas_field *f1, const tbslas_field *f2, tbslas_tree *con, int bc,
                               int timestep, double dt, int nrk, double *out_vals, int mem) {
  return semilag_impl(f1, f2, con, bc, nullptr, 0, timestep, dt, nrk, out_vals, nullptr, mem);
}

int tbslas_b200_semilag_insitu_dep(const tbslas_field *f1, const tbslas_field *f2, tbslas_tree *con, int bc,
                                   int timestep, double dt, int nrk, double *out_vals, double *out_dep, int mem) {
  return semilag_impl(f1, f2, con, bc, nullptr, 0, timestep, dt, nrk, out_vals, out_dep, mem);
}

// ---------------------------------------------------------------- values -> coefficients
int tbslas_b200_set_pt2coeff(tbslas_ctx *ctx, int q, const double *M) {
  if (!ctx || !M) return TBSLAS_ERR_INVALID;
  if (q < 1 || q > TBSLAS_MAX_CHEB_DEG) return fail(ctx, TBSLAS_ERR_UNSUPPORTED, "degree %d not supported", q);
  return set_pt2coeff(ctx, q, M);
}

int tbslas_b200_has_pt2coeff(tbslas_ctx *ctx, int q, int *has) {
  if (!ctx || !has || q < 1 || q > TBSLAS_MAX_CHEB_DEG) return TBSLAS_ERR_INVALID;
  *has = ctx->pt2coeff[q].d_M != nullptr;
  return TBSLAS_OK;
}

int tbslas_b200_tree_set_grid_values(tbslas_tree *t, const double *vals, int point_major, int mem) {
  if (!t || (t->n_leaf && !vals)) return TBSLAS_ERR_INVALID;
  tbslas_ctx *ctx = t->ctx;
  const size_t d = t->q + 1, m = t->n_leaf * t->dof * d * d * d;
  TB_TRY(tree_coeff_ready(t));
  HostIO io{ctx, mem};
  void *dv;
  TB_TRY(io.h2d(WS_VAL_A, vals, sizeof(double) * m, &dv));
  TB_TRY(launch_refit(ctx, t, (const double *)dv, point_major));
  return io.finish();
}

int tbslas_b200_semilag_insitu_update(const tbslas_field *f1, const tbslas_field *f2, tbslas_tree *con,
                                      int bc, int timestep, double dt, int nrk) {
  if (!con) return TBSLAS_ERR_INVALID;
  tbslas_ctx *ctx = con->ctx;
  const size_t d = con->q + 1, n = con->n_leaf * d * d * d;
  void *dval;
  TB_TRY(ws_get(ctx, WS_VAL_C, sizeof(double) * con->dof * (n + 1), &dval));
  TB_TRY(semilag_impl(f1, f2, con, bc, nullptr, 0, timestep, dt, nrk, (double *)dval, nullptr, TBSLAS_MEM_DEVICE));
  // new coefficients in place: every evaluation of `con` above is stream-ordered before this
  return launch_refit(ctx, con, (const double *)dval, /*point_major=*/1);
}

// ---------------------------------------------------------------- cubic grid
struct tbslas_grid {
  tbslas_ctx *ctx = nullptr;
  int n_reg = 0, dof = 0;
  double *d_grid = nullptr;  // [dof][n_reg^3], resident
};

int tbslas_b200_grid_create(tbslas_ctx *ctx, const double *grid, int n_reg, int dof, int mem, tbslas_grid **out) {
  if (!ctx || !out) return TBSLAS_ERR_INVALID;
  *out = nullptr;
  if (n_reg < 4 || dof < 1 || !grid)
    return fail(ctx, TBSLAS_ERR_INVALID, "grid_create: bad argument (n_reg=%d, dof=%d)", n_reg, dof);
  tbslas_grid *g = new tbslas_grid();
  g->ctx = ctx;
  g->n_reg = n_reg;
  g->dof = dof;
  const size_t bytes = sizeof(double) * dof * (size_t)n_reg * n_reg * n_reg;
  if (cudaMalloc(&g->d_grid, bytes) != cudaSuccess) {
    delete g;
    return fail(ctx, TBSLAS_ERR_NOMEM, "grid_create: cudaMalloc(%zu bytes)", bytes);
  }
  *out = g;
  return tbslas_b200_grid_update(g, grid, mem);
}

int tbslas_b200_grid_update(tbslas_grid *g, const double *grid, int mem) {
  if (!g || !grid) return TBSLAS_ERR_INVALID;
  tbslas_ctx *ctx = g->ctx;
  const size_t bytes = sizeof(double) * g->dof * (size_t)g->n_reg * g->n_reg * g->n_reg;
  StageScope sc(ctx, ST_H2D, (double)bytes, 0);
  TB_CUDA(ctx, cudaMemcpyAsync(g->d_grid, grid, bytes,
                               mem == TBSLAS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                               ctx->stream));
  if (mem == TBSLAS_MEM_HOST) TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TBSLAS_OK;
}

int tbslas_b200_grid_destroy(tbslas_grid *g) {
  if (!g) return TBSLAS_ERR_INVALID;
  cudaStreamSynchronize(g->ctx->stream);
  cudaFree(g->d_grid);
  delete g;
  return TBSLAS_OK;
}

// fast_interp on a resident grid; host buffers: the queries stream in and the values out in chunks
// (copy-in, kernel and copy-out of consecutive chunks overlap), only 24 + 8*dof B per query cross PCIe
int tbslas_b200_grid_eval(tbslas_grid *g, const double *pos, size_t n, double *out, int mem) {
  if (!g) return TBSLAS_ERR_INVALID;
  tbslas_ctx *ctx = g->ctx;
  if (n && (!pos || !out)) return fail(ctx, TBSLAS_ERR_INVALID, "grid_eval: null buffer");
  if (mem == TBSLAS_MEM_DEVICE) return launch_cubic_grid(ctx, g->d_grid, g->n_reg, g->dof, pos, n, out);
  PipeSpec sp;
  sp.h_pos = pos;
  sp.h_val = out;
  sp.val_dof = g->dof;
  return run_host_pipeline(
      ctx, sp, n, [](PipeBufs &, size_t, size_t) { return (int)TBSLAS_OK; },
      [&](PipeBufs &B, size_t m, size_t) { return launch_cubic_grid(ctx, g->d_grid, g->n_reg, g->dof, B.pos, m, B.val); });
}

// one-shot form (the grid travels with the call): tbslas::fast_interp's own signature
int tbslas_b200_cubic_eval(tbslas_ctx *ctx, const double *grid, int n_reg, int dof, const double *pos,
                           size_t n, double *out, int mem) {
  if (!ctx) return TBSLAS_ERR_INVALID;
  if (n_reg < 4 || dof < 1 || !grid || (n && (!pos || !out)))
    return fail(ctx, TBSLAS_ERR_INVALID, "cubic_eval: bad argument (n_reg=%d, dof=%d)", n_reg, dof);
  if (mem == TBSLAS_MEM_DEVICE) return launch_cubic_grid(ctx, grid, n_reg, dof, pos, n, out);
  tbslas_grid *g = nullptr;
  TB_TRY(tbslas_b200_grid_create(ctx, grid, n_reg, dof, mem, &g));
  const int rc = tbslas_b200_grid_eval(g, pos, n, out, mem);
  tbslas_b200_grid_destroy(g);
  return rc;
}

// ---------------------------------------------------------------- next rows
int tbslas_b200_collect_grid_points(tbslas_tree *t, double *out_pos, int mem) {
  if (!t || !out_pos) return TBSLAS_ERR_INVALID;
  tbslas_ctx *ctx = t->ctx;
  const size_t d = t->q + 1, n = t->n_leaf * d * d * d;
  HostIO io{ctx, mem};
  void *dout;
  TB_TRY(io.out_buf(WS_POS_C, out_pos, sizeof(double) * 3 * n, &dout));
  TB_TRY(launch_grid_points(ctx, t, (double *)dout));
  TB_TRY(io.d2h(out_pos, dout, sizeof(double) * 3 * n));
  return io.finish();
}

int tbslas_b200_new_nodes(int q, double *out) {
  if (q < 0 || q > TBSLAS_MAX_CHEB_DEG || !out) return TBSLAS_ERR_INVALID;
  tb::new_nodes_host(q, out);
  return TBSLAS_OK;
}

// ---------------------------------------------------------------- host shard logic
uint64_t tbslas_b200_point_key(double x, double y, double z, int bc) {
  return point_key(x, y, z, bc == TBSLAS_PERIODIC);
}

int tbslas_b200_owner_of_key(uint64_t key, const uint64_t *splitters, int nranks) {
  int owner = 0;
  for (int r = 1; r < nranks; r++)
    if (splitters[r] <= key) owner = r;
  return owner;
}

int tbslas_b200_cubic_time_weights(const double times[4], double t, double w[4]) {
  if (!times || !w) return TBSLAS_ERR_INVALID;
  cubic_time_weights(times, t, w);
  return TBSLAS_OK;
}

int tbslas_b200_partition_leaves(size_t n_leaf, int nranks, size_t *first) {
  if (nranks < 1 || !first) return TBSLAS_ERR_INVALID;
  for (int r = 0; r <= nranks; r++) first[r] = (size_t)r * n_leaf / nranks;
  return TBSLAS_OK;
}

int tbslas_b200_partition_leaves_weighted(size_t n_leaf, const double *weight, int nranks, size_t *first) {
  if (nranks < 1 || !first || (n_leaf && !weight)) return TBSLAS_ERR_INVALID;
  double total = 0.0;
  for (size_t j = 0; j < n_leaf; j++) {
    if (!(weight[j] >= 0.0)) return TBSLAS_ERR_INVALID;  // negative or NaN
    total += weight[j];
  }
  if (!(total > 0.0)) return tbslas_b200_partition_leaves(n_leaf, nranks, first);
  // rank r starts at the first leaf whose weight prefix reaches r/nranks of the total
  size_t j = 0;
  double run = 0.0;
  first[0] = 0;
  for (int r = 1; r < nranks; r++) {
    const double target = total * (double)r / (double)nranks;
    while (j < n_leaf && run + 0.5 * weight[j] < target) run += weight[j++];
    first[r] = j;
  }
  first[nranks] = n_leaf;
  return TBSLAS_OK;
}

int tbslas_b200_tree_last_point_counts(tbslas_tree *t, uint32_t *counts, int mem) {
  if (!t || (t->n_leaf && !counts)) return TBSLAS_ERR_INVALID;
  tbslas_ctx *ctx = t->ctx;
  if (!t->n_leaf) return TBSLAS_OK;
  if (!t->pt_count_valid)
    return fail(ctx, TBSLAS_ERR_INVALID, "tree_last_point_counts: this tree has not been evaluated yet");
  TB_CUDA(ctx, cudaMemcpyAsync(counts, t->d_pt_count, sizeof(uint32_t) * t->n_leaf,
                               mem == TBSLAS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                               ctx->stream));
  if (mem == TBSLAS_MEM_HOST) TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TBSLAS_OK;
}

int tbslas_b200_tree_tail_norm(tbslas_tree *t, double *tail, int mem) {
  if (!t || (t->n_leaf && !tail)) return TBSLAS_ERR_INVALID;
  tbslas_ctx *ctx = t->ctx;
  if (!t->n_leaf) return TBSLAS_OK;
  TB_TRY(tree_coeff_ready(t));
  HostIO io{ctx, mem};
  void *d;
  TB_TRY(io.out_buf(WS_VAL_A, tail, sizeof(double) * t->n_leaf, &d));
  TB_TRY(launch_tail_norm(ctx, t, (double *)d));
  TB_TRY(io.d2h(tail, d, sizeof(double) * t->n_leaf));
  return io.finish();
}

// ---------------------------------------------------------------- instrumentation
int tbslas_b200_profile_enable(tbslas_ctx *ctx, int on) {
  if (!ctx) return TBSLAS_ERR_INVALID;
  TB_TRY(prof_collect(ctx));
  ctx->prof = on != 0;
  return TBSLAS_OK;
}
int tbslas_b200_profile_reset(tbslas_ctx *ctx) {
  if (!ctx) return TBSLAS_ERR_INVALID;
  TB_TRY(prof_collect(ctx));
  for (int s = 0; s < ST_COUNT; s++) {
    ctx->acc_ms[s] = 0;
    ctx->acc_units[s] = 0;
    ctx->acc_launch[s] = 0;
  }
  return TBSLAS_OK;
}
int tbslas_b200_profile_num_stages(void) { return ST_COUNT; }
const char *tbslas_b200_profile_stage_name(int s) {
  return (s >= 0 && s < ST_COUNT) ? kStageNames[s] : "";
}
const char *tbslas_b200_profile_reference_tag(int s) {
  return (s >= 0 && s < ST_COUNT) ? kStageRefTags[s] : "";
}
int tbslas_b200_profile_get(tbslas_ctx *ctx, int stage, double *ms, long long *launches, double *units) {
  if (!ctx || stage < 0 || stage >= ST_COUNT) return TBSLAS_ERR_INVALID;
  TB_TRY(prof_collect(ctx));
  if (ms) *ms = ctx->acc_ms[stage];
  if (launches) *launches = ctx->acc_launch[stage];
  if (units) *units = ctx->acc_units[stage];
  return TBSLAS_OK;
}
long long tbslas_b200_kernel_launches(tbslas_ctx *ctx) { return ctx ? ctx->launches : 0; }

int tbslas_b200_fp64_peak(tbslas_ctx *ctx, int reps, double *tflops) {
  if (!ctx || !tflops) return TBSLAS_ERR_INVALID;
  return run_fp64_peak(ctx, reps < 1 ? 1 : reps, tflops);
}

}  // extern "C"

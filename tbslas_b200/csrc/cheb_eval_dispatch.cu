// Degree dispatch for the Chebyshev evaluation: the persistent warp-pipelined kernel
// (cheb_eval_wt.cuh, one fully unrolled instantiation per degree 1..19, compiled in
// cheb_eval_inst_*.cu) whenever one coefficient buffer per warp fits in shared memory at two
// CTAs per SM; otherwise -- very wide dof -- the degree-generic kernel (cheb_eval_generic.cu).
#include "cheb_eval_wt.cuh"

namespace tb {

#define TB_DECL(Q) extern template int launch_cheb_eval_wt<Q, eval_ppt(Q)>(tbslas_ctx *, const EvalArgs &);
TB_DECL(1) TB_DECL(2) TB_DECL(3) TB_DECL(4) TB_DECL(5) TB_DECL(6) TB_DECL(7) TB_DECL(8)
TB_DECL(9) TB_DECL(10) TB_DECL(11) TB_DECL(12) TB_DECL(13) TB_DECL(14)
TB_DECL(15) TB_DECL(16) TB_DECL(17) TB_DECL(18) TB_DECL(19)
#undef TB_DECL
static_assert(TBSLAS_MAX_CHEB_DEG == 19, "one unrolled instantiation per supported degree");
extern template int launch_cheb_eval_q<8, eval_ppt(8)>(tbslas_ctx *, const EvalArgs &);
extern template int launch_cheb_eval_q<14, eval_ppt(14)>(tbslas_ctx *, const EvalArgs &);

constexpr int kMaxUnrolledDeg = 19;  // every supported degree has its unrolled instantiation
// One coefficient block + staging per one-warp CTA.  Up to 27 KB eight CTAs share an SM (two warps per
// sub-partition, what the kernel is tuned for); larger blocks (dof > 3 at q = 14, q >= 17 with three
// components) run with fewer resident CTAs -- the launcher sizes the grid by the occupancy -- which still beats
// the degree-generic kernel several times over (profiles/r02_eval_degree_sweep.json).  Beyond two CTAs per
// SM's worth of shared memory the generic kernel takes over.
constexpr size_t kWtSmemLimit = 100 * 1024;
int launch_cheb_eval_generic(tbslas_ctx *ctx, const EvalArgs &a);

// 0: warp-pipelined kernel, 1: one-tile-per-CTA kernel, 2: generic kernel
// (variant: the context's A/B switch, TBSLAS_EVAL_VARIANT read once at tbslas_b200_init)
static int eval_kind(const tbslas_tree *t) {
  const int q = t->q, variant = t->ctx->opt.eval_variant;
  const size_t stride = t->stride;
  if (q < 1 || q > TBSLAS_MAX_CHEB_DEG) return -1;
  if (q > kMaxUnrolledDeg) return 2;
  if (variant == 1 && (q == 8 || q == 14)) return 1;
  if (variant == 2) return 2;
  // (13 * 32 * PPT: the largest staging area, that of the grid-base epilogue)
  const size_t smem = (stride + 13 * 32 * (size_t)eval_ppt(q) + 16) * sizeof(double);
  return smem <= kWtSmemLimit ? 0 : 2;
}

int eval_tile_points(const tbslas_tree *t) {
  switch (eval_kind(t)) {
    case 0: return 32 * eval_ppt(t->q);
    case 1: return kEvalThreads * eval_ppt(t->q);
    case 2: return kEvalThreads;
    default: return 0;
  }
}

bool eval_needs_tile_map(const tbslas_tree *t) { return eval_kind(t) != 0; }
bool eval_supports_grid_base(const tbslas_tree *t) { return eval_kind(t) == 0 && t->dof == 3; }

int launch_cheb_eval(tbslas_ctx *ctx, const EvalArgs &a) {
  StageScope sc(ctx, ST_CHEB_EVAL, (double)a.n, 1);
  const int kind = eval_kind(a.tree);
  if (a.epilogue == EPI_AXPY_GRID && kind != 0)
    return fail(ctx, TBSLAS_ERR_UNSUPPORTED, "grid-base epilogue needs the persistent evaluation kernel");
  if (kind == 2) return launch_cheb_eval_generic(ctx, a);
  if (kind == 1)
    return a.tree->q == 8 ? launch_cheb_eval_q<8, eval_ppt(8)>(ctx, a) : launch_cheb_eval_q<14, eval_ppt(14)>(ctx, a);
  switch (a.tree->q) {
#define TB_CASE(Q) \
  case Q:          \
    return launch_cheb_eval_wt<Q, eval_ppt(Q)>(ctx, a);
    TB_CASE(1) TB_CASE(2) TB_CASE(3) TB_CASE(4) TB_CASE(5) TB_CASE(6) TB_CASE(7) TB_CASE(8)
    TB_CASE(9) TB_CASE(10) TB_CASE(11) TB_CASE(12) TB_CASE(13) TB_CASE(14)
    TB_CASE(15) TB_CASE(16) TB_CASE(17) TB_CASE(18) TB_CASE(19)
#undef TB_CASE
    default:
      return fail(ctx, TBSLAS_ERR_UNSUPPORTED, "Chebyshev degree %d not supported", a.tree->q);
  }
}

}  // namespace tb

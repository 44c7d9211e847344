// Degree dispatch for the Chebyshev evaluation kernel: one fully unrolled
// instantiation per degree 1..14 (compiled in cheb_eval_inst_*.cu), the degree-generic
// kernel (cheb_eval_generic.cu) for 15..TBSLAS_MAX_CHEB_DEG.
#include <cstdlib>

#include "cheb_eval.cuh"

namespace tb {

#define TB_DECL(Q) extern template int launch_cheb_eval_q<Q, eval_ppt(Q)>(tbslas_ctx *, const EvalArgs &);
TB_DECL(1) TB_DECL(2) TB_DECL(3) TB_DECL(4) TB_DECL(5) TB_DECL(6) TB_DECL(7) TB_DECL(8)
TB_DECL(9) TB_DECL(10) TB_DECL(11) TB_DECL(12) TB_DECL(13) TB_DECL(14)
#undef TB_DECL
// experimental variants of the high-degree kernel (selected by TBSLAS_EVAL_VARIANT)
extern template int launch_cheb_eval_q<14, 3, false>(tbslas_ctx *, const EvalArgs &);
extern template int launch_cheb_eval_q<14, 4, true>(tbslas_ctx *, const EvalArgs &);
extern template int launch_cheb_eval_q<14, 3, true>(tbslas_ctx *, const EvalArgs &);
extern template int launch_cheb_eval_q<14, 2, false, 1, false>(tbslas_ctx *, const EvalArgs &);
extern template int launch_cheb_eval_q<14, 2, false, 1, true>(tbslas_ctx *, const EvalArgs &);
extern template int launch_cheb_eval_q<14, 2, false, 4, false>(tbslas_ctx *, const EvalArgs &);

constexpr int kMaxUnrolledDeg = 14;  // beyond this nvcc stops unrolling: generic kernel
int launch_cheb_eval_generic(tbslas_ctx *ctx, const EvalArgs &a);

static int eval_variant() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("TBSLAS_EVAL_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

int eval_tile_points(int q) {
  if (q < 1 || q > TBSLAS_MAX_CHEB_DEG) return 0;
  if (q == 14 && eval_variant() == 1) return kEvalThreads * 3 * kEvalBatches;
  if (q == 14 && eval_variant() == 2) return kEvalThreads * 4 * kEvalBatches;
  if (q == 14 && eval_variant() == 3) return kEvalThreads * 3 * kEvalBatches;
  if (q == 14 && (eval_variant() == 4 || eval_variant() == 5)) return kEvalThreads * 2;
  return q <= kMaxUnrolledDeg ? kEvalThreads * eval_ppt(q) * kEvalBatches : kEvalThreads;
}

int launch_cheb_eval(tbslas_ctx *ctx, const EvalArgs &a) {
  StageScope sc(ctx, ST_CHEB_EVAL, (double)a.n, 1);
  if (a.tree->q == 14 && eval_variant() == 1) return launch_cheb_eval_q<14, 3, false>(ctx, a);
  if (a.tree->q == 14 && eval_variant() == 2) return launch_cheb_eval_q<14, 4, true>(ctx, a);
  if (a.tree->q == 14 && eval_variant() == 3) return launch_cheb_eval_q<14, 3, true>(ctx, a);
  if (a.tree->q == 14 && eval_variant() == 4) return launch_cheb_eval_q<14, 2, false, 1, false>(ctx, a);
  if (a.tree->q == 14 && eval_variant() == 5) return launch_cheb_eval_q<14, 2, false, 1, true>(ctx, a);
  if (a.tree->q == 14 && eval_variant() == 6) return launch_cheb_eval_q<14, 2, false, 4, false>(ctx, a);
  switch (a.tree->q) {
#define TB_CASE(Q) \
  case Q:          \
    return launch_cheb_eval_q<Q, eval_ppt(Q)>(ctx, a);
    TB_CASE(1) TB_CASE(2) TB_CASE(3) TB_CASE(4) TB_CASE(5) TB_CASE(6) TB_CASE(7) TB_CASE(8)
    TB_CASE(9) TB_CASE(10) TB_CASE(11) TB_CASE(12) TB_CASE(13) TB_CASE(14)
#undef TB_CASE
    default:
      if (a.tree->q > kMaxUnrolledDeg && a.tree->q <= TBSLAS_MAX_CHEB_DEG)
        return launch_cheb_eval_generic(ctx, a);
      return fail(ctx, TBSLAS_ERR_UNSUPPORTED, "Chebyshev degree %d not supported", a.tree->q);
  }
}

}  // namespace tb

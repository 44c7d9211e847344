// The one-tile-per-CTA kernel of cheb_eval.cuh, kept for A/B measurements against the
// persistent warp-pipelined kernel (TBSLAS_EVAL_VARIANT=1; degrees 8 and 14 only).
#include "cheb_eval.cuh"
namespace tb {
template int launch_cheb_eval_q<8, eval_ppt(8)>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_q<14, eval_ppt(14)>(tbslas_ctx *, const EvalArgs &);
}  // namespace tb

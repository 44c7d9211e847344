// Experimental variants of the degree-14 kernel (see cheb_eval_dispatch.cu).
#include "cheb_eval.cuh"
namespace tb {
template int launch_cheb_eval_q<14, 3, false>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_q<14, 4, true>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_q<14, 3, true>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_q<14, 2, false, 1, false>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_q<14, 2, false, 1, true>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_q<14, 2, false, 4, false>(tbslas_ctx *, const EvalArgs &);
}  // namespace tb

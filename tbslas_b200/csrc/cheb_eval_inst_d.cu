// Explicit instantiations of the evaluation kernel (split across files to compile in parallel).
#include "cheb_eval_wt.cuh"
namespace tb {
template int launch_cheb_eval_wt<13, eval_ppt(13)>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_wt<14, eval_ppt(14)>(tbslas_ctx *, const EvalArgs &);
}  // namespace tb

// Explicit instantiations of the evaluation kernel (split across files to compile in parallel).
#include "cheb_eval_wt.cuh"
namespace tb {
template int launch_cheb_eval_wt<11, eval_ppt(11)>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_wt<12, eval_ppt(12)>(tbslas_ctx *, const EvalArgs &);
}  // namespace tb

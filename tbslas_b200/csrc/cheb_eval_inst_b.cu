// Explicit instantiations of the evaluation kernel (split across files to compile in parallel).
#include "cheb_eval_wt.cuh"
namespace tb {
template int launch_cheb_eval_wt<8, eval_ppt(8)>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_wt<9, eval_ppt(9)>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_wt<10, eval_ppt(10)>(tbslas_ctx *, const EvalArgs &);
}  // namespace tb

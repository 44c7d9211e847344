// Explicit instantiations of the evaluation kernel (split across files to compile in parallel).
#include "cheb_eval_wt.cuh"
namespace tb {
template int launch_cheb_eval_wt<1, eval_ppt(1)>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_wt<2, eval_ppt(2)>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_wt<3, eval_ppt(3)>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_wt<4, eval_ppt(4)>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_wt<5, eval_ppt(5)>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_wt<6, eval_ppt(6)>(tbslas_ctx *, const EvalArgs &);
template int launch_cheb_eval_wt<7, eval_ppt(7)>(tbslas_ctx *, const EvalArgs &);
}  // namespace tb

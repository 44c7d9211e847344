// Explicit instantiations of the evaluation kernel (split across files to compile in parallel).
#include "cheb_eval_wt.cuh"
namespace tb {
template int launch_cheb_eval_wt<18, eval_ppt(18)>(tbslas_ctx *, const EvalArgs &);
}  // namespace tb

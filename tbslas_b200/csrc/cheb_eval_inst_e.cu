// Explicit instantiations of the evaluation kernel (split across files to compile in parallel).
#include "cheb_eval_wt.cuh"
namespace tb {
template int launch_cheb_eval_wt<15, eval_ppt(15)>(tbslas_ctx *, const EvalArgs &);
}  // namespace tb

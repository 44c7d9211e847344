// Explicit instantiations of the evaluation kernel (split across files to compile in parallel).
#include "cheb_eval_wt.cuh"
namespace tb {
template int launch_cheb_eval_wt<16, eval_ppt(16)>(tbslas_ctx *, const EvalArgs &);
}  // namespace tb

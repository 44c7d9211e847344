// Multi-GPU point redistribution (placeholder until the exchange path lands).
#include "common.cuh"
namespace tb {
int comm_tree_splitters(tbslas_tree *t) { return fail(t->ctx, TBSLAS_ERR_UNSUPPORTED, "multi-rank not built"); }
int comm_eval_outsiders(tbslas_tree *t, int, const double *, size_t, const int32_t *, const uint32_t *,
                        const uint32_t *, int, double *, const double *, double, int32_t *) {
  return fail(t->ctx, TBSLAS_ERR_UNSUPPORTED, "multi-rank not built");
}
void comm_destroy(tbslas_ctx *) {}
}  // namespace tb
extern "C" {
int tbslas_b200_comm_unique_id(void *) { return TBSLAS_ERR_UNSUPPORTED; }
int tbslas_b200_comm_init(tbslas_ctx *, int, int, const void *) { return TBSLAS_ERR_UNSUPPORTED; }
int tbslas_b200_comm_rank(tbslas_ctx *ctx, int *rank, int *nranks) {
  if (!ctx) return TBSLAS_ERR_INVALID;
  if (rank) *rank = ctx->rank;
  if (nranks) *nranks = ctx->nranks;
  return TBSLAS_OK;
}
}

// K9: ranking metrics over top-K lists.  Stands in for the pure-Python set probing of topKMetrics
// (/root/reference/trainers/topKmetrics.py:74-99): tp = #(u, rank) whose (u, item) is a positive,
// hits = #users with at least one tp; the remaining fields (fp, fn, tn, precision, recall, hitRate)
// follow from these counts on the host.  NDCG@k (not in the reference; binary relevance, IDCG over
// min(|pos_u|, k)) is accumulated alongside.  Positives are a CSR over users with sorted item lists;
// one thread per user, k binary searches each: 4k B of ids read per user + L2-resident CSR probes.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
topk_metrics_kernel(const int32_t* __restrict__ ids, int64_t U, int k, const int32_t* __restrict__ user_ids,
                    const int64_t* __restrict__ indptr, const int32_t* __restrict__ items, int64_t csr_users,
                    unsigned long long* __restrict__ counts, double* __restrict__ ndcg_sum) {
  __shared__ double red[32];
  unsigned int tp = 0, hits = 0;
  double ndcg = 0.0;
  for (int64_t r = int64_t(blockIdx.x) * kThreads + threadIdx.x; r < U; r += int64_t(gridDim.x) * kThreads) {
    const int64_t u = user_ids ? int64_t(__ldg(user_ids + r)) : r;
    int64_t lo0 = 0, hi0 = 0;
    if (u >= 0 && u < csr_users) { lo0 = __ldg(indptr + u); hi0 = __ldg(indptr + u + 1); }
    bool hit = false;
    double dcg = 0.0;
    for (int j = 0; j < k; ++j) {
      const int32_t it = __ldg(ids + r * k + j);
      int64_t lo = lo0, hi = hi0;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(items + mid) < it) lo = mid + 1; else hi = mid;
      }
      if (it >= 0 && lo < hi0 && __ldg(items + lo) == it) {
        ++tp; hit = true;
        dcg += 1.0 / log2(double(j + 2));
      }
    }
    hits += hit;
    const int64_t m = (hi0 - lo0 < int64_t(k)) ? (hi0 - lo0) : int64_t(k);
    if (m > 0) {
      double idcg = 0.0;
      for (int j = 0; j < m; ++j) idcg += 1.0 / log2(double(j + 2));
      ndcg += dcg / idcg;
    }
  }
  // block reduction, then one atomic per block and counter
  const double s_nd = block_sum_double(ndcg, red);
  __syncthreads();
  const double s_tp = block_sum_double(double(tp), red);
  __syncthreads();
  const double s_hit = block_sum_double(double(hits), red);
  if (threadIdx.x == 0) {
    atomicAdd(counts + 0, (unsigned long long)(s_tp + 0.5));
    atomicAdd(counts + 1, (unsigned long long)(s_hit + 0.5));
    atomicAdd(ndcg_sum, s_nd);
  }
}

}  // namespace

extern "C" int brk_topk_metrics(brk_ctx* ctx, const int32_t* ids, int64_t U, int32_t k, const int32_t* user_ids,
                                const int64_t* pos_indptr, const int32_t* pos_items, int64_t csr_users,
                                int64_t* counts_out, double* ndcg_sum_out, void* stream) {
  BRK_REQUIRE(ctx && ids && pos_indptr && pos_items && counts_out && ndcg_sum_out, BRK_E_ARG,
              "brk_topk_metrics: null argument");
  BRK_REQUIRE(U > 0 && k >= 1 && csr_users > 0, BRK_E_ARG, "brk_topk_metrics: U=%lld k=%d", (long long)U, k);
  cudaStream_t st = (cudaStream_t)stream;
  BRK_CUDA(cudaMemsetAsync(counts_out, 0, 2 * sizeof(int64_t), st));
  BRK_CUDA(cudaMemsetAsync(ndcg_sum_out, 0, sizeof(double), st));
  int64_t need = (U + kThreads - 1) / kThreads;
  const int64_t cap = int64_t(ctx->sm_count) * 8;
  topk_metrics_kernel<<<int(need < cap ? need : cap), kThreads, 0, st>>>(
      ids, U, k, user_ids, pos_indptr, pos_items, csr_users, reinterpret_cast<unsigned long long*>(counts_out),
      ndcg_sum_out);
  BRK_LAUNCH_CHECK();
  return 0;
}

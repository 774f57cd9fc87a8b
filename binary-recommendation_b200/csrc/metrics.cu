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


// ---- per-user AUC and average precision at k (src/models/bpr.py:230-289: full_auc, mean_average_precision_k) -----------
// The reference scores one user against the whole catalog (bpr_predict), hands the scores to sklearn's roc_auc_score
// and, for MAP, sorts a dict of all items per user in Python.  Both numbers only need, for every POSITIVE p of the
// user, how many items beat it:  gt = #{j: s_j > s_p},  eq_before = #{j < p: s_j == s_p}  (Python's stable
// sorted(..., reverse=True) keeps the earlier item first among equals) and the same counts restricted to positives:
//   rank_p = gt + eq_before (0-based position in the sorted catalog),
//   AUC    = sum_p (#negatives below p + 0.5 #negatives equal to p) / (n_pos n_neg)        (ties count half, as sklearn),
//   AP@k   = sum_{p: rank_p < k} (1 + #{p': rank_p' < rank_p}) / (rank_p + 1)  /  min(len(actual), k).
// One CTA per user row of a score matrix [R, I]; a warp per positive, lanes striding over the catalog: n_pos I
// compares per user, scores L2-resident.  Positions of the positives: CSR (indptr int64 [R+1], cols int32 sorted).
__global__ void __launch_bounds__(kThreads)
rank_eval_kernel(const float* __restrict__ scores, int64_t I, const int64_t* __restrict__ indptr,
                 const int32_t* __restrict__ cols, const int32_t* __restrict__ actual_len, int k,
                 int32_t* __restrict__ rank_ws, double* __restrict__ part_ws, double* __restrict__ out) {
  __shared__ double red[32];
  const int64_t r = blockIdx.x;
  const float* __restrict__ s = scores + r * I;
  const int64_t lo = indptr[r], hi = indptr[r + 1];
  const int np = int(hi - lo);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kThreads / 32;
  for (int a = warp; a < np; a += nw) {
    const int p = cols[lo + a];
    const float sp = s[p];
    int gt = 0, eqb = 0, eq = 0;
    for (int64_t j = lane; j < I; j += 32) {
      const float v = s[j];
      gt += v > sp;
      eq += (v == sp) && j != p;
      eqb += (v == sp) && j < p;
    }
    int gtp = 0, eqp = 0;
    for (int b = lane; b < np; b += 32) {
      const int q = cols[lo + b];
      const float v = s[q];
      gtp += v > sp;
      eqp += (v == sp) && q != p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gt += __shfl_xor_sync(0xffffffffu, gt, o); eq += __shfl_xor_sync(0xffffffffu, eq, o);
      eqb += __shfl_xor_sync(0xffffffffu, eqb, o); gtp += __shfl_xor_sync(0xffffffffu, gtp, o);
      eqp += __shfl_xor_sync(0xffffffffu, eqp, o);
    }
    if (lane == 0) {
      const int64_t n_neg = I - np;
      const int64_t gt_neg = gt - gtp, eq_neg = eq - eqp;
      rank_ws[lo + a] = gt + eqb;
      part_ws[lo + a] = double(n_neg - gt_neg - eq_neg) + 0.5 * double(eq_neg);
    }
  }
  __syncthreads();
  double auc = 0.0, ap = 0.0;
  for (int a = threadIdx.x; a < np; a += kThreads) {
    auc += part_ws[lo + a];
    const int ra = rank_ws[lo + a];
    if (ra < k) {
      int before = 0;
      for (int b = 0; b < np; ++b) before += rank_ws[lo + b] < ra;
      ap += double(1 + before) / double(ra + 1);
    }
  }
  auc = block_sum_double(auc, red);
  __syncthreads();
  ap = block_sum_double(ap, red);
  if (threadIdx.x == 0) {
    const double n_neg = double(I - np);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    out[2 * r] = (np > 0 && n_neg > 0) ? auc / (double(np) * n_neg) : nan;      // sklearn raises on a single class
    const int al = actual_len ? actual_len[r] : np;
    const int den = al < k ? al : k;
    out[2 * r + 1] = den > 0 ? ap / double(den) : nan;                            // the reference divides by zero here
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

extern "C" int brk_rank_eval_rows(brk_ctx* ctx, const float* scores, int64_t R, int64_t I, const int64_t* pos_indptr,
                                  const int32_t* pos_cols, const int32_t* actual_len, int32_t k, int32_t* rank_ws,
                                  double* part_ws, double* out, void* stream) {
  BRK_REQUIRE(ctx && scores && pos_indptr && pos_cols && rank_ws && part_ws && out, BRK_E_ARG,
              "brk_rank_eval_rows: null argument");
  BRK_REQUIRE(R > 0 && I > 0 && I < (int64_t(1) << 31) && k >= 1, BRK_E_ARG, "brk_rank_eval_rows: R=%lld I=%lld k=%d",
              (long long)R, (long long)I, k);
  BRK_REQUIRE(R < (int64_t(1) << 31), BRK_E_ARG, "brk_rank_eval_rows: R=%lld rows per call", (long long)R);
  rank_eval_kernel<<<int(R), kThreads, 0, (cudaStream_t)stream>>>(scores, I, pos_indptr, pos_cols, actual_len, k, rank_ws,
                                                                  part_ws, out);
  BRK_LAUNCH_CHECK();
  return 0;
}

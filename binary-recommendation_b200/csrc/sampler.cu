// K10: counter-based negative sampling, Philox4x32-10.  The stream ("brk sampler v2") is defined
// in oracle/philox.py; the reference samples with the host's global RNG
// (/root/reference/src/models/NeuMFModel.py:104-105) or enumerates exhaustively
// (/root/reference/src/models/BPRModel.py:111-119).  One thread per sample; ~4 B id read +
// a binary search in the user's (L2-resident) sorted positive list + 4 B written.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr uint32_t kTagNeumf = 0x4Eu;

__global__ void __launch_bounds__(kThreads)
philox_bpr_negatives(const int32_t* __restrict__ users, int64_t n, int64_t first_index, uint32_t seed,
                     uint32_t epoch, uint32_t num_items, const int64_t* __restrict__ indptr,
                     const int32_t* __restrict__ items, int32_t* __restrict__ neg) {
  const int64_t stride = int64_t(gridDim.x) * kThreads;
  for (int64_t b = int64_t(blockIdx.x) * kThreads + threadIdx.x; b < n; b += stride) {
    neg[b] = brk_sample_bpr_negative(uint64_t(first_index + b), __ldg(users + b), seed, epoch, num_items, indptr, items);
  }
}

__global__ void __launch_bounds__(kThreads)
philox_neumf_negatives(const int32_t* __restrict__ pos_users, const int32_t* __restrict__ pos_items,
                       uint32_t num_pos, int64_t n, int64_t first_index, uint32_t seed, uint32_t epoch,
                       int32_t* __restrict__ neg_users, int32_t* __restrict__ neg_items) {
  const int64_t stride = int64_t(gridDim.x) * kThreads;
  for (int64_t b = int64_t(blockIdx.x) * kThreads + threadIdx.x; b < n; b += stride) {
    const uint64_t idx = uint64_t(first_index + b);
    const uint4 w = philox4x32_10(make_uint4(uint32_t(idx), uint32_t(idx >> 32), 0u, kTagNeumf), seed, epoch);
    neg_users[b] = __ldg(pos_users + __umulhi(w.x, num_pos));
    neg_items[b] = __ldg(pos_items + __umulhi(w.y, num_pos));
  }
}

__global__ void __launch_bounds__(kThreads)
philox_raw(const uint4* __restrict__ ctr, int64_t n, uint32_t k0, uint32_t k1, uint4* __restrict__ out) {
  const int64_t stride = int64_t(gridDim.x) * kThreads;
  for (int64_t i = int64_t(blockIdx.x) * kThreads + threadIdx.x; i < n; i += stride)
    out[i] = philox4x32_10(ctr[i], k0, k1);
}

int grid_1d(const brk_ctx* ctx, int64_t n) {
  int64_t need = (n + kThreads - 1) / kThreads;
  const int64_t cap = int64_t(ctx->sm_count) * (2048 / kThreads);
  return int(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace

extern "C" int brk_philox_bpr_negatives(brk_ctx* ctx, const int32_t* users, int64_t n, int64_t first_index,
                                        uint32_t seed, uint32_t epoch, int32_t num_items,
                                        const int64_t* csr_indptr, const int32_t* csr_items,
                                        int32_t* neg, void* stream) {
  BRK_REQUIRE(ctx && (n == 0 || (users && neg && csr_indptr && csr_items)), BRK_E_ARG,
              "brk_philox_bpr_negatives: null argument");
  BRK_REQUIRE(n >= 0 && num_items > 0 && first_index >= 0, BRK_E_ARG,
              "brk_philox_bpr_negatives: n=%lld num_items=%d first_index=%lld", (long long)n, num_items,
              (long long)first_index);
  if (n == 0) return 0;
  philox_bpr_negatives<<<grid_1d(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(
      users, n, first_index, seed, epoch, uint32_t(num_items), csr_indptr, csr_items, neg);
  BRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int brk_philox_neumf_negatives(brk_ctx* ctx, const int32_t* pos_users, const int32_t* pos_items,
                                          int64_t num_pos, int64_t n, int64_t first_index, uint32_t seed,
                                          uint32_t epoch, int32_t* neg_users, int32_t* neg_items,
                                          void* stream) {
  BRK_REQUIRE(ctx && (n == 0 || (pos_users && pos_items && neg_users && neg_items)), BRK_E_ARG,
              "brk_philox_neumf_negatives: null argument");
  BRK_REQUIRE(n >= 0 && num_pos > 0 && num_pos < (int64_t(1) << 32) && first_index >= 0, BRK_E_ARG,
              "brk_philox_neumf_negatives: n=%lld num_pos=%lld", (long long)n, (long long)num_pos);
  if (n == 0) return 0;
  philox_neumf_negatives<<<grid_1d(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(
      pos_users, pos_items, uint32_t(num_pos), n, first_index, seed, epoch, neg_users, neg_items);
  BRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int brk_philox4x32_10(brk_ctx* ctx, const uint32_t* ctr, int64_t n, uint32_t key0, uint32_t key1,
                                 uint32_t* out, void* stream) {
  BRK_REQUIRE(ctx && (n == 0 || (ctr && out)), BRK_E_ARG, "brk_philox4x32_10: null argument");
  BRK_REQUIRE(n >= 0 && brk_aligned16(ctr) && brk_aligned16(out), BRK_E_ALIGN,
              "brk_philox4x32_10: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  philox_raw<<<grid_1d(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(ctr), n, key0, key1, reinterpret_cast<uint4*>(out));
  BRK_LAUNCH_CHECK();
  return 0;
}

// Device input pipeline: what sits on the caller's side of the
// training step in the reference --
//   * bootstrapDataset (/root/reference/src/models/NeuMFModel.py:102-123): sample negatives, label, row-shuffle
//     the merged frame, batch, shuffle the batches -- done there with pandas on the host, O(N) per call;
//   * pd.unique / StringLookup vocabularies (/root/reference/trainers/loadBinaryMovieLens.py:16-19,58-61,
//     /root/reference/trainers/twoTower.py:33-36).
// Here: a keyed counter-based permutation ("brk perm v1", oracle/pipeline.py: Feistel network with Philox-derived
// round keys + cycle walking -- no sort, no memory), ONE kernel that writes the shuffled
// (user, item, label) frame of an epoch (permutation, positive copy / Philox negative with optional collision
// rejection against the per-user sorted positive lists, label), and a hash-table factorisation that reproduces
// pd.unique's first-occurrence order bit for bit.  All of it is integer work bound by HBM / L2 atomics:
//   epoch build  : 12 B written per interaction + one 8-byte random read of the source pair (sector-granular);
//   factorisation: 8 B key read + 4 B id written per element, one 64-bit CAS + one 32-bit MIN per element.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr uint32_t kTagPerm = 0x5Eu;
constexpr uint32_t kTagNeumf = 0x4Eu;
constexpr int kPermRounds = 6;
constexpr int kMaxAttempts = 8;
constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int32_t kNoPos = 0x7F7F7F7F;     // what a 0x7F byte fill leaves: larger than every position

// "brk perm v1" (oracle/pipeline.py): six-round Feistel network over bits = max(2, ceil(log2 n)) bits (left part
// bits / 2, right part the rest; the widths swap every round), cycle walking.  Philox4x32-10 supplies the six round
// key pairs (computed once on the host, passed by value); the per-element round function is a two-multiply 32-bit
// mixer, so one evaluation costs 12 multiplies and the kernels below run at memory speed.
struct PermKeys {
  uint32_t k0[kPermRounds], k1[kPermRounds];
  uint64_t n;
  int32_t hl, hr;
};

inline uint32_t mulhi32(uint32_t a, uint32_t b) { return uint32_t((uint64_t(a) * b) >> 32); }
inline void philox_host(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

inline PermKeys make_perm_keys(uint64_t n, uint32_t seed, uint32_t epoch, uint32_t salt) {
  PermKeys k;
  for (int r = 0; r < kPermRounds; ++r) {
    uint32_t c[4] = {uint32_t(r), 0u, salt, kTagPerm};
    philox_host(c, seed, epoch);
    k.k0[r] = c[0];
    k.k1[r] = c[1];
  }
  int bits = 2;
  while (bits < 63 && (uint64_t(1) << bits) < n) ++bits;
  k.n = n;
  k.hl = bits / 2;
  k.hr = bits - k.hl;
  return k;
}

BRK_HD inline uint32_t perm_mix32(uint32_t x, uint32_t k0, uint32_t k1) {
  x ^= k0;
  x ^= x >> 16; x *= 0x7FEB352Du;
  x ^= x >> 15; x += k1; x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}

BRK_HD inline uint64_t feistel_perm(uint64_t j, const PermKeys& k) {
  const uint64_t ml = (uint64_t(1) << k.hl) - 1, mr = (uint64_t(1) << k.hr) - 1;
  uint64_t x = j;
  do {
    uint64_t L = x >> k.hr, R = x & mr;
    for (int r = 0; r < kPermRounds; r += 2) {          // two rounds per trip: the part widths swap and swap back
      const uint64_t t = L ^ (uint64_t(perm_mix32(uint32_t(R), k.k0[r], k.k1[r])) & ml);            // round r:     (L, R) <- (R, t)
      const uint64_t t2 = R ^ (uint64_t(perm_mix32(uint32_t(t), k.k0[r + 1], k.k1[r + 1])) & mr);   // round r + 1: (L, R) <- (t, t2)
      L = t;
      R = t2;
    }
    x = (L << k.hr) | R;
  } while (x >= k.n);
  return x;
}

__global__ void __launch_bounds__(kThreads)
epoch_permutation_kernel(const PermKeys keys, int64_t first, int64_t count, int64_t* __restrict__ out) {
  const int64_t stride = int64_t(gridDim.x) * kThreads;
  for (int64_t j = int64_t(blockIdx.x) * kThreads + threadIdx.x; j < count; j += stride)
    out[j] = int64_t(feistel_perm(uint64_t(first + j), keys));
}

__device__ __forceinline__ bool is_positive(int32_t u, int32_t i, const int64_t* __restrict__ indptr,
                                            const int32_t* __restrict__ items) {
  int64_t lo = __ldg(indptr + u), hi = __ldg(indptr + u + 1);
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int32_t v = __ldg(items + mid);
    if (v == i) return true;
    if (v < i) lo = mid + 1; else hi = mid;
  }
  return false;
}

// One thread per output row: row j of the shuffled frame is source row perm(j) of concat(positives, negatives).
template <bool REJECT>
__global__ void __launch_bounds__(kThreads)
neumf_epoch_build_kernel(const int32_t* __restrict__ pos_users, const int32_t* __restrict__ pos_items,
                         uint32_t num_pos, const PermKeys keys, int64_t first, int64_t count, uint32_t seed,
                         uint32_t epoch, const int64_t* __restrict__ indptr, const int32_t* __restrict__ csr_items,
                         int32_t* __restrict__ users, int32_t* __restrict__ items, float* __restrict__ labels) {
  const int64_t stride = int64_t(gridDim.x) * kThreads;
  for (int64_t j = int64_t(blockIdx.x) * kThreads + threadIdx.x; j < count; j += stride) {
    const uint64_t s = feistel_perm(uint64_t(first + j), keys);
    int32_t u, i;
    float y;
    if (s < num_pos) {
      u = __ldg(pos_users + s);
      i = __ldg(pos_items + s);
      y = 1.0f;
    } else {
      const uint64_t idx = s - num_pos;
      y = 0.0f;
      int attempt = 0;
      while (true) {
        const uint4 w = philox4x32_10(make_uint4(uint32_t(idx), uint32_t(idx >> 32), uint32_t(attempt), kTagNeumf),
                                      seed, epoch);
        u = __ldg(pos_users + __umulhi(w.x, num_pos));
        i = __ldg(pos_items + __umulhi(w.y, num_pos));
        if (!REJECT || ++attempt == kMaxAttempts || !is_positive(u, i, indptr, csr_items)) break;
      }
    }
    users[j] = u;
    items[j] = i;
    labels[j] = y;
  }
}

// ---- factorisation ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t k) {      // murmur3 finaliser
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}

__global__ void __launch_bounds__(kThreads)
vocab_insert_kernel(const unsigned long long* __restrict__ keys, int64_t n, unsigned long long* __restrict__ tk,
                    int32_t* __restrict__ tv, uint64_t cap_mask, int32_t* __restrict__ slot_of) {
  const int64_t stride = int64_t(gridDim.x) * kThreads;
  for (int64_t pos = int64_t(blockIdx.x) * kThreads + threadIdx.x; pos < n; pos += stride) {
    const unsigned long long key = __ldg(keys + pos);
    uint64_t slot = mix64(key) & cap_mask;
    while (true) {
      unsigned long long prev = tk[slot];
      if (prev == kEmptyKey) prev = atomicCAS(tk + slot, kEmptyKey, key);
      if (prev == kEmptyKey || prev == key) break;
      slot = (slot + 1) & cap_mask;
    }
    atomicMin(tv + slot, int32_t(pos));
    slot_of[pos] = int32_t(slot);
  }
}

__global__ void __launch_bounds__(kThreads)
vocab_flag_kernel(const unsigned long long* __restrict__ tk, const int32_t* __restrict__ tv, int64_t cap,
                  uint32_t* __restrict__ rank) {
  const int64_t stride = int64_t(gridDim.x) * kThreads;
  for (int64_t s = int64_t(blockIdx.x) * kThreads + threadIdx.x; s < cap; s += stride)
    if (tk[s] != kEmptyKey) rank[tv[s]] = 1u;
}

// Exclusive prefix sum over n uint32 in three launches (tile sums, scan of the tile sums by one block, tile scans).
constexpr int kScanTile = 4096;     // 256 threads x 16
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t x, uint32_t* smem, uint32_t* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) smem[w] = incl;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < (kThreads / 32) ? smem[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    if (lane < (kThreads / 32)) smem[lane] = s;        // inclusive over warps
  }
  __syncthreads();
  const uint32_t base = w ? smem[w - 1] : 0u;
  if (total) *total = smem[kThreads / 32 - 1];
  __syncthreads();
  return base + incl - x;
}

__global__ void __launch_bounds__(kThreads)
scan_tile_sums_kernel(const uint32_t* __restrict__ v, int64_t n, uint32_t* __restrict__ sums) {
  __shared__ uint32_t smem[32];
  const int64_t base = int64_t(blockIdx.x) * kScanTile;
  uint32_t s = 0;
  for (int k = threadIdx.x; k < kScanTile; k += kThreads) s += (base + k < n) ? v[base + k] : 0u;
  uint32_t total;
  block_exclusive_scan(s, smem, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kThreads)
scan_sums_kernel(uint32_t* __restrict__ sums, int64_t n_tiles, int64_t* __restrict__ total_out) {
  __shared__ uint32_t smem[32];
  uint32_t carry = 0;
  for (int64_t base = 0; base < n_tiles; base += kThreads) {
    const int64_t k = base + threadIdx.x;
    const uint32_t x = k < n_tiles ? sums[k] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(x, smem, &total);
    if (k < n_tiles) sums[k] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *total_out = int64_t(carry);
}

__global__ void __launch_bounds__(kThreads)
scan_tiles_kernel(uint32_t* __restrict__ v, int64_t n, const uint32_t* __restrict__ sums) {
  __shared__ uint32_t smem[32];
  const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(threadIdx.x) * 16;
  uint32_t x[16], s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) { x[k] = (base + k < n) ? v[base + k] : 0u; s += x[k]; }
  uint32_t run = sums[blockIdx.x] + block_exclusive_scan(s, smem, nullptr);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (base + k < n) v[base + k] = run;
    run += x[k];
  }
}

__global__ void __launch_bounds__(kThreads)
vocab_assign_kernel(const unsigned long long* __restrict__ keys, int64_t n, const int32_t* __restrict__ tv,
                    const int32_t* __restrict__ slot_of, const uint32_t* __restrict__ rank, int32_t offset,
                    int32_t* __restrict__ ids, unsigned long long* __restrict__ vocab) {
  const int64_t stride = int64_t(gridDim.x) * kThreads;
  for (int64_t pos = int64_t(blockIdx.x) * kThreads + threadIdx.x; pos < n; pos += stride) {
    const int32_t f = tv[slot_of[pos]];
    const uint32_t r = rank[f];
    ids[pos] = int32_t(r) + offset;
    if (vocab && f == pos) vocab[r] = keys[pos];
  }
}

__global__ void __launch_bounds__(kThreads)
vocab_finalize_kernel(const unsigned long long* __restrict__ tk, int32_t* __restrict__ tv, int64_t cap,
                      const uint32_t* __restrict__ rank, int32_t offset) {
  const int64_t stride = int64_t(gridDim.x) * kThreads;
  for (int64_t s = int64_t(blockIdx.x) * kThreads + threadIdx.x; s < cap; s += stride)
    if (tk[s] != kEmptyKey) tv[s] = int32_t(rank[tv[s]]) + offset;
}

__global__ void __launch_bounds__(kThreads)
vocab_lookup_kernel(const unsigned long long* __restrict__ keys, int64_t n, const unsigned long long* __restrict__ tk,
                    const int32_t* __restrict__ tv, uint64_t cap_mask, int32_t oov, int32_t* __restrict__ ids) {
  const int64_t stride = int64_t(gridDim.x) * kThreads;
  for (int64_t pos = int64_t(blockIdx.x) * kThreads + threadIdx.x; pos < n; pos += stride) {
    const unsigned long long key = __ldg(keys + pos);
    uint64_t slot = mix64(key) & cap_mask;
    int32_t id = oov;
    while (true) {
      const unsigned long long k = __ldg(tk + slot);
      if (k == key) { id = __ldg(tv + slot); break; }
      if (k == kEmptyKey) break;
      slot = (slot + 1) & cap_mask;
    }
    ids[pos] = id;
  }
}

int grid_1d(const brk_ctx* ctx, int64_t n) {
  int64_t need = (n + kThreads - 1) / kThreads;
  const int64_t cap = int64_t(ctx->sm_count) * (2048 / kThreads);
  return int(need < 1 ? 1 : (need < cap ? need : cap));
}

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
inline int64_t scan_tiles(int64_t n) { return (n + kScanTile - 1) / kScanTile; }

}  // namespace

extern "C" int brk_epoch_permutation(brk_ctx* ctx, int64_t n, int64_t first, int64_t count, uint32_t seed,
                                     uint32_t epoch, uint32_t salt, int64_t* out, void* stream) {
  BRK_REQUIRE(ctx && (count == 0 || out), BRK_E_ARG, "brk_epoch_permutation: null argument");
  BRK_REQUIRE(n >= 1 && first >= 0 && count >= 0 && first + count <= n, BRK_E_ARG,
              "brk_epoch_permutation: n=%lld first=%lld count=%lld", (long long)n, (long long)first, (long long)count);
  if (count == 0) return 0;
  epoch_permutation_kernel<<<grid_1d(ctx, count), kThreads, 0, (cudaStream_t)stream>>>(
      make_perm_keys(uint64_t(n), seed, epoch, salt), first, count, out);
  BRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int brk_epoch_permutation_host(int64_t n, int64_t first, int64_t count, uint32_t seed, uint32_t epoch,
                                          uint32_t salt, int64_t* out_host) {
  BRK_REQUIRE(count == 0 || out_host, BRK_E_ARG, "brk_epoch_permutation_host: null argument");
  BRK_REQUIRE(n >= 1 && first >= 0 && count >= 0 && first + count <= n, BRK_E_ARG,
              "brk_epoch_permutation_host: n=%lld first=%lld count=%lld", (long long)n, (long long)first,
              (long long)count);
  const PermKeys keys = make_perm_keys(uint64_t(n), seed, epoch, salt);
  for (int64_t j = 0; j < count; ++j) out_host[j] = int64_t(feistel_perm(uint64_t(first + j), keys));
  return 0;
}

extern "C" int brk_neumf_epoch_build(brk_ctx* ctx, const int32_t* pos_users, const int32_t* pos_items,
                                     int64_t num_pos, int64_t n_neg, int64_t first, int64_t count, uint32_t seed,
                                     uint32_t epoch, int32_t reject, const int64_t* csr_indptr,
                                     const int32_t* csr_items, int32_t* users, int32_t* items, float* labels,
                                     void* stream) {
  BRK_REQUIRE(ctx && pos_users && pos_items && (count == 0 || (users && items && labels)), BRK_E_ARG,
              "brk_neumf_epoch_build: null argument");
  BRK_REQUIRE(num_pos > 0 && num_pos < (int64_t(1) << 32) && n_neg >= 0 && first >= 0 && count >= 0 &&
                  first + count <= num_pos + n_neg,
              BRK_E_ARG, "brk_neumf_epoch_build: num_pos=%lld n_neg=%lld first=%lld count=%lld", (long long)num_pos,
              (long long)n_neg, (long long)first, (long long)count);
  BRK_REQUIRE(!reject || (csr_indptr && csr_items), BRK_E_ARG,
              "brk_neumf_epoch_build: collision rejection needs the per-user sorted positive lists");
  if (count == 0) return 0;
  const PermKeys keys = make_perm_keys(uint64_t(num_pos + n_neg), seed, epoch, 0u);
  const int grid = grid_1d(ctx, count);
  if (reject)
    neumf_epoch_build_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        pos_users, pos_items, uint32_t(num_pos), keys, first, count, seed, epoch, csr_indptr, csr_items, users, items,
        labels);
  else
    neumf_epoch_build_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        pos_users, pos_items, uint32_t(num_pos), keys, first, count, seed, epoch, nullptr, nullptr, users, items,
        labels);
  BRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t brk_vocab_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

extern "C" int64_t brk_vocab_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return int64_t(align256(size_t(n) * 4) * 2 + align256(size_t(scan_tiles(n) + 1) * 4));
}

extern "C" int brk_vocab_build_u64(brk_ctx* ctx, const uint64_t* keys, int64_t n, int32_t offset,
                                   uint64_t* table_keys, int32_t* table_vals, int64_t capacity, int32_t* ids,
                                   uint64_t* vocab, int64_t* n_unique, void* workspace, void* stream) {
  BRK_REQUIRE(ctx && table_keys && table_vals && n_unique && (n == 0 || (keys && ids && workspace)), BRK_E_ARG,
              "brk_vocab_build_u64: null argument");
  BRK_REQUIRE(n >= 0 && n < int64_t(kNoPos) && capacity >= 2 * n && capacity >= 2 &&
                  (capacity & (capacity - 1)) == 0,
              BRK_E_ARG, "brk_vocab_build_u64: n=%lld capacity=%lld (need a power of two >= 2n)", (long long)n,
              (long long)capacity);
  cudaStream_t st = (cudaStream_t)stream;
  BRK_CUDA(cudaMemsetAsync(table_keys, 0xFF, size_t(capacity) * 8, st));
  BRK_CUDA(cudaMemsetAsync(table_vals, 0x7F, size_t(capacity) * 4, st));
  BRK_CUDA(cudaMemsetAsync(n_unique, 0, 8, st));
  if (n == 0) return 0;
  char* ws = static_cast<char*>(workspace);
  int32_t* slot_of = reinterpret_cast<int32_t*>(ws);
  uint32_t* rank = reinterpret_cast<uint32_t*>(ws + align256(size_t(n) * 4));
  uint32_t* sums = reinterpret_cast<uint32_t*>(ws + 2 * align256(size_t(n) * 4));
  BRK_CUDA(cudaMemsetAsync(rank, 0, size_t(n) * 4, st));
  auto* tk = reinterpret_cast<unsigned long long*>(table_keys);
  auto* k = reinterpret_cast<const unsigned long long*>(keys);
  vocab_insert_kernel<<<grid_1d(ctx, n), kThreads, 0, st>>>(k, n, tk, table_vals, uint64_t(capacity - 1), slot_of);
  vocab_flag_kernel<<<grid_1d(ctx, capacity), kThreads, 0, st>>>(tk, table_vals, capacity, rank);
  const int64_t tiles = scan_tiles(n);
  scan_tile_sums_kernel<<<int(tiles), kThreads, 0, st>>>(rank, n, sums);
  scan_sums_kernel<<<1, kThreads, 0, st>>>(sums, tiles, n_unique);
  scan_tiles_kernel<<<int(tiles), kThreads, 0, st>>>(rank, n, sums);
  vocab_assign_kernel<<<grid_1d(ctx, n), kThreads, 0, st>>>(k, n, table_vals, slot_of, rank, offset, ids,
                                                            reinterpret_cast<unsigned long long*>(vocab));
  vocab_finalize_kernel<<<grid_1d(ctx, capacity), kThreads, 0, st>>>(tk, table_vals, capacity, rank, offset);
  BRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int brk_vocab_lookup_u64(brk_ctx* ctx, const uint64_t* keys, int64_t n, const uint64_t* table_keys,
                                    const int32_t* table_vals, int64_t capacity, int32_t oov, int32_t* ids,
                                    void* stream) {
  BRK_REQUIRE(ctx && table_keys && table_vals && (n == 0 || (keys && ids)), BRK_E_ARG,
              "brk_vocab_lookup_u64: null argument");
  BRK_REQUIRE(n >= 0 && capacity >= 2 && (capacity & (capacity - 1)) == 0, BRK_E_ARG,
              "brk_vocab_lookup_u64: n=%lld capacity=%lld", (long long)n, (long long)capacity);
  if (n == 0) return 0;
  vocab_lookup_kernel<<<grid_1d(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned long long*>(keys), n, reinterpret_cast<const unsigned long long*>(table_keys),
      table_vals, uint64_t(capacity - 1), oov, ids);
  BRK_LAUNCH_CHECK();
  return 0;
}

// K1 gather_rows and K5 scatter_add_rows (mode 0: vector atomics; mode 1: per-CTA sort + segment-reduce).  HBM-bound byte movers:
//   gather  : 4d B read + 4d B written per row (+4 B id)
//   scatter : 4d B read (vals) + 2*4d B read-modify-write of the accumulator row per occurrence
// Mapping: a group of LPR = pow2(d/4) <= 32 lanes owns one row and moves it as float4 chunks, so a
// warp covers 32/LPR rows per request wave with fully coalesced 128-bit accesses; ROWS_PER_ITER
// rows per group are kept in flight to cover HBM latency.  Grid = SM count x resident CTAs.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kRowsPerIter = 4;

template <int LPR>
__global__ void __launch_bounds__(kThreads)
gather_rows_vec(const float* __restrict__ table, int d4, const int32_t* __restrict__ ids,
                int64_t n, float* __restrict__ out) {
  const int lane_in = threadIdx.x & (LPR - 1);
  const int64_t group = (int64_t(blockIdx.x) * kThreads + threadIdx.x) / LPR;
  const int64_t n_groups = int64_t(gridDim.x) * kThreads / LPR;
  const float4* __restrict__ t4 = reinterpret_cast<const float4*>(table);
  float4* __restrict__ o4 = reinterpret_cast<float4*>(out);
  for (int64_t b0 = group * kRowsPerIter; b0 < n; b0 += n_groups * kRowsPerIter) {
    int64_t row[kRowsPerIter];
#pragma unroll
    for (int r = 0; r < kRowsPerIter; ++r) row[r] = (b0 + r < n) ? int64_t(__ldg(ids + b0 + r)) : -1;
    for (int c = lane_in; c < d4; c += LPR) {
      float4 v[kRowsPerIter];
#pragma unroll
      for (int r = 0; r < kRowsPerIter; ++r)
        if (row[r] >= 0) v[r] = ldg_nc_f4(t4 + row[r] * d4 + c);
#pragma unroll
      for (int r = 0; r < kRowsPerIter; ++r)
        if (row[r] >= 0) stg_na_f4(o4 + (b0 + r) * d4 + c, v[r]);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
gather_rows_scalar(const float* __restrict__ table, int d, const int32_t* __restrict__ ids, int64_t n,
                   float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * kThreads + threadIdx.x) >> 5;
  const int64_t n_warps = int64_t(gridDim.x) * kThreads >> 5;
  for (int64_t b = warp; b < n; b += n_warps) {
    const int64_t row = __ldg(ids + b);
    for (int c = lane; c < d; c += 32) out[b * d + c] = __ldg(table + row * d + c);
  }
}

template <int LPR>
__global__ void __launch_bounds__(kThreads)
scatter_add_rows_vec(float* __restrict__ acc, int d4, const int32_t* __restrict__ ids, int64_t n,
                     const float* __restrict__ vals, uint32_t* __restrict__ touched) {
  const int lane_in = threadIdx.x & (LPR - 1);
  const int64_t group = (int64_t(blockIdx.x) * kThreads + threadIdx.x) / LPR;
  const int64_t n_groups = int64_t(gridDim.x) * kThreads / LPR;
  const float4* __restrict__ v4 = reinterpret_cast<const float4*>(vals);
  for (int64_t b0 = group * kRowsPerIter; b0 < n; b0 += n_groups * kRowsPerIter) {
    int64_t row[kRowsPerIter];
#pragma unroll
    for (int r = 0; r < kRowsPerIter; ++r) row[r] = (b0 + r < n) ? int64_t(__ldg(ids + b0 + r)) : -1;
    if (touched != nullptr && lane_in == 0) {
      // early plain read of the bitmask word; the bit is set with a fire-and-forget RED.OR
      uint32_t seen[kRowsPerIter];
#pragma unroll
      for (int r = 0; r < kRowsPerIter; ++r) seen[r] = row[r] >= 0 ? __ldg(touched + (row[r] >> 5)) : 0xffffffffu;
#pragma unroll
      for (int r = 0; r < kRowsPerIter; ++r) {
        const uint32_t bit = 1u << (row[r] & 31);
        if (row[r] >= 0 && !(seen[r] & bit)) atomicOr(touched + (row[r] >> 5), bit);
      }
    }
    for (int c = lane_in; c < d4; c += LPR) {
      float4 v[kRowsPerIter];
#pragma unroll
      for (int r = 0; r < kRowsPerIter; ++r)
        if (row[r] >= 0) v[r] = ldg_nc_f4(v4 + (b0 + r) * d4 + c);
#pragma unroll
      for (int r = 0; r < kRowsPerIter; ++r)
        if (row[r] >= 0) red_add_f4(acc + (row[r] * d4 + c) * 4, v[r]);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
scatter_add_rows_scalar(float* __restrict__ acc, int d, const int32_t* __restrict__ ids, int64_t n,
                        const float* __restrict__ vals, uint32_t* __restrict__ touched) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * kThreads + threadIdx.x) >> 5;
  const int64_t n_warps = int64_t(gridDim.x) * kThreads >> 5;
  for (int64_t b = warp; b < n; b += n_warps) {
    const int64_t row = __ldg(ids + b);
    for (int c = lane; c < d; c += 32) atomicAdd(acc + row * d + c, __ldg(vals + b * d + c));
    if (touched != nullptr && lane == 0) atomicOr(touched + (row >> 5), 1u << (row & 31));
  }
}

// Mode 1, sort-and-segment-reduce: a CTA sorts a tile of kSortTile (id, position) pairs in shared memory (bitonic,
// by id then position), then lane groups walk aligned chunks of kRun sorted entries, add the value rows of equal
// ids in registers and send ONE 16-byte RED per run and column chunk.  With heavily repeated ids (small hot
// tables: a 6040-row table hit by 1 M ids) this cuts the same-address REDs that serialise in L2 by up to kRun;
// with mostly distinct ids it only adds the sort, so the caller picks the mode by the duplicate factor n / rows.
constexpr int kSortTile = 1024;
constexpr int kRun = 16;

template <int LPR>
__global__ void __launch_bounds__(kThreads)
scatter_add_rows_sorted(float* __restrict__ acc, int d4, const int32_t* __restrict__ ids, int64_t n,
                        const float* __restrict__ vals, uint32_t* __restrict__ touched) {
  __shared__ int32_t key[kSortTile];
  __shared__ int32_t pos[kSortTile];
  constexpr int GROUPS = kThreads / LPR;
  const int t = threadIdx.x, grp = t / LPR, lane_in = t % LPR;
  const float4* __restrict__ v4 = reinterpret_cast<const float4*>(vals);
  const int64_t n_tiles = (n + kSortTile - 1) / kSortTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kSortTile;
    for (int i = t; i < kSortTile; i += kThreads) {
      key[i] = (base + i < n) ? __ldg(ids + base + i) : 0x7fffffff;
      pos[i] = i;
    }
    __syncthreads();
    // bitonic sort of (key, pos), ascending; pos breaks ties so the order (and the fp32 sum order) is reproducible
    for (int k = 2; k <= kSortTile; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = t; i < kSortTile; i += kThreads) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const int32_t ka = key[i], kb = key[ixj], pa = pos[i], pb = pos[ixj];
            const bool up = (i & k) == 0;
            const bool gt = ka > kb || (ka == kb && pa > pb);
            if (gt == up) { key[i] = kb; key[ixj] = ka; pos[i] = pb; pos[ixj] = pa; }
          }
        }
        __syncthreads();
      }
    }
    // segment-reduce: chunk c = sorted entries [c*kRun, (c+1)*kRun)
    for (int c = grp; c < kSortTile / kRun; c += GROUPS) {
      const int i0 = c * kRun;
      if (key[i0] == 0x7fffffff) break;                      // padding only from here on (sorted last)
      for (int cc = lane_in; cc < d4; cc += LPR) {
        float4 v[kRun];
#pragma unroll
        for (int r = 0; r < kRun; ++r)
          v[r] = key[i0 + r] != 0x7fffffff ? ldg_nc_f4(v4 + (base + pos[i0 + r]) * d4 + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 run = v[0];
        int32_t cur = key[i0];
#pragma unroll
        for (int r = 1; r < kRun; ++r) {
          const int32_t kr = key[i0 + r];
          if (kr != cur) {
            red_add_f4(acc + (int64_t(cur) * d4 + cc) * 4, run);
            if (kr == 0x7fffffff) { cur = kr; break; }
            cur = kr; run = v[r];
          } else {
            run.x += v[r].x; run.y += v[r].y; run.z += v[r].z; run.w += v[r].w;
          }
        }
        if (cur != 0x7fffffff) red_add_f4(acc + (int64_t(cur) * d4 + cc) * 4, run);
      }
      if (touched != nullptr && lane_in == 0) {
        int32_t prev = -1;
        for (int r = 0; r < kRun; ++r) {
          const int32_t kr = key[i0 + r];
          if (kr == 0x7fffffff) break;
          if (kr != prev) { atomicOr(touched + (kr >> 5), 1u << (kr & 31)); prev = kr; }
        }
      }
    }
    __syncthreads();
  }
}

int grid_for(const brk_ctx* ctx, int64_t work_items, int items_per_block) {
  int64_t need = (work_items + items_per_block - 1) / items_per_block;
  int64_t cap = int64_t(ctx->sm_count) * (2048 / kThreads);
  if (need < 1) need = 1;
  return int(need < cap ? need : cap);
}

}  // namespace

extern "C" int brk_gather_rows(brk_ctx* ctx, const float* table, int64_t rows, int32_t d,
                               const int32_t* ids, int64_t n, float* out, void* stream) {
  BRK_REQUIRE(ctx && table && (n == 0 || (ids && out)), BRK_E_ARG, "brk_gather_rows: null argument");
  BRK_REQUIRE(rows > 0 && d > 0 && n >= 0, BRK_E_ARG, "brk_gather_rows: rows=%lld d=%d n=%lld",
              (long long)rows, d, (long long)n);
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if ((d & 3) == 0 && brk_aligned16(table) && brk_aligned16(out)) {
    const int d4 = d >> 2;
    const int lpr = brk_lanes_per_row(d4);
    const int grid = grid_for(ctx, n, (kThreads / lpr) * kRowsPerIter);
    switch (lpr) {
      case 1:  gather_rows_vec<1><<<grid, kThreads, 0, st>>>(table, d4, ids, n, out); break;
      case 2:  gather_rows_vec<2><<<grid, kThreads, 0, st>>>(table, d4, ids, n, out); break;
      case 4:  gather_rows_vec<4><<<grid, kThreads, 0, st>>>(table, d4, ids, n, out); break;
      case 8:  gather_rows_vec<8><<<grid, kThreads, 0, st>>>(table, d4, ids, n, out); break;
      case 16: gather_rows_vec<16><<<grid, kThreads, 0, st>>>(table, d4, ids, n, out); break;
      default: gather_rows_vec<32><<<grid, kThreads, 0, st>>>(table, d4, ids, n, out); break;
    }
  } else {
    const int grid = grid_for(ctx, n, kThreads / 32);
    gather_rows_scalar<<<grid, kThreads, 0, st>>>(table, d, ids, n, out);
  }
  BRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int brk_scatter_add_rows(brk_ctx* ctx, float* acc, int64_t rows, int32_t d,
                                    const int32_t* ids, int64_t n, const float* vals,
                                    uint32_t* touched, int32_t mode, void* stream) {
  BRK_REQUIRE(ctx && acc && (n == 0 || (ids && vals)), BRK_E_ARG, "brk_scatter_add_rows: null argument");
  BRK_REQUIRE(rows > 0 && d > 0 && n >= 0, BRK_E_ARG, "brk_scatter_add_rows: rows=%lld d=%d n=%lld",
              (long long)rows, d, (long long)n);
  BRK_REQUIRE(mode == 0 || mode == 1, BRK_E_ARG, "brk_scatter_add_rows: mode %d (0 = vector atomics, 1 = sort-and-segment-reduce)", mode);
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 1 && (d & 3) == 0 && d <= 128 && brk_aligned16(acc) && brk_aligned16(vals)) {
    const int d4 = d >> 2;
    const int lpr = brk_lanes_per_row(d4);
    const int64_t tiles = (n + kSortTile - 1) / kSortTile;
    const int64_t cap = int64_t(ctx->sm_count) * 8;
    const int grid = int(tiles < cap ? tiles : cap);
    switch (lpr) {
      case 1:  scatter_add_rows_sorted<1><<<grid, kThreads, 0, st>>>(acc, d4, ids, n, vals, touched); break;
      case 2:  scatter_add_rows_sorted<2><<<grid, kThreads, 0, st>>>(acc, d4, ids, n, vals, touched); break;
      case 4:  scatter_add_rows_sorted<4><<<grid, kThreads, 0, st>>>(acc, d4, ids, n, vals, touched); break;
      case 8:  scatter_add_rows_sorted<8><<<grid, kThreads, 0, st>>>(acc, d4, ids, n, vals, touched); break;
      case 16: scatter_add_rows_sorted<16><<<grid, kThreads, 0, st>>>(acc, d4, ids, n, vals, touched); break;
      default: scatter_add_rows_sorted<32><<<grid, kThreads, 0, st>>>(acc, d4, ids, n, vals, touched); break;
    }
    BRK_LAUNCH_CHECK();
    return 0;
  }
  if ((d & 3) == 0 && brk_aligned16(acc) && brk_aligned16(vals)) {
    const int d4 = d >> 2;
    const int lpr = brk_lanes_per_row(d4);
    const int grid = grid_for(ctx, n, (kThreads / lpr) * kRowsPerIter);
    switch (lpr) {
      case 1:  scatter_add_rows_vec<1><<<grid, kThreads, 0, st>>>(acc, d4, ids, n, vals, touched); break;
      case 2:  scatter_add_rows_vec<2><<<grid, kThreads, 0, st>>>(acc, d4, ids, n, vals, touched); break;
      case 4:  scatter_add_rows_vec<4><<<grid, kThreads, 0, st>>>(acc, d4, ids, n, vals, touched); break;
      case 8:  scatter_add_rows_vec<8><<<grid, kThreads, 0, st>>>(acc, d4, ids, n, vals, touched); break;
      case 16: scatter_add_rows_vec<16><<<grid, kThreads, 0, st>>>(acc, d4, ids, n, vals, touched); break;
      default: scatter_add_rows_vec<32><<<grid, kThreads, 0, st>>>(acc, d4, ids, n, vals, touched); break;
    }
  } else {
    const int grid = grid_for(ctx, n, kThreads / 32);
    scatter_add_rows_scalar<<<grid, kThreads, 0, st>>>(acc, d, ids, n, vals, touched);
  }
  BRK_LAUNCH_CHECK();
  return 0;
}


// ---- row-sharded tables over NVLink peer memory: the "all-to-all of looked-up rows" and "of row gradients" as ops -------
// Row r of the table lives on rank r % world at local row r / world (include/brk_b200.h, brk_shards).  The gather reads
// peer rows with ordinary 16-byte loads, the scatter-add sends 16-byte REDs into the owner's accumulator and marks the
// owner's touched bit: the exchange IS the memory instruction, there is no staging buffer and no separate collective.
// (The fused NeuMF / BPR kernels do the same inline; these are the stand-alone forms.)
namespace {
__global__ void __launch_bounds__(256) gather_rows_sharded_kernel(brk_shards sh, int d4, const int32_t* __restrict__ ids, int64_t n,
                                                                 float4* __restrict__ out, int lpr) {
  const int64_t gt = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t row = gt / lpr; const int c = int(gt % lpr);
  if (row >= n) return;
  const int64_t id = __ldg(ids + row);
  const int o = int(id % sh.world); const int64_t l = id / sh.world;
  const float4* src = reinterpret_cast<const float4*>(sh.w[o]) + l * d4;
  for (int k = c; k < d4; k += lpr) out[row * d4 + k] = ldg_nc_f4(src + k);
}
__global__ void __launch_bounds__(256) scatter_add_rows_sharded_kernel(brk_shards sh, int d4, const int32_t* __restrict__ ids, int64_t n,
                                                                      const float4* __restrict__ vals, int lpr) {
  const int64_t gt = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t row = gt / lpr; const int c = int(gt % lpr);
  if (row >= n) return;
  const int64_t id = __ldg(ids + row);
  const int o = int(id % sh.world); const int64_t l = id / sh.world;
  float* dst = sh.g[o] + l * d4 * 4;
  for (int k = c; k < d4; k += lpr) red_add_f4(dst + 4 * k, __ldg(vals + row * d4 + k));
  if (c == 0 && sh.touched[o] != nullptr) atomicOr(sh.touched[o] + (l >> 5), 1u << (l & 31));
}
}  // namespace

extern "C" int brk_gather_rows_sharded(brk_ctx* ctx, const brk_shards* sh, int32_t d, const int32_t* ids, int64_t n, float* out,
                                       void* stream) {
  BRK_REQUIRE(ctx && sh && ids && out, BRK_E_ARG, "brk_gather_rows_sharded: null argument");
  BRK_REQUIRE(sh->world >= 1 && sh->world <= BRK_MAX_PEERS && d > 0 && (d & 3) == 0 && n >= 0, BRK_E_ARG,
              "brk_gather_rows_sharded: world=%d d=%d (d must be a multiple of 4)", sh->world, d);
  for (int p = 0; p < sh->world; ++p)
    BRK_REQUIRE(sh->w[p] && brk_aligned16(sh->w[p]), BRK_E_ALIGN, "brk_gather_rows_sharded: shard %d missing or not 16-byte aligned", p);
  BRK_REQUIRE(brk_aligned16(out), BRK_E_ALIGN, "brk_gather_rows_sharded: out is not 16-byte aligned");
  if (n == 0) return 0;
  const int d4 = d / 4, lpr = brk_lanes_per_row(d4);
  const int64_t threads = n * lpr;
  gather_rows_sharded_kernel<<<unsigned((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*sh, d4, ids, n, reinterpret_cast<float4*>(out), lpr);
  BRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int brk_scatter_add_rows_sharded(brk_ctx* ctx, const brk_shards* sh, int32_t d, const int32_t* ids, int64_t n,
                                            const float* values, void* stream) {
  BRK_REQUIRE(ctx && sh && ids && values, BRK_E_ARG, "brk_scatter_add_rows_sharded: null argument");
  BRK_REQUIRE(sh->world >= 1 && sh->world <= BRK_MAX_PEERS && d > 0 && (d & 3) == 0 && n >= 0, BRK_E_ARG,
              "brk_scatter_add_rows_sharded: world=%d d=%d (d must be a multiple of 4)", sh->world, d);
  for (int p = 0; p < sh->world; ++p)
    BRK_REQUIRE(sh->g[p] && brk_aligned16(sh->g[p]), BRK_E_ALIGN, "brk_scatter_add_rows_sharded: accumulator %d missing or not 16-byte aligned", p);
  BRK_REQUIRE(brk_aligned16(values), BRK_E_ALIGN, "brk_scatter_add_rows_sharded: values are not 16-byte aligned");
  if (n == 0) return 0;
  const int d4 = d / 4, lpr = brk_lanes_per_row(d4);
  const int64_t threads = n * lpr;
  scatter_add_rows_sharded_kernel<<<unsigned((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*sh, d4, ids, n,
                                                                                                       reinterpret_cast<const float4*>(values), lpr);
  BRK_LAUNCH_CHECK();
  return 0;
}

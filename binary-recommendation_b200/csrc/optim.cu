// K6: fused optimizers consuming the dense gradient accumulators written by the fwd/bwd kernels.
//   adam_dense_keras : exact Keras Adam; every element moves (dense-equivalent for sparse grads).
//                      Traffic per element: w,m,v read+write + g read + g zeroed = 32 B.
//   adam_rows        : lazy Adam over rows flagged in the touched bitmask.  Per touched row:
//                      6*4d (w,m,v r+w) + 4d (g read) + 4d (g zeroed) B; the bitmask scan adds rows/8 B.
//   adagrad_rows     : Keras sparse Adagrad over touched rows: 4*4d + 4d + 4d B per touched row.
// Reference call sites: Adam(1e-3) /root/reference/src/models/NeuMFModel.py:89, BPRModel.py:70,
// bpr.py:201; Adam(lr=0.005) trainers/NFC_plain.py:153; "Adagrad" 0.1 trainers/twoTower.py:278-279.
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxTabs = 16;

struct TabList {
  brk_table t[kMaxTabs];
  int n;
};

struct AdamOp {
  float alpha, b1, b2, omb1, omb2, eps;
  __device__ __forceinline__ void operator()(float& w, float& m, float& v, float g) const {
    m = b1 * m + omb1 * g;
    v = b2 * v + omb2 * g * g;
    w -= alpha * m / (sqrtf(v) + eps);
  }
};
struct AdagradOp {
  float lr, eps;
  __device__ __forceinline__ void operator()(float& w, float& acc, float& /*unused*/, float g) const {
    acc += g * g;
    w -= lr * g / (sqrtf(acc) + eps);
  }
};

template <class Op, bool HAS_V>
__device__ __forceinline__ void apply4(const Op& op, float* w, float* m, float* v, float* g, int64_t i4) {
  float4 w4 = reinterpret_cast<float4*>(w)[i4];
  float4 m4 = reinterpret_cast<float4*>(m)[i4];
  float4 v4 = HAS_V ? reinterpret_cast<float4*>(v)[i4] : make_float4(0, 0, 0, 0);
  const float4 g4 = reinterpret_cast<float4*>(g)[i4];
  op(w4.x, m4.x, v4.x, g4.x);
  op(w4.y, m4.y, v4.y, g4.y);
  op(w4.z, m4.z, v4.z, g4.z);
  op(w4.w, m4.w, v4.w, g4.w);
  reinterpret_cast<float4*>(w)[i4] = w4;
  reinterpret_cast<float4*>(m)[i4] = m4;
  if (HAS_V) reinterpret_cast<float4*>(v)[i4] = v4;
  reinterpret_cast<float4*>(g)[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
}
template <class Op, bool HAS_V>
__device__ __forceinline__ void apply1(const Op& op, float* w, float* m, float* v, float* g, int64_t i) {
  float ww = w[i], mm = m[i], vv = HAS_V ? v[i] : 0.f;
  op(ww, mm, vv, g[i]);
  w[i] = ww; m[i] = mm;
  if (HAS_V) v[i] = vv;
  g[i] = 0.f;
}

// Optimizer state in device memory (see brk_b200.h): state[0] = t as int64, state[1] = beta1^t and
// state[2] = beta2^t as doubles (running products; no pow() on the device).
__device__ __forceinline__ float adam_alpha(const brk_adam_hyper& h, const int64_t* state) {
  __shared__ float s_alpha;
  if (threadIdx.x == 0) {
    const double* pw = reinterpret_cast<const double*>(state);
    const double p1 = pw[1] * double(h.beta1), p2 = pw[2] * double(h.beta2);
    s_alpha = float(double(h.lr) * sqrt(1.0 - p2) / (1.0 - p1));
  }
  __syncthreads();
  return s_alpha;
}

__device__ __forceinline__ void advance_step_last_block(int64_t* state, const brk_adam_hyper& h,
                                                        unsigned int* ticket, int advance) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      if (advance) {
        double* pw = reinterpret_cast<double*>(state);
        state[0] += 1;
        pw[1] *= double(h.beta1);
        pw[2] *= double(h.beta2);
      }
      *ticket = 0u;
      __threadfence();
    }
  }
}

template <class Op, bool HAS_V>
__device__ __forceinline__ void dense_pass(const TabList& tl, const Op& op) {
  const int64_t tid = int64_t(blockIdx.x) * kThreads + threadIdx.x;
  const int64_t nthr = int64_t(gridDim.x) * kThreads;
  for (int k = 0; k < tl.n; ++k) {
    const brk_table& t = tl.t[k];
    const int64_t numel = t.rows * t.d;
    const bool vec = brk_aligned16(t.w) && brk_aligned16(t.m) && brk_aligned16(t.g) && (!HAS_V || brk_aligned16(t.v));
    const int64_t n4 = vec ? (numel >> 2) : 0;
    for (int64_t i = tid; i < n4; i += nthr) apply4<Op, HAS_V>(op, t.w, t.m, t.v, t.g, i);
    for (int64_t i = n4 * 4 + tid; i < numel; i += nthr) apply1<Op, HAS_V>(op, t.w, t.m, t.v, t.g, i);
    if (t.touched != nullptr) {
      const int64_t nwords = (t.rows + 31) >> 5;
      for (int64_t i = tid; i < nwords; i += nthr) t.touched[i] = 0u;
    }
  }
}

// Row-sparse pass: warps scan the touched bitmask 32 words at a time, clear what they read, expand the set bits
// into a per-warp row list in shared memory and update those rows; a row is handled by `lpr` lanes as float4
// chunks and every lane group keeps kRowsInFlight rows' loads in flight (the pass is DRAM-latency bound: at
// BASELINE.json configs[3] densities a warp finds 3-30 rows per 1024 scanned).
// Rows in flight per lane group.  4 costs 128 registers per thread (2 CTAs per SM); 2 would fit 3 CTAs per SM but
// measured slower at every touched-row count (65 k of 20 M rows: 42 vs 37 us; 1 M: 397 vs 384 us).
constexpr int kRowsInFlight = 4;

template <class Op, bool HAS_V>
__device__ __forceinline__ void rows_pass(const TabList& tl, const Op& op) {
  __shared__ int32_t s_rows[kThreads / 32][1024 + 32 * kRowsInFlight];   // touched rows waiting for the warp's row loop
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t(blockIdx.x) * kThreads + threadIdx.x) >> 5;
  const int64_t n_warps = (int64_t(gridDim.x) * kThreads) >> 5;
  int32_t* my_rows = s_rows[wib];
  for (int k = 0; k < tl.n; ++k) {
    const brk_table& t = tl.t[k];
    const bool vec = (t.d & 3) == 0 && brk_aligned16(t.w) && brk_aligned16(t.m) && brk_aligned16(t.g) &&
                     (!HAS_V || brk_aligned16(t.v));
    const int chunks = vec ? (t.d >> 2) : t.d;   // per-row work items (float4 or float)
    int lpr = 1;
    while (lpr < chunks && lpr < 32) lpr <<= 1;
    const int gpw = 32 / lpr;
    const int grp = lane / lpr, lane_in = lane % lpr;
    const int64_t nwords = (t.rows + 31) >> 5;
    // words per warp and pass: 32 for big tables, fewer when the table is too small to give every warp something
    // to do.  The wpw words of one pass are taken `stride` words apart (lane l reads word l * stride + slot), not
    // next to each other: with skewed ids the touched rows crowd into a few neighbouring words (ids drawn as
    // floor(rows * r^3) put ~700 of a 65 536 batch's rows into the first 1024), and a warp that owned that whole
    // neighbourhood worked alone for 0.6 ms while the rest of the grid had finished (20 M-row tables: 601 -> the
    // uniform-id time).  Spread out, no warp gets more than one word of any neighbourhood.
    int wpw = 32;
    while (wpw > 1 && nwords < n_warps * wpw) wpw >>= 1;
    const int64_t stride = (nwords + wpw - 1) / wpw;            // slots: [0, stride)
    // The bitmask words of slot s+1 are loaded before the rows of slot s are touched (one DRAM round trip per slot
    // instead of two), and sparse slots are merged: rows pile up in the list until one full round of the row loop
    // (gpw * kRowsInFlight rows) is there, so a warp that finds 3 rows per slot pays one row round trip per 3 slots.
    auto load_word = [&](int64_t slot_) -> uint32_t {
      const int64_t widx_ = int64_t(lane) * stride + slot_;
      return (slot_ < stride && lane < wpw && widx_ < nwords) ? t.touched[widx_] : 0u;
    };
    int have = 0;                                               // rows waiting in the list
    uint32_t word_next = load_word(warp);
    for (int64_t slot = warp; slot < stride; slot += n_warps) {
      const uint32_t word = word_next;
      const int64_t widx = int64_t(lane) * stride + slot;
      word_next = load_word(slot + n_warps);
      if (word) t.touched[widx] = 0u;
      const bool last_slot = slot + n_warps >= stride;
      if (__ballot_sync(0xffffffffu, word != 0u) == 0u && !(last_slot && have > 0)) continue;
      // exclusive prefix sum of the per-lane popcounts -> each lane expands its word into the list
      const int cnt = __popc(word);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      int pos = have + incl - cnt;
      have += __shfl_sync(0xffffffffu, incl, 31);
      uint32_t w = word;
      while (w) {
        const int bit = __ffs(w) - 1;
        w &= w - 1;
        my_rows[pos++] = int32_t(widx * 32 + bit);              // absolute row (rows < 2^31)
      }
      if (have < gpw * kRowsInFlight && !last_slot) continue;   // not a full round yet: take the next slot too
      const int total = have;
      have = 0;
      __syncwarp();
      if (vec && chunks <= lpr) {
        // one float4 per lane and row: kRowsInFlight rows per group, loads first, then compute and store
        for (int j0 = 0; j0 < total; j0 += gpw * kRowsInFlight) {
          float4 w4[kRowsInFlight], m4[kRowsInFlight], v4[kRowsInFlight], g4[kRowsInFlight];
          int64_t idx[kRowsInFlight];
#pragma unroll
          for (int q = 0; q < kRowsInFlight; ++q) {
            const int j = j0 + q * gpw + grp;
            idx[q] = -1;
            if (j < total && lane_in < chunks) {
              idx[q] = int64_t(my_rows[j]) * chunks + lane_in;
              w4[q] = reinterpret_cast<const float4*>(t.w)[idx[q]];
              m4[q] = reinterpret_cast<const float4*>(t.m)[idx[q]];
              if (HAS_V) v4[q] = reinterpret_cast<const float4*>(t.v)[idx[q]];
              g4[q] = reinterpret_cast<const float4*>(t.g)[idx[q]];
            }
          }
#pragma unroll
          for (int q = 0; q < kRowsInFlight; ++q) {
            if (idx[q] >= 0) {
              float4 vv = HAS_V ? v4[q] : make_float4(0.f, 0.f, 0.f, 0.f);
              op(w4[q].x, m4[q].x, vv.x, g4[q].x); op(w4[q].y, m4[q].y, vv.y, g4[q].y);
              op(w4[q].z, m4[q].z, vv.z, g4[q].z); op(w4[q].w, m4[q].w, vv.w, g4[q].w);
              reinterpret_cast<float4*>(t.w)[idx[q]] = w4[q];
              reinterpret_cast<float4*>(t.m)[idx[q]] = m4[q];
              if (HAS_V) reinterpret_cast<float4*>(t.v)[idx[q]] = vv;
              reinterpret_cast<float4*>(t.g)[idx[q]] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
      } else {
        for (int j = grp; j < total; j += gpw) {
          const int64_t row = my_rows[j];
          for (int c = lane_in; c < chunks; c += lpr) {
            if (vec) apply4<Op, HAS_V>(op, t.w, t.m, t.v, t.g, row * chunks + c);
            else     apply1<Op, HAS_V>(op, t.w, t.m, t.v, t.g, row * chunks + c);
          }
        }
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(kThreads)
adam_dense_kernel(TabList tl, brk_adam_hyper h, int64_t* step_dev, unsigned int* ticket, int advance) {
  AdamOp op;
  op.alpha = adam_alpha(h, step_dev);
  op.b1 = h.beta1; op.b2 = h.beta2; op.omb1 = 1.0f - h.beta1; op.omb2 = 1.0f - h.beta2; op.eps = h.eps;
  dense_pass<AdamOp, true>(tl, op);
  advance_step_last_block(step_dev, h, ticket, advance);
}

__global__ void __launch_bounds__(kThreads, 2)
adam_rows_kernel(TabList tl, brk_adam_hyper h, int64_t* step_dev, unsigned int* ticket, int advance) {
  AdamOp op;
  op.alpha = adam_alpha(h, step_dev);
  op.b1 = h.beta1; op.b2 = h.beta2; op.omb1 = 1.0f - h.beta1; op.omb2 = 1.0f - h.beta2; op.eps = h.eps;
  rows_pass<AdamOp, true>(tl, op);
  advance_step_last_block(step_dev, h, ticket, advance);
}

__global__ void __launch_bounds__(kThreads, 2) adagrad_rows_kernel(TabList tl, float lr, float eps) {
  AdagradOp op{lr, eps};
  rows_pass<AdagradOp, false>(tl, op);
}
__global__ void __launch_bounds__(kThreads) adagrad_dense_kernel(TabList tl, float lr, float eps) {
  AdagradOp op{lr, eps};
  dense_pass<AdagradOp, false>(tl, op);
}

int pack(const char* who, const brk_table* tabs, int32_t n_tabs, bool need_v, bool need_touched, TabList* out,
         int64_t* work_items) {
  BRK_REQUIRE(tabs != nullptr && n_tabs > 0 && n_tabs <= kMaxTabs, BRK_E_ARG, "%s: n_tabs=%d (1..%d)", who,
              n_tabs, kMaxTabs);
  int64_t work = 0;
  for (int k = 0; k < n_tabs; ++k) {
    const brk_table& t = tabs[k];
    BRK_REQUIRE(t.w && t.g && t.m && (!need_v || t.v), BRK_E_ARG, "%s: table %d lacks w/g/m/v", who, k);
    BRK_REQUIRE(!need_touched || t.touched, BRK_E_ARG, "%s: table %d lacks the touched bitmask", who, k);
    BRK_REQUIRE(t.rows > 0 && t.d > 0, BRK_E_ARG, "%s: table %d rows=%lld d=%d", who, k, (long long)t.rows, t.d);
    out->t[k] = t;
    work += need_touched ? (t.rows + 31) / 32 * 32 : (t.rows * t.d + 3) / 4;  // threads wanted
  }
  out->n = n_tabs;
  *work_items = work;
  return 0;
}

// Row passes: a persistent grid of exactly the CTAs that are resident at once (the 8-per-SM cap of grid_for ran
// the 2-per-SM row kernels in four waves, each with its own latency tail: 65 k touched rows of 20 M took 50 us,
// 37 us now).
template <class K>
int rows_grid(const brk_ctx* ctx, K kernel, int64_t threads_wanted, int* grid) {
  int occ = 0;
  BRK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, 0));
  const int64_t need = (threads_wanted + kThreads - 1) / kThreads, cap = int64_t(ctx->sm_count) * (occ < 1 ? 1 : occ);
  *grid = int(need < 1 ? 1 : (need < cap ? need : cap));
  return 0;
}
int grid_for(const brk_ctx* ctx, int64_t threads_wanted) {
  int64_t need = (threads_wanted + kThreads - 1) / kThreads;
  const int64_t cap = int64_t(ctx->sm_count) * (2048 / kThreads);
  return int(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace

extern "C" int brk_adam_dense_keras(brk_ctx* ctx, const brk_table* tabs, int32_t n_tabs, brk_adam_hyper h,
                                    int64_t* step_dev, int32_t advance_step, void* stream) {
  BRK_REQUIRE(ctx && step_dev, BRK_E_ARG, "brk_adam_dense_keras: null argument");
  TabList tl; int64_t work;
  if (int rc = pack("brk_adam_dense_keras", tabs, n_tabs, true, false, &tl, &work)) return rc;
  adam_dense_kernel<<<grid_for(ctx, work), kThreads, 0, (cudaStream_t)stream>>>(tl, h, step_dev, ctx->tickets + 1,
                                                                               advance_step);
  BRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int brk_adam_rows(brk_ctx* ctx, const brk_table* tabs, int32_t n_tabs, brk_adam_hyper h,
                             int64_t* step_dev, int32_t advance_step, void* stream) {
  BRK_REQUIRE(ctx && step_dev, BRK_E_ARG, "brk_adam_rows: null argument");
  TabList tl; int64_t work;
  if (int rc = pack("brk_adam_rows", tabs, n_tabs, true, true, &tl, &work)) return rc;
  int grid = 1;
  if (int rc = rows_grid(ctx, adam_rows_kernel, work, &grid)) return rc;
  adam_rows_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(tl, h, step_dev, ctx->tickets + 2, advance_step);
  BRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int brk_adagrad_rows(brk_ctx* ctx, const brk_table* tabs, int32_t n_tabs, float lr, float eps,
                                void* stream) {
  BRK_REQUIRE(ctx, BRK_E_ARG, "brk_adagrad_rows: null context");
  TabList tl; int64_t work;
  if (int rc = pack("brk_adagrad_rows", tabs, n_tabs, false, true, &tl, &work)) return rc;
  int grid = 1;
  if (int rc = rows_grid(ctx, adagrad_rows_kernel, work, &grid)) return rc;
  adagrad_rows_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(tl, lr, eps);
  BRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int brk_adagrad_dense(brk_ctx* ctx, const brk_table* tabs, int32_t n_tabs, float lr, float eps,
                                 void* stream) {
  BRK_REQUIRE(ctx, BRK_E_ARG, "brk_adagrad_dense: null context");
  TabList tl; int64_t work;
  if (int rc = pack("brk_adagrad_dense", tabs, n_tabs, false, false, &tl, &work)) return rc;
  adagrad_dense_kernel<<<grid_for(ctx, work), kThreads, 0, (cudaStream_t)stream>>>(tl, lr, eps);
  BRK_LAUNCH_CHECK();
  return 0;
}

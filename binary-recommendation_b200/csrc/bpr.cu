// Fused BPR triplet forward + backward (K1+K4+K5 in one pass; no activations ever reach HBM).
// Reference graph: /root/reference/src/models/BPRModel.py:49-74, loss :124-144 (1 - sigmoid).
// Per triplet: 3 rows gathered (3*4d B) and 3 row gradients reduced into the dense accumulators
// (3*4d B of RED traffic) => 1536 B of algorithmic traffic at d = 64.
// Mapping: a group of LPR lanes owns a triplet, each lane holds NCH float4 chunks of u, p and n in
// registers; the dot product is a group shuffle-reduction; gradients leave as 16-byte vector REDs.
#include "common.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace {

constexpr int kThreads = 256;

// Touched-row bitmask: the word is read early (plain load, issued with the row loads; a stale
// value only costs a redundant RED) and the bit is set with a fire-and-forget RED.OR.
__device__ __forceinline__ uint32_t peek_touched(const uint32_t* touched, int64_t row) {
  return touched ? __ldg(touched + (row >> 5)) : 0xffffffffu;
}
__device__ __forceinline__ void mark_touched(uint32_t* touched, int64_t row, uint32_t seen) {
  const uint32_t bit = 1u << (row & 31);
  if (!(seen & bit)) atomicOr(touched + (row >> 5), bit);
}

// Adds this block's partial loss; the last block to arrive publishes the mean and resets the slot.
__device__ __forceinline__ void finish_loss(double part, double inv_batch, double* acc, unsigned int* ticket,
                                            float* loss_out) {
  __shared__ double red[32];
  const double tot = block_sum_double(part, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc, tot);
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      __threadfence();
      const double sum = atomicAdd(acc, 0.0);
      if (loss_out) loss_out[0] = float(sum * inv_batch);
      *acc = 0.0;
      *ticket = 0u;
      __threadfence();
    }
  }
}

template <int LPR, int NCH, bool TRAIN>
__global__ void __launch_bounds__(kThreads)
bpr_vec(const float* __restrict__ Wu, const float* __restrict__ Wi, float* __restrict__ Gu,
        float* __restrict__ Gi, uint32_t* __restrict__ Tu, uint32_t* __restrict__ Ti, int d4,
        const int32_t* __restrict__ uid, const int32_t* __restrict__ pid, const int32_t* __restrict__ nid,
        int64_t batch, float inv_batch, double inv_local, double* loss_acc, unsigned int* ticket, float* __restrict__ out) {
  constexpr int GPW = 32 / LPR;  // groups (triplets) per warp
  const int lane_in = threadIdx.x & (LPR - 1);
  const int64_t group = (int64_t(blockIdx.x) * kThreads + threadIdx.x) / LPR;
  const int64_t n_groups = int64_t(gridDim.x) * kThreads / LPR;
  const int64_t warp_first = group - ((threadIdx.x & 31) / LPR);  // first triplet of this warp
  const float4* __restrict__ Wu4 = reinterpret_cast<const float4*>(Wu);
  const float4* __restrict__ Wi4 = reinterpret_cast<const float4*>(Wi);
  float loss_local = 0.f;

  for (int64_t wb = warp_first; wb < batch; wb += n_groups) {
    const int64_t b = wb + (threadIdx.x & 31) / LPR;
    const bool valid = b < batch;
    int64_t ru = 0, rp = 0, rn = 0;
    if (valid) { ru = __ldg(uid + b); rp = __ldg(pid + b); rn = __ldg(nid + b); }
    uint32_t su = 0xffffffffu, sp = 0xffffffffu, sn = 0xffffffffu;
    if (TRAIN && valid && lane_in == 0) { su = peek_touched(Tu, ru); sp = peek_touched(Ti, rp); sn = peek_touched(Ti, rn); }
    float4 u[NCH], p[NCH], n[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c = lane_in + k * LPR;
      if (valid && c < d4) {
        u[k] = __ldg(Wu4 + ru * d4 + c);
        p[k] = __ldg(Wi4 + rp * d4 + c);
        n[k] = __ldg(Wi4 + rn * d4 + c);
      } else {
        u[k] = p[k] = n[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      // difference first, then the product: mirrors x = <u,p> - <u,n> up to fp32 reassociation
      dot = fmaf(u[k].x, p[k].x - n[k].x, dot);
      dot = fmaf(u[k].y, p[k].y - n[k].y, dot);
      dot = fmaf(u[k].z, p[k].z - n[k].z, dot);
      dot = fmaf(u[k].w, p[k].w - n[k].w, dot);
    }
    const float x = group_sum<LPR>(dot);
    if constexpr (!TRAIN) {
      if (valid && lane_in == 0) out[b] = x;
    } else {
      const float s = sigmoidf_acc(x);
      if (valid && lane_in == 0) loss_local += 1.0f - s;
      const float g = -s * (1.0f - s) * inv_batch;
      if (valid) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          const int c = lane_in + k * LPR;
          if (c < d4) {
            const float4 du = make_float4(g * (p[k].x - n[k].x), g * (p[k].y - n[k].y),
                                          g * (p[k].z - n[k].z), g * (p[k].w - n[k].w));
            const float4 dp = make_float4(g * u[k].x, g * u[k].y, g * u[k].z, g * u[k].w);
            const float4 dn = make_float4(-dp.x, -dp.y, -dp.z, -dp.w);
            red_add_f4(Gu + (ru * d4 + c) * 4, du);
            red_add_f4(Gi + (rp * d4 + c) * 4, dp);
            red_add_f4(Gi + (rn * d4 + c) * 4, dn);
          }
        }
        if (lane_in == 0) { mark_touched(Tu, ru, su); mark_touched(Ti, rp, sp); mark_touched(Ti, rn, sn); }
      }
    }
  }
  (void)GPW;
  if constexpr (TRAIN) finish_loss(double(loss_local), inv_local, loss_acc, ticket, out);
}

// Generic width (d % 4 != 0 or unaligned): one warp per triplet, rows re-read for the backward.
template <bool TRAIN>
__global__ void __launch_bounds__(kThreads)
bpr_scalar(const float* __restrict__ Wu, const float* __restrict__ Wi, float* __restrict__ Gu,
           float* __restrict__ Gi, uint32_t* __restrict__ Tu, uint32_t* __restrict__ Ti, int d,
           const int32_t* __restrict__ uid, const int32_t* __restrict__ pid, const int32_t* __restrict__ nid,
           int64_t batch, float inv_batch, double inv_local, double* loss_acc, unsigned int* ticket, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * kThreads + threadIdx.x) >> 5;
  const int64_t n_warps = int64_t(gridDim.x) * kThreads >> 5;
  float loss_local = 0.f;
  for (int64_t b = warp; b < batch; b += n_warps) {
    const int64_t ru = __ldg(uid + b), rp = __ldg(pid + b), rn = __ldg(nid + b);
    uint32_t su = 0xffffffffu, sp = 0xffffffffu, sn = 0xffffffffu;
    if (TRAIN && lane == 0) { su = peek_touched(Tu, ru); sp = peek_touched(Ti, rp); sn = peek_touched(Ti, rn); }
    float dot = 0.f;
    for (int c = lane; c < d; c += 32)
      dot = fmaf(__ldg(Wu + ru * d + c), __ldg(Wi + rp * d + c) - __ldg(Wi + rn * d + c), dot);
    const float x = warp_sum(dot);
    if constexpr (!TRAIN) {
      if (lane == 0) out[b] = x;
    } else {
      const float s = sigmoidf_acc(x);
      if (lane == 0) loss_local += 1.0f - s;
      const float g = -s * (1.0f - s) * inv_batch;
      for (int c = lane; c < d; c += 32) {
        const float uu = __ldg(Wu + ru * d + c), pp = __ldg(Wi + rp * d + c), nn = __ldg(Wi + rn * d + c);
        atomicAdd(Gu + ru * d + c, g * (pp - nn));
        atomicAdd(Gi + rp * d + c, g * uu);
        atomicAdd(Gi + rn * d + c, -g * uu);
      }
      if (lane == 0) { mark_touched(Tu, ru, su); mark_touched(Ti, rp, sp); mark_touched(Ti, rn, sn); }
    }
  }
  if constexpr (TRAIN) finish_loss(double(loss_local), inv_local, loss_acc, ticket, out);
}

// ---- whole training steps in ONE cooperative kernel ------------------------------------------------
// K steps x (fused fwd/bwd -> grid.sync -> exact Keras Adam over both tables -> grid.sync): no kernel
// launch and no host involvement between steps; both phases are single-wave and latency-bound, so the
// launch gaps and cold starts between separate kernels were ~40 % of the step.
struct BprCoopParams {
  brk_table user, item;
  const int32_t* u; const int32_t* p; const int32_t* n;
  const int64_t* batch_index;     // device [n_steps] (long calls) ...
  int64_t inline_index[16];       // ... or by value (n_steps <= 16: no H2D copy at all)
  int32_t use_inline;
  int64_t total, batch;
  int32_t n_steps;
  brk_adam_hyper h;
  int64_t* state;                 // [t, beta1^t, beta2^t]
  float* losses;                  // device [n_steps] or null
  double* loss_acc;               // 2 doubles (ping-pong), zero on entry
};

__device__ __forceinline__ void adam4(float4& w, float4& m, float4& v, const float4 g, float alpha, float b1, float b2,
                                      float eps) {
  const float omb1 = 1.f - b1, omb2 = 1.f - b2;
  m.x = b1 * m.x + omb1 * g.x; v.x = b2 * v.x + omb2 * g.x * g.x; w.x -= alpha * m.x / (sqrtf(v.x) + eps);
  m.y = b1 * m.y + omb1 * g.y; v.y = b2 * v.y + omb2 * g.y * g.y; w.y -= alpha * m.y / (sqrtf(v.y) + eps);
  m.z = b1 * m.z + omb1 * g.z; v.z = b2 * v.z + omb2 * g.z * g.z; w.z -= alpha * m.z / (sqrtf(v.z) + eps);
  m.w = b1 * m.w + omb1 * g.w; v.w = b2 * v.w + omb2 * g.w * g.w; w.w -= alpha * m.w / (sqrtf(v.w) + eps);
}

template <int LPR, int NCH>
__global__ void __launch_bounds__(kThreads) bpr_steps_coop(const BprCoopParams P) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[32];
  const int d4 = P.user.d >> 2;
  const int lane_in = threadIdx.x & (LPR - 1);
  const int sub = (threadIdx.x & 31) / LPR;
  const int64_t group = (int64_t(blockIdx.x) * kThreads + threadIdx.x) / LPR;
  const int64_t n_groups = int64_t(gridDim.x) * kThreads / LPR;
  const int64_t warp_first = group - sub;
  const int64_t tid = int64_t(blockIdx.x) * kThreads + threadIdx.x, nthr = int64_t(gridDim.x) * kThreads;
  // no __restrict__/read-only path here: phase 2 of this same kernel rewrites the tables, so rows are
  // read with ld.global.cg (L2 only) -- another SM's L1 could otherwise hold a stale line across steps
  const float4* Wu4 = reinterpret_cast<const float4*>(P.user.w);
  const float4* Wi4 = reinterpret_cast<const float4*>(P.item.w);
  const double* pw = reinterpret_cast<const double*>(P.state);
  double p1 = pw[1], p2 = pw[2];                       // running beta powers, advanced per step in registers
  const brk_table tabs[2] = {P.user, P.item};

  for (int s = 0; s < P.n_steps; ++s) {
    const int64_t off = (P.use_inline ? P.inline_index[s] : P.batch_index[s]) * P.batch;
    const int64_t batch = (off + P.batch <= P.total) ? P.batch : P.total - off;
    const float inv_batch = 1.0f / float(batch);
    const int32_t* uid = P.u + off; const int32_t* pid = P.p + off; const int32_t* nid = P.n + off;
    // ---- phase 1: fused gather + loss + gradient REDs (same math as bpr_vec) ----
    float loss_local = 0.f;
    for (int64_t wb = warp_first; wb < batch; wb += n_groups) {
      const int64_t b = wb + sub;
      const bool valid = b < batch;
      int64_t ru = 0, rp = 0, rn = 0;
      if (valid) { ru = uid[b]; rp = pid[b]; rn = nid[b]; }
      float4 u[NCH], p[NCH], n[NCH];
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int c = lane_in + k * LPR;
        if (valid && c < d4) {           // plain loads: the tables are rewritten by phase 2 of the same kernel
          u[k] = __ldcg(Wu4 + ru * d4 + c); p[k] = __ldcg(Wi4 + rp * d4 + c); n[k] = __ldcg(Wi4 + rn * d4 + c);
        } else {
          u[k] = p[k] = n[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        dot = fmaf(u[k].x, p[k].x - n[k].x, dot); dot = fmaf(u[k].y, p[k].y - n[k].y, dot);
        dot = fmaf(u[k].z, p[k].z - n[k].z, dot); dot = fmaf(u[k].w, p[k].w - n[k].w, dot);
      }
      const float x = group_sum<LPR>(dot);
      const float sg = sigmoidf_acc(x);
      if (valid && lane_in == 0) loss_local += 1.0f - sg;
      const float g = -sg * (1.0f - sg) * inv_batch;
      if (valid) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          const int c = lane_in + k * LPR;
          if (c < d4) {
            const float4 dp = make_float4(g * u[k].x, g * u[k].y, g * u[k].z, g * u[k].w);
            red_add_f4(P.user.g + (ru * d4 + c) * 4, make_float4(g * (p[k].x - n[k].x), g * (p[k].y - n[k].y),
                                                                 g * (p[k].z - n[k].z), g * (p[k].w - n[k].w)));
            red_add_f4(P.item.g + (rp * d4 + c) * 4, dp);
            red_add_f4(P.item.g + (rn * d4 + c) * 4, make_float4(-dp.x, -dp.y, -dp.z, -dp.w));
          }
        }
      }
    }
    const double part = block_sum_double(double(loss_local), red);
    if (threadIdx.x == 0) atomicAdd(P.loss_acc + (s & 1), part);
    grid.sync();
    // ---- phase 2: exact Keras Adam over every element of both tables; zero the accumulators ----
    p1 *= double(P.h.beta1); p2 *= double(P.h.beta2);
    const float alpha = float(double(P.h.lr) * sqrt(1.0 - p2) / (1.0 - p1));
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int64_t n4 = tabs[k].rows * tabs[k].d / 4;
      float4* w4 = reinterpret_cast<float4*>(tabs[k].w); float4* g4 = reinterpret_cast<float4*>(tabs[k].g);
      float4* m4 = reinterpret_cast<float4*>(tabs[k].m); float4* v4 = reinterpret_cast<float4*>(tabs[k].v);
      for (int64_t i = tid; i < n4; i += nthr) {
        float4 w = w4[i], m = m4[i], v = v4[i];
        adam4(w, m, v, g4[i], alpha, P.h.beta1, P.h.beta2, P.h.eps);
        w4[i] = w; m4[i] = m; v4[i] = v; g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (tid == 0) {
      if (P.losses) P.losses[s] = float(P.loss_acc[s & 1] / double(batch));
      P.loss_acc[s & 1] = 0.0;
    }
    grid.sync();
  }
  if (tid == 0) {
    double* pwo = reinterpret_cast<double*>(P.state);
    P.state[0] += P.n_steps; pwo[1] = p1; pwo[2] = p2;
  }
}

template <bool TRAIN>
int launch_bpr(brk_ctx* ctx, const float* Wu, const float* Wi, float* Gu, float* Gi, uint32_t* Tu, uint32_t* Ti,
               int d, const int32_t* u, const int32_t* p, const int32_t* n, int64_t batch, int64_t global_batch,
               float* out, cudaStream_t st) {
  // gradients are scaled by 1/global_batch (data-parallel replicas sum their accumulators); the
  // reported loss is the mean over the local batch
  const float inv_batch = 1.0f / float(global_batch > 0 ? global_batch : batch);
  const double inv_local = 1.0 / double(batch);
  double* acc = ctx->loss_acc + 0;
  unsigned int* ticket = ctx->tickets + 0;
  const int64_t cap = int64_t(ctx->sm_count) * (2048 / kThreads);
  const bool vec = (d & 3) == 0 && d <= 512 && brk_aligned16(Wu) && brk_aligned16(Wi) &&
                   (!TRAIN || (brk_aligned16(Gu) && brk_aligned16(Gi)));
  if (vec) {
    const int d4 = d >> 2;
    const int lpr = brk_lanes_per_row(d4);
    const int nch = (d4 + lpr - 1) / lpr;
    int64_t need = (batch * lpr + kThreads - 1) / kThreads;
    const int grid = int(need < 1 ? 1 : (need < cap ? need : cap));
#define BRK_BPR_CASE(L, N)                                                                          \
  bpr_vec<L, N, TRAIN><<<grid, kThreads, 0, st>>>(Wu, Wi, Gu, Gi, Tu, Ti, d4, u, p, n, batch, inv_batch, \
                                                  inv_local, acc, ticket, out)
    if (lpr == 1) BRK_BPR_CASE(1, 1);
    else if (lpr == 2) BRK_BPR_CASE(2, 1);
    else if (lpr == 4) BRK_BPR_CASE(4, 1);
    else if (lpr == 8) BRK_BPR_CASE(8, 1);
    else if (lpr == 16) BRK_BPR_CASE(16, 1);
    else if (nch == 1) BRK_BPR_CASE(32, 1);
    else if (nch == 2) BRK_BPR_CASE(32, 2);
    else if (nch == 3) BRK_BPR_CASE(32, 3);
    else BRK_BPR_CASE(32, 4);
#undef BRK_BPR_CASE
  } else {
    int64_t need = (batch * 32 + kThreads - 1) / kThreads;
    const int grid = int(need < 1 ? 1 : (need < cap ? need : cap));
    bpr_scalar<TRAIN><<<grid, kThreads, 0, st>>>(Wu, Wi, Gu, Gi, Tu, Ti, d, u, p, n, batch, inv_batch, inv_local,
                                                 acc, ticket, out);
  }
  BRK_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int brk_bpr_fwd_bwd(brk_ctx* ctx, const brk_table* user, const brk_table* item,
                               const int32_t* u, const int32_t* p, const int32_t* n, int64_t batch,
                               int64_t global_batch, float* loss_out, void* stream) {
  BRK_REQUIRE(ctx && user && item && u && p && n, BRK_E_ARG, "brk_bpr_fwd_bwd: null argument");
  BRK_REQUIRE(user->w && user->g && item->w && item->g, BRK_E_ARG, "brk_bpr_fwd_bwd: table w/g missing");
  BRK_REQUIRE(user->d == item->d && user->d > 0, BRK_E_ARG, "brk_bpr_fwd_bwd: user d=%d item d=%d", user->d,
              item->d);
  BRK_REQUIRE(batch > 0, BRK_E_ARG, "brk_bpr_fwd_bwd: batch=%lld", (long long)batch);
  return launch_bpr<true>(ctx, user->w, item->w, user->g, item->g, user->touched, item->touched, user->d, u, p,
                          n, batch, global_batch, loss_out, (cudaStream_t)stream);
}

extern "C" int brk_bpr_scores(brk_ctx* ctx, const float* user_w, const float* item_w, int32_t d,
                              const int32_t* u, const int32_t* p, const int32_t* n, int64_t batch,
                              float* x_out, void* stream) {
  BRK_REQUIRE(ctx && user_w && item_w && u && p && n && x_out, BRK_E_ARG, "brk_bpr_scores: null argument");
  BRK_REQUIRE(d > 0 && batch > 0, BRK_E_ARG, "brk_bpr_scores: d=%d batch=%lld", d, (long long)batch);
  return launch_bpr<false>(ctx, user_w, item_w, nullptr, nullptr, nullptr, nullptr, d, u, p, n, batch, 0, x_out,
                           (cudaStream_t)stream);
}

static bool coop_eligible(const brk_table* user, const brk_table* item, int lazy_adam) {
  const int d = user->d;
  return !lazy_adam && (d & 3) == 0 && d <= 128 && user->m && user->v && item->m && item->v &&
         brk_aligned16(user->w) && brk_aligned16(item->w) && brk_aligned16(user->g) && brk_aligned16(item->g) &&
         ((user->rows * d) & 3) == 0 && ((item->rows * d) & 3) == 0 && getenv("BRK_NO_COOP") == nullptr;
}

static int launch_coop_steps(brk_ctx* ctx, const brk_table* user, const brk_table* item, const int32_t* u,
                             const int32_t* p, const int32_t* n, int64_t total, int64_t batch,
                             const int64_t* batch_index_host, int32_t n_steps, brk_adam_hyper h, int64_t* step_dev,
                             float* losses, cudaStream_t st) {
  const int d = user->d;
    BprCoopParams P;
    P.use_inline = n_steps <= 16;
    if (P.use_inline) {
      for (int k = 0; k < n_steps; ++k) P.inline_index[k] = batch_index_host[k];
    } else {
      // batch indices to the device (the context keeps a small scratch buffer)
      const size_t need = size_t(n_steps) * sizeof(int64_t) + 64;
      if (ctx->scratch_bytes < need) {
        if (ctx->scratch) BRK_CUDA(cudaFree(ctx->scratch));
        BRK_CUDA(cudaMalloc(&ctx->scratch, need * 2));
        ctx->scratch_bytes = need * 2;
      }
      BRK_CUDA(cudaMemcpyAsync(ctx->scratch, batch_index_host, size_t(n_steps) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    }
    P.user = *user; P.item = *item; P.u = u; P.p = p; P.n = n;
    P.batch_index = reinterpret_cast<const int64_t*>(ctx->scratch);
    P.total = total; P.batch = batch; P.n_steps = n_steps; P.h = h; P.state = step_dev; P.losses = losses;
    P.loss_acc = ctx->loss_acc + 2;
    const int lpr = brk_lanes_per_row(d >> 2);
    void* fn = nullptr;
    switch (lpr) {
      case 1: fn = (void*)bpr_steps_coop<1, 1>; break;
      case 2: fn = (void*)bpr_steps_coop<2, 1>; break;
      case 4: fn = (void*)bpr_steps_coop<4, 1>; break;
      case 8: fn = (void*)bpr_steps_coop<8, 1>; break;
      case 16: fn = (void*)bpr_steps_coop<16, 1>; break;
      default: fn = (void*)bpr_steps_coop<32, 1>; break;
    }
    int per_sm = 0;
    BRK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, 0));
    BRK_REQUIRE(per_sm > 0, BRK_E_STATE, "brk_bpr_train_steps: cooperative kernel does not fit");
    int64_t want = (batch * lpr + kThreads - 1) / kThreads;
    const int64_t cap = int64_t(per_sm) * ctx->sm_count;
    const int grid = int(want < cap ? (want < 1 ? 1 : want) : cap);
    void* args[] = {(void*)&P};
    BRK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreads), args, 0, st));
    return 0;
}

// Multi-step driver: one C call enqueues n_steps x (fused fwd/bwd + optimizer) so that the host
// cost per step is two kernel launches and nothing else (model.fit's inner loop,
// /root/reference/src/models/BPRModel.py:109).  batch_index_host[k] selects which batch of the
// device-resident triplet arrays step k consumes (Keras shuffles batches per epoch).
extern "C" int brk_bpr_train_steps(brk_ctx* ctx, const brk_table* user, const brk_table* item,
                                   const int32_t* u, const int32_t* p, const int32_t* n, int64_t total,
                                   int64_t batch, const int64_t* batch_index_host, int32_t n_steps,
                                   brk_adam_hyper h, int32_t lazy_adam, int64_t* step_dev, float* losses,
                                   void* stream) {
  BRK_REQUIRE(ctx && user && item && u && p && n && batch_index_host && step_dev, BRK_E_ARG,
              "brk_bpr_train_steps: null argument");
  BRK_REQUIRE(total > 0 && batch > 0 && n_steps >= 0, BRK_E_ARG, "brk_bpr_train_steps: total=%lld batch=%lld",
              (long long)total, (long long)batch);
  const int64_t n_batches = (total + batch - 1) / batch;
  for (int k = 0; k < n_steps; ++k)
    BRK_REQUIRE(batch_index_host[k] >= 0 && batch_index_host[k] < n_batches, BRK_E_ARG,
                "brk_bpr_train_steps: batch index %lld of %lld", (long long)batch_index_host[k], (long long)n_batches);
  if (n_steps == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool coop_ok = coop_eligible(user, item, lazy_adam);
  if (coop_ok) return launch_coop_steps(ctx, user, item, u, p, n, total, batch, batch_index_host, n_steps, h, step_dev, losses, st);
  brk_table tabs[2] = {*user, *item};
  for (int k = 0; k < n_steps; ++k) {
    const int64_t bi = batch_index_host[k];
    const int64_t off = bi * batch;
    const int64_t cnt = (off + batch <= total) ? batch : total - off;
    int rc = brk_bpr_fwd_bwd(ctx, user, item, u + off, p + off, n + off, cnt, 0, losses ? losses + k : nullptr, stream);
    if (rc) return rc;
    rc = lazy_adam ? brk_adam_rows(ctx, tabs, 2, h, step_dev, 1, stream)
                   : brk_adam_dense_keras(ctx, tabs, 2, h, step_dev, 1, stream);
    if (rc) return rc;
  }
  return 0;
}

// End-to-end steps from HOST buffers in one call (the public-API path bench.py's `e2e` times):
// per step  H2D(user ids, positive ids)  ->  Philox negatives on device  ->  fused fwd/bwd  ->  Adam
// ->  D2H(loss).  u_host/p_host/losses_host must be page-locked for the copies to be asynchronous.
// d_stage: device int32 scratch of 6*batch (two slots of u,p,n); d_losses: device [n_steps].
extern "C" int brk_bpr_train_steps_host(brk_ctx* ctx, const brk_table* user, const brk_table* item,
                                        const int32_t* u_host, const int32_t* p_host, int64_t total, int64_t batch,
                                        const int64_t* batch_index_host, int32_t n_steps, uint32_t seed, uint32_t epoch,
                                        int32_t num_items, const int64_t* csr_indptr, const int32_t* csr_items,
                                        brk_adam_hyper h, int32_t lazy_adam, int64_t* step_dev, int32_t* d_stage,
                                        float* d_losses, float* losses_host, void* stream) {
  BRK_REQUIRE(ctx && user && item && u_host && p_host && batch_index_host && step_dev && d_stage && d_losses &&
                  csr_indptr && csr_items, BRK_E_ARG, "brk_bpr_train_steps_host: null argument");
  BRK_REQUIRE(total > 0 && batch > 0 && n_steps >= 0, BRK_E_ARG, "brk_bpr_train_steps_host: total=%lld batch=%lld",
              (long long)total, (long long)batch);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_batches = (total + batch - 1) / batch;
  for (int k = 0; k < n_steps; ++k)
    BRK_REQUIRE(batch_index_host[k] >= 0 && batch_index_host[k] < n_batches, BRK_E_ARG,
                "brk_bpr_train_steps_host: batch index %lld of %lld", (long long)batch_index_host[k], (long long)n_batches);
  if (n_steps == 0) return 0;
  if (!ctx->copy_ready) {
    BRK_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int q = 0; q < 2; ++q) {
      BRK_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready[q], cudaEventDisableTiming));
      BRK_CUDA(cudaEventCreateWithFlags(&ctx->ev_done[q], cudaEventDisableTiming));
    }
    ctx->copy_ready = 1;
  }
  cudaStream_t cs = ctx->copy_stream;
  const bool coop_ok = coop_eligible(user, item, lazy_adam);
  brk_table tabs[2] = {*user, *item};
  const int64_t zero_index = 0;
  // the copy stream must not run ahead of work already queued on `st` that still uses the staging slots
  BRK_CUDA(cudaEventRecord(ctx->ev_done[0], st));
  BRK_CUDA(cudaEventRecord(ctx->ev_done[1], st));
  auto stage = [&](int k) -> int {                 // H2D of step k's ids into slot k&1, on the copy stream
    const int64_t off = batch_index_host[k] * batch;
    const int64_t cnt = (off + batch <= total) ? batch : total - off;
    int32_t* du = d_stage + (k & 1) * 3 * batch;
    BRK_CUDA(cudaStreamWaitEvent(cs, ctx->ev_done[k & 1], 0));          // slot free (step k-2 finished)
    BRK_CUDA(cudaMemcpyAsync(du, u_host + off, cnt * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
    BRK_CUDA(cudaMemcpyAsync(du + batch, p_host + off, cnt * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
    BRK_CUDA(cudaEventRecord(ctx->ev_ready[k & 1], cs));
    return 0;
  };
  if (int rc = stage(0)) return rc;
  for (int k = 0; k < n_steps; ++k) {
    if (k + 1 < n_steps) { if (int rc = stage(k + 1)) return rc; }       // prefetch the next batch's ids
    const int64_t off = batch_index_host[k] * batch;
    const int64_t cnt = (off + batch <= total) ? batch : total - off;
    int32_t* du = d_stage + (k & 1) * 3 * batch;
    int32_t* dp = du + batch;
    int32_t* dn = dp + batch;
    BRK_CUDA(cudaStreamWaitEvent(st, ctx->ev_ready[k & 1], 0));
    int rc = brk_philox_bpr_negatives(ctx, du, cnt, off, seed, epoch, num_items, csr_indptr, csr_items, dn, stream);
    if (rc) return rc;
    if (coop_ok) {
      rc = launch_coop_steps(ctx, user, item, du, dp, dn, cnt, cnt, &zero_index, 1, h, step_dev, d_losses + k, st);
      if (rc) return rc;
    } else {
      rc = brk_bpr_fwd_bwd(ctx, user, item, du, dp, dn, cnt, 0, d_losses + k, stream);
      if (rc) return rc;
      rc = lazy_adam ? brk_adam_rows(ctx, tabs, 2, h, step_dev, 1, stream)
                     : brk_adam_dense_keras(ctx, tabs, 2, h, step_dev, 1, stream);
      if (rc) return rc;
    }
    BRK_CUDA(cudaEventRecord(ctx->ev_done[k & 1], st));
    if (losses_host) {                               // the step's result goes back on the copy stream
      BRK_CUDA(cudaStreamWaitEvent(cs, ctx->ev_done[k & 1], 0));
      BRK_CUDA(cudaMemcpyAsync(losses_host + k, d_losses + k, sizeof(float), cudaMemcpyDeviceToHost, cs));
    }
  }
  // rejoin: everything the copy stream did is ordered before whatever follows on `st`
  BRK_CUDA(cudaEventRecord(ctx->ev_ready[0], cs));
  BRK_CUDA(cudaStreamWaitEvent(st, ctx->ev_ready[0], 0));
  return 0;
}

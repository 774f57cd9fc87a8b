// NeuMF forward / backward for ANY layer widths (runtime E, EMF, H1, H2, H3): the catch-all behind the templated
// instances of neumf2.cu / neumf_tc.cu / neumf_fused.cu.  `numFactor` is a free attribute of the reference model
// (/root/reference/src/models/RModel.py:35; NeuMFModel.py:53-83 derives Dense(F), Dense(F // 2), Dense(F // 4) from
// it), so a drop-in must accept every value; the He et al. variant (Hadamard GMF vector, no BatchNorm) is covered for
// every width as well.  fp32 on the CUDA cores, one small kernel per graph node, every intermediate feature-major
// [features][batch] in a library-owned scratch (coalesced along the batch).  Nothing here is tuned: the hot
// configurations have their own kernels; this path is for correctness at unusual widths and is tested against the
// same oracle (tests/test_gpu_neumf.py).
//
// All reductions over the batch are one block per output with a fixed-order tree: results are bit-reproducible
// except for the embedding-gradient atomics.
#include "neumf_common.cuh"

namespace ngen {

using v2::drop16_bits; using v2::kBnEps; using v2::kBnMomentum; using v2::kDropScale;

constexpr int NT = 256;

__device__ __forceinline__ float actf(float x, int act) { return act == 0 ? fmaxf(x, 0.f) : 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float actg(float h, int act) { return act == 0 ? (h > 0.f ? 1.f : 0.f) : h * (1.f - h); }
__device__ __forceinline__ float keep_factor(uint64_t idx, int f, int layer, uint32_t seed, uint32_t epoch) {
  return ((drop16_bits(idx, f >> 4, layer, seed, epoch) >> (f & 15)) & 1u) ? kDropScale : 0.f;
}

// dst[(col0 + k) * B + s] = table[ids[s] * d + k]  (* layer-0 keep factor of feature col0 + k)
__global__ void gather_fm(const float* __restrict__ table, const int32_t* __restrict__ ids, int d, int64_t B, float* dst, int col0,
                          int dropout, int64_t first, uint32_t seed, uint32_t epoch) {
  const int64_t idx = int64_t(blockIdx.x) * NT + threadIdx.x;
  if (idx >= B * d) return;
  const int k = int(idx / B); const int64_t s = idx % B;
  float v = __ldg(table + int64_t(__ldg(ids + s)) * d + k);
  if (dropout) v *= keep_factor(uint64_t(first + s), col0 + k, 0, seed, epoch);
  dst[int64_t(col0 + k) * B + s] = v;
}
// out[j][s] = act(sum_k in[k][s] W[k][j] + b[j])      (W is the Keras kernel [K][N])
__global__ void dense_fwd(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ b, int K, int N,
                          int64_t B, float* out, int act) {
  const int64_t idx = int64_t(blockIdx.x) * NT + threadIdx.x;
  if (idx >= B * N) return;
  const int j = int(idx / B); const int64_t s = idx % B;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(in[int64_t(k) * B + s], __ldg(W + int64_t(k) * N + j), acc);
  acc += __ldg(b + j);
  out[idx] = act < 0 ? acc : actf(acc, act);
}
// out[k][s] = (sum_j dz[j][s] W[k][j]) * keep factor of feature k in `layer` (layer < 0: none)
__global__ void dense_bwd_in(const float* __restrict__ dz, const float* __restrict__ W, int K, int N, int64_t B, float* out, int layer,
                             int64_t first, uint32_t seed, uint32_t epoch) {
  const int64_t idx = int64_t(blockIdx.x) * NT + threadIdx.x;
  if (idx >= B * K) return;
  const int k = int(idx / B); const int64_t s = idx % B;
  float acc = 0.f;
  for (int j = 0; j < N; ++j) acc = fmaf(dz[int64_t(j) * B + s], __ldg(W + int64_t(k) * N + j), acc);
  if (layer >= 0) acc *= keep_factor(uint64_t(first + s), k, layer, seed, epoch);
  out[idx] = acc;
}
__device__ __forceinline__ double block_sum(double x, double* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = x;
  __syncthreads();
  x = threadIdx.x < NT / 32 ? sm[threadIdx.x] : 0.0;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  }
  __syncthreads();
  return x;                                                 // valid in thread 0
}
// one block per feature j: sa[j] = sum_s a[j][s];  sb[j] = sum_s a[j][s] * (b ? b[j][s] : a[j][s])
// mode 1: b is h and the second factor is xhat = (h - mean) * rstd  (BatchNorm-backward sums)
__global__ void col_sums(const float* __restrict__ a, const float* __restrict__ b, int64_t B, double* sa, double* sb, int mode,
                         const float* __restrict__ mean, const float* __restrict__ rstd) {
  __shared__ double sm[NT / 32];
  const int j = blockIdx.x;
  double x = 0.0, y = 0.0;
  for (int64_t s = threadIdx.x; s < B; s += NT) {
    const float av = a[int64_t(j) * B + s];
    float bv = b ? b[int64_t(j) * B + s] : av;
    if (mode == 1) bv = (bv - mean[j]) * rstd[j];
    x += double(av); y += double(av) * double(bv);
  }
  x = block_sum(x, sm); y = block_sum(y, sm);
  if (threadIdx.x == 0) { sa[j] = x; sb[j] = y; }
}
// mean / rstd / var of a BatchNorm layer from its batch sums (training) or its moving statistics
__global__ void bn_stats(const double* sum, const double* sq, const float* mov_mean, const float* mov_var, int H, int64_t B,
                         int training, float* mean, float* rstd, float* var) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= H) return;
  float mu, vv;
  if (training) {
    const double m = sum[f] / double(B);
    mu = float(m); vv = float(fmax(sq[f] / double(B) - m * m, 0.0));
  } else { mu = mov_mean[f]; vv = mov_var[f]; }
  mean[f] = mu; var[f] = vv; rstd[f] = 1.0f / sqrtf(vv + kBnEps);
}
// a[j][s] = dropout(gamma[j] * (h - mean[j]) * rstd[j] + beta[j]);  bn == 0: a = dropout(h)
__global__ void bn_apply(const float* __restrict__ h, const float* mean, const float* rstd, const float* gamma, const float* beta,
                         int H, int64_t B, float* a, int bn, int dropout, int layer, int64_t first, uint32_t seed, uint32_t epoch) {
  const int64_t idx = int64_t(blockIdx.x) * NT + threadIdx.x;
  if (idx >= B * H) return;
  const int j = int(idx / B); const int64_t s = idx % B;
  float y = h[idx];
  if (bn) y = gamma[j] * ((y - mean[j]) * rstd[j]) + beta[j];
  if (dropout) y *= keep_factor(uint64_t(first + s), j, layer, seed, epoch);
  a[idx] = y;
}
// dz[j][s] = (bn ? gamma rstd (dy - sum_dy / B - xhat sum_dyx / B) : dy) * act'(h)
__global__ void bn_bwd(const float* __restrict__ dy, const float* __restrict__ h, const float* mean, const float* rstd,
                       const float* gamma, const double* sd, const double* se, int H, int64_t B, float* dz, int bn, int act) {
  const int64_t idx = int64_t(blockIdx.x) * NT + threadIdx.x;
  if (idx >= B * H) return;
  const int j = int(idx / B);
  const float hv = h[idx];
  float d = dy[idx];
  if (bn) {
    const float xh = (hv - mean[j]) * rstd[j];
    d = gamma[j] * rstd[j] * (d - float(sd[j] / double(B)) - xh * float(se[j] / double(B)));
  }
  dz[idx] = d * actg(hv, act);
}
// one block per (k, j): dW[k][j] += sum_s in[k][s] dz[j][s];  blocks with k == K: db[j] += sum_s dz[j][s]
__global__ void wgrad(const float* __restrict__ in, const float* __restrict__ dz, int K, int N, int64_t B, float* dW, float* db) {
  __shared__ double sm[NT / 32];
  const int k = blockIdx.x / N, j = blockIdx.x % N;
  double x = 0.0;
  for (int64_t s = threadIdx.x; s < B; s += NT)
    x += double(k < K ? in[int64_t(k) * B + s] : 1.f) * double(dz[int64_t(j) * B + s]);
  x = block_sum(x, sm);
  if (threadIdx.x == 0) { if (k < K) dW[int64_t(k) * N + j] += float(x); else db[j] += float(x); }
}
// MF part of the head: mode 0 mf[0][s] = <uMF[u], iMF[i]>;  mode 1 mf[f][s] = uMF[u][f] * iMF[i][f]
__global__ void mf_forward(const float* __restrict__ um, const float* __restrict__ im, const int32_t* __restrict__ u,
                           const int32_t* __restrict__ i, int EMF, int64_t B, float* mf, int mode) {
  const int64_t s = int64_t(blockIdx.x) * NT + threadIdx.x;
  if (s >= B) return;
  const float* a = um + int64_t(__ldg(u + s)) * EMF; const float* b = im + int64_t(__ldg(i + s)) * EMF;
  float dot = 0.f;
  for (int f = 0; f < EMF; ++f) {
    const float p = __ldg(a + f) * __ldg(b + f);
    if (mode) mf[int64_t(f) * B + s] = p; else dot = fmaf(__ldg(a + f), __ldg(b + f), dot);
  }
  if (!mode) mf[s] = dot;
}
// logit, prediction, loss, d loss / d logit, dz3
__global__ void head(const float* __restrict__ h3, const float* __restrict__ mf, const float* __restrict__ w4, int H3, int HM,
                     const float* __restrict__ y, int64_t B, int64_t global_B, int loss_kind, int act, int training, float* out,
                     float* dl, float* dz3, double* loss_acc) {
  __shared__ double sm[NT / 32];
  const int64_t s = int64_t(blockIdx.x) * NT + threadIdx.x;
  float lloc = 0.f;
  if (s < B) {
    float logit = __ldg(w4 + H3 + HM);
    for (int j = 0; j < H3; ++j) logit = fmaf(h3[int64_t(j) * B + s], __ldg(w4 + j), logit);
    for (int f = 0; f < HM; ++f) logit = fmaf(mf[int64_t(f) * B + s], __ldg(w4 + H3 + f), logit);
    const float o = 1.0f / (1.0f + expf(-logit));
    out[s] = o;
    if (y) {
      const float yv = __ldg(y + s), invB = 1.0f / float(global_B);
      float dlogit;
      if (loss_kind == 0) { const float e = o - yv; lloc = e * e; dlogit = 2.f * e * o * (1.f - o) * invB; }
      else { lloc = fmaxf(logit, 0.f) - logit * yv + log1pf(expf(-fabsf(logit))); dlogit = (o - yv) * invB; }
      if (training) {
        dl[s] = dlogit;
        for (int j = 0; j < H3; ++j) dz3[int64_t(j) * B + s] = dlogit * __ldg(w4 + j) * actg(h3[int64_t(j) * B + s], act);
      }
    }
  }
  const double t = block_sum(double(lloc), sm);
  if (threadIdx.x == 0 && y) atomicAdd(loss_acc, t);
}
// head weight gradients: one block per input z of the head (h3 rows, MF rows, the constant 1 of b4)
__global__ void head_wgrad(const float* __restrict__ h3, const float* __restrict__ mf, const float* __restrict__ dl, int H3, int HM,
                           int64_t B, float* dW4) {
  __shared__ double sm[NT / 32];
  const int r = blockIdx.x;
  const float* z = r < H3 ? h3 + int64_t(r) * B : (r < H3 + HM ? mf + int64_t(r - H3) * B : nullptr);
  double x = 0.0;
  for (int64_t s = threadIdx.x; s < B; s += NT) x += double(z ? z[s] : 1.f) * double(dl[s]);
  x = block_sum(x, sm);
  if (threadIdx.x == 0) dW4[r] += float(x);
}
// MF embedding gradients
__global__ void mf_backward(const float* __restrict__ um, const float* __restrict__ im, float* gu, float* gi, uint32_t* tu, uint32_t* ti,
                            const int32_t* __restrict__ u, const int32_t* __restrict__ i, const float* __restrict__ dl,
                            const float* __restrict__ w4mf, int EMF, int64_t B, int mode) {
  const int64_t idx = int64_t(blockIdx.x) * NT + threadIdx.x;
  if (idx >= B * EMF) return;
  const int64_t s = idx / EMF; const int f = int(idx % EMF);
  const int64_t ur = __ldg(u + s), ir = __ldg(i + s);
  const float c = dl[s] * __ldg(w4mf + (mode ? f : 0));
  atomicAdd(gu + ur * EMF + f, c * __ldg(im + ir * EMF + f));
  atomicAdd(gi + ir * EMF + f, c * __ldg(um + ur * EMF + f));
  if (f == 0) {
    if (tu) atomicOr(tu + (ur >> 5), 1u << (ur & 31));
    if (ti) atomicOr(ti + (ir >> 5), 1u << (ir & 31));
  }
}
// g[ids[s]][k] += dx[col0 + k][s]
__global__ void scatter_fm(const float* __restrict__ dx, const int32_t* __restrict__ ids, int d, int64_t B, float* g, uint32_t* touched,
                           int col0) {
  const int64_t idx = int64_t(blockIdx.x) * NT + threadIdx.x;
  if (idx >= B * d) return;
  const int64_t s = idx / d; const int k = int(idx % d);
  const int64_t r = __ldg(ids + s);
  atomicAdd(g + r * d + k, dx[int64_t(col0 + k) * B + s]);
  if (k == 0 && touched) atomicOr(touched + (r >> 5), 1u << (r & 31));
}
__global__ void add_bn_param_grads(const double* sd, const double* se, int H, float* gbeta, float* ggamma) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < H) { gbeta[f] += float(sd[f]); ggamma[f] += float(se[f]); }
}
__global__ void finish(double* acc, int n_acc, int loss_off, int64_t B, float* loss_out, float* bn_moving, const float* mean1,
                       const float* var1, const float* mean2, const float* var2, int H1, int H2, int update_moving) {
  const int t = threadIdx.x;
  if (update_moving) {
    for (int f = t; f < H1; f += blockDim.x) {
      bn_moving[f] = bn_moving[f] * kBnMomentum + mean1[f] * (1.f - kBnMomentum);
      bn_moving[H1 + f] = bn_moving[H1 + f] * kBnMomentum + var1[f] * (1.f - kBnMomentum);
    }
    for (int f = t; f < H2; f += blockDim.x) {
      bn_moving[2 * H1 + f] = bn_moving[2 * H1 + f] * kBnMomentum + mean2[f] * (1.f - kBnMomentum);
      bn_moving[2 * H1 + H2 + f] = bn_moving[2 * H1 + H2 + f] * kBnMomentum + var2[f] * (1.f - kBnMomentum);
    }
  }
  if (t == 0 && loss_out) loss_out[0] = float(acc[loss_off] / double(B));
  __syncthreads();
  for (int j = t; j < n_acc; j += blockDim.x) acc[j] = 0.0;
}

static inline unsigned blocks(int64_t n) { return unsigned((n + NT - 1) / NT); }

}  // namespace ngen

// Runs the step for any widths.  `m` carries plain (unsharded) tables.  Scratch is owned by the context and grown on
// demand: (4 E + 4 H1 + 4 H2 + 2 H3 + HM + 1) * batch + 3 (H1 + H2) floats.
int brk_neumf_step_generic(brk_ctx* ctx, const brk_neumf_model* m, const int32_t* u, const int32_t* i, const float* y,
                           int64_t B, int64_t global_batch, int64_t first_index, int32_t training, uint32_t seed, uint32_t epoch,
                           const brk_neumf_workspace* ws, float* out, float* loss_out, cudaStream_t st) {
  using namespace ngen;
  const int E = m->E, EMF = m->EMF > 0 ? m->EMF : m->E, H1 = m->H1, H2 = m->H2, H3 = m->H3;
  const int had = m->mf_mode != 0, bn = m->no_batch_norm ? 0 : 1, HM = had ? EMF : 1, K0 = 2 * E, act = m->act;
  BRK_REQUIRE(E > 0 && EMF > 0 && H1 > 0 && H2 > 0 && H3 > 0, BRK_E_ARG, "brk_neumf_step: widths E=%d EMF=%d H=(%d,%d,%d)", E, EMF, H1, H2, H3);
  const int dropout = (m->dropout != 0 && training) ? 1 : 0;
  const int64_t gB = global_batch > 0 ? global_batch : B;
  // dense block offsets (include/brk_b200.h)
  const int oW1 = 0, ob1 = oW1 + K0 * H1, og1 = ob1 + H1, obe1 = og1 + H1, oW2 = obe1 + H1, ob2 = oW2 + H1 * H2, og2 = ob2 + H2,
            obe2 = og2 + H2, oW3 = obe2 + H2, ob3 = oW3 + H2 * H3, oW4 = ob3 + H3;
  const float* W = m->dense.w; float* G = m->dense.g;
  // accumulator block (doubles): s1 q1 s2 q2 d2 e2 d1 e1 loss  (brk_neumf_acc_doubles)
  double* acc = ws->acc;
  double *s1 = acc, *q1 = s1 + H1, *s2 = q1 + H1, *q2 = s2 + H2, *d2 = q2 + H2, *e2 = d2 + H2, *d1 = e2 + H2, *e1 = d1 + H1;
  const int loss_off = 4 * H1 + 4 * H2, n_acc = loss_off + 1;
  const size_t need = size_t(4 * E + 4 * H1 + 4 * H2 + 2 * H3 + HM + 1) * size_t(B) + 3 * size_t(H1 + H2) + 64;
  if (ctx->neumf_gen_floats < need) {
    if (ctx->neumf_gen) BRK_CUDA(cudaFree(ctx->neumf_gen));
    ctx->neumf_gen = nullptr; ctx->neumf_gen_floats = 0;
    BRK_CUDA(cudaMalloc(&ctx->neumf_gen, need * sizeof(float)));
    ctx->neumf_gen_floats = need;
  }
  float* p = ctx->neumf_gen;
  auto take = [&](size_t n) { float* r = p; p += n; return r; };
  float *x0 = take(size_t(K0) * B), *h1 = take(size_t(H1) * B), *a1 = take(size_t(H1) * B), *h2 = take(size_t(H2) * B),
        *a2 = take(size_t(H2) * B), *h3 = take(size_t(H3) * B), *mf = take(size_t(HM) * B), *dl = take(size_t(B)),
        *dz3 = take(size_t(H3) * B), *dy2 = take(size_t(H2) * B), *dy1 = take(size_t(H1) * B), *dx0 = take(size_t(K0) * B);
  float *mean1 = take(H1), *rstd1 = take(H1), *var1 = take(H1), *mean2 = take(H2), *rstd2 = take(H2), *var2 = take(H2);
  float* dz2 = a2;   // a2 is dead once dW3 has read it; dz1 likewise reuses a1 after dW2
  float* dz1 = a1;

  gather_fm<<<blocks(B * E), NT, 0, st>>>(m->uMLP.w, u, E, B, x0, 0, dropout, first_index, seed, epoch);
  gather_fm<<<blocks(B * E), NT, 0, st>>>(m->iMLP.w, i, E, B, x0, E, dropout, first_index, seed, epoch);
  dense_fwd<<<blocks(B * H1), NT, 0, st>>>(x0, W + oW1, W + ob1, K0, H1, B, h1, act);
  if (bn && training) col_sums<<<H1, NT, 0, st>>>(h1, nullptr, B, s1, q1, 0, nullptr, nullptr);
  if (bn) bn_stats<<<(H1 + 127) / 128, 128, 0, st>>>(s1, q1, m->bn_moving, m->bn_moving + H1, H1, B, training, mean1, rstd1, var1);
  bn_apply<<<blocks(B * H1), NT, 0, st>>>(h1, mean1, rstd1, W + og1, W + obe1, H1, B, a1, bn, dropout, 1, first_index, seed, epoch);
  dense_fwd<<<blocks(B * H2), NT, 0, st>>>(a1, W + oW2, W + ob2, H1, H2, B, h2, act);
  if (bn && training) col_sums<<<H2, NT, 0, st>>>(h2, nullptr, B, s2, q2, 0, nullptr, nullptr);
  if (bn) bn_stats<<<(H2 + 127) / 128, 128, 0, st>>>(s2, q2, m->bn_moving + 2 * H1, m->bn_moving + 2 * H1 + H2, H2, B, training, mean2, rstd2, var2);
  bn_apply<<<blocks(B * H2), NT, 0, st>>>(h2, mean2, rstd2, W + og2, W + obe2, H2, B, a2, bn, dropout, 2, first_index, seed, epoch);
  dense_fwd<<<blocks(B * H3), NT, 0, st>>>(a2, W + oW3, W + ob3, H2, H3, B, h3, act);
  mf_forward<<<blocks(B), NT, 0, st>>>(m->uMF.w, m->iMF.w, u, i, EMF, B, mf, had);
  head<<<blocks(B), NT, 0, st>>>(h3, mf, W + oW4, H3, HM, y, B, gB, m->loss, act, training, out, dl, dz3, acc + loss_off);
  if (training) {
    head_wgrad<<<H3 + HM + 1, NT, 0, st>>>(h3, mf, dl, H3, HM, B, G + oW4);
    mf_backward<<<blocks(B * EMF), NT, 0, st>>>(m->uMF.w, m->iMF.w, m->uMF.g, m->iMF.g, m->uMF.touched, m->iMF.touched, u, i, dl,
                                                W + oW4 + H3, EMF, B, had);
    wgrad<<<(H2 + 1) * H3, NT, 0, st>>>(a2, dz3, H2, H3, B, G + oW3, G + ob3);
    dense_bwd_in<<<blocks(B * H2), NT, 0, st>>>(dz3, W + oW3, H2, H3, B, dy2, dropout ? 2 : -1, first_index, seed, epoch);
    if (bn) {
      col_sums<<<H2, NT, 0, st>>>(dy2, h2, B, d2, e2, 1, mean2, rstd2);
      add_bn_param_grads<<<(H2 + 127) / 128, 128, 0, st>>>(d2, e2, H2, G + obe2, G + og2);
    }
    bn_bwd<<<blocks(B * H2), NT, 0, st>>>(dy2, h2, mean2, rstd2, W + og2, d2, e2, H2, B, dz2, bn, act);
    wgrad<<<(H1 + 1) * H2, NT, 0, st>>>(a1, dz2, H1, H2, B, G + oW2, G + ob2);
    dense_bwd_in<<<blocks(B * H1), NT, 0, st>>>(dz2, W + oW2, H1, H2, B, dy1, dropout ? 1 : -1, first_index, seed, epoch);
    if (bn) {
      col_sums<<<H1, NT, 0, st>>>(dy1, h1, B, d1, e1, 1, mean1, rstd1);
      add_bn_param_grads<<<(H1 + 127) / 128, 128, 0, st>>>(d1, e1, H1, G + obe1, G + og1);
    }
    bn_bwd<<<blocks(B * H1), NT, 0, st>>>(dy1, h1, mean1, rstd1, W + og1, d1, e1, H1, B, dz1, bn, act);
    wgrad<<<(K0 + 1) * H1, NT, 0, st>>>(x0, dz1, K0, H1, B, G + oW1, G + ob1);
    dense_bwd_in<<<blocks(B * K0), NT, 0, st>>>(dz1, W + oW1, K0, H1, B, dx0, dropout ? 0 : -1, first_index, seed, epoch);
    scatter_fm<<<blocks(B * E), NT, 0, st>>>(dx0, u, E, B, m->uMLP.g, m->uMLP.touched, 0);
    scatter_fm<<<blocks(B * E), NT, 0, st>>>(dx0, i, E, B, m->iMLP.g, m->iMLP.touched, E);
  }
  finish<<<1, 256, 0, st>>>(acc, n_acc, loss_off, B, y ? loss_out : nullptr, m->bn_moving, mean1, var1, mean2, var2, H1, H2,
                            (training && bn) ? 1 : 0);
  BRK_LAUNCH_CHECK();
  return 0;
}

// NeuMF on the 5th-generation tensor cores (tcgen05, TF32 operands, fp32 accumulation in TMEM).
//
// Every activation / gradient tile of 128 samples is staged in shared memory sample-major (rows of 32 floats
// = 128 B, one region of ROWS*128 B per block of 32 columns), in one of the two swizzles the tensor core reads
// fp32 (TF32) operands through:
//   "KM" K-major, SWIZZLE_128B        : 8-row groups of 1024 B, 16-byte chunk index ^ (row & 7)
//   "MN" MN-major, SWIZZLE_128B_BASE32B: 4-row groups of  512 B, 32-byte chunk index ^ (row & 3)
// and the three products of a layer are
//   forward      Y[s][j]  = sum_k X[s][k] Wt[j][k]     A = X  (KM),  B = W^T image (KM)
//   input grad   dX[s][k] = sum_j dZ[s][j] W[k][j]     A = dZ (KM),  B = W image, Keras [in][out] (KM)
//   weight grad  dW[k][j] = sum_s X[s][k] dZ[s][j]     A = X  (MN: rows are the reduction index), B = dZ (MN)
// so no tile is ever transposed: the sample-major rows are simply written with the swizzle of the view(s)
// that will read them.  One elected thread issues the MMAs (M = 128 samples or features, N = layer
// width, K = 8 per instruction); the accumulators are read back with tcgen05.ld, one TMEM lane (= one sample,
// or one weight row) per thread.
//
// This file starts with the building blocks and a self-test entry (brk_tc_selftest) that checks each
// descriptor form against a host product; the NeuMF phases are built from the same helpers.
#include "common.cuh"
#include "tc.cuh"

namespace ntc {

constexpr int kThreads = 128;

// byte offset of the 16-byte chunk c4 (= col / 4) of `row` in a tile of `rows` rows
__device__ __forceinline__ uint32_t km_off16(int rows, int row, int c4) {     // K-major view, SWIZZLE_128B
  const int kb = c4 >> 3, c = c4 & 7, r8 = row & 7;
  return uint32_t(kb) * uint32_t(rows) * 128u + uint32_t(row >> 3) * 1024u + uint32_t(r8) * 128u + uint32_t((c ^ r8) << 4);
}
__device__ __forceinline__ uint32_t mn_off16(int rows, int row, int c4) {     // MN-major view, SWIZZLE_128B_BASE32B
  const int kb = c4 >> 3, c32 = (c4 & 7) >> 1, half = c4 & 1, r4 = row & 3;
  return uint32_t(kb) * uint32_t(rows) * 128u + uint32_t(row >> 2) * 512u + uint32_t(r4) * 128u + uint32_t((c32 ^ r4) << 5) +
         uint32_t(half << 4);
}

// D[M x N] (+)= A * B over K, operands in the tile layout above.
//   A_MN == 0: A tile has M rows, K columns.       A_MN == 1: A tile has K rows, M columns (read transposed).
//   B_MN == 0: B tile has N rows, K columns.       B_MN == 1: B tile has K rows, N columns.
// a_rows / b_rows: ROWS of the respective tiles.  Issued by ONE thread.
template <int M, int N, int A_MN, int B_MN>
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_base, int a_rows, uint32_t b_base, int b_rows,
                                           int K, bool accumulate_first) {
  constexpr uint32_t idesc = tc::idesc_tf32_f32(M, N, A_MN, B_MN);
  for (int ks = 0; ks < K / 8; ++ks) {
    uint64_t ad, bd;
    if (A_MN == 0) ad = tc::smem_desc_sw128_ex(a_base + uint32_t(ks >> 2) * uint32_t(a_rows) * 128u + uint32_t(ks & 3) * 32u, 16, 1024);
    else           ad = tc::smem_desc_sw128_base32(a_base + uint32_t(ks) * 1024u, uint32_t(a_rows) * 128u, 512);
    if (B_MN == 0) bd = tc::smem_desc_sw128_ex(b_base + uint32_t(ks >> 2) * uint32_t(b_rows) * 128u + uint32_t(ks & 3) * 32u, 16, 1024);
    else           bd = tc::smem_desc_sw128_base32(b_base + uint32_t(ks) * 1024u, uint32_t(b_rows) * 128u, 512);
    tc::mma_tf32_ss(d_tmem, ad, bd, idesc, (ks != 0 || accumulate_first) ? 1u : 0u);
  }
}

// ---- self-test ---------------------------------------------------------------------------------------------
// Stages Ag [a_rows x a_cols] and Bg [b_rows x b_cols] (row-major fp32, cols multiples of 32, rows multiples of
// 8) into swizzled tiles, runs one product and dumps TMEM lanes 0..127 x N columns to out [128 x N].
template <int M, int N, int A_MN, int B_MN>
__global__ void __launch_bounds__(kThreads) selftest_kernel(const float* __restrict__ Ag, int a_rows, int a_cols,
                                                            const float* __restrict__ Bg, int b_rows, int b_cols,
                                                            int K, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* As = smem;
  uint8_t* Bs = As + size_t(a_rows) * a_cols * 4;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x;
  auto stage = [&](uint8_t* dst, const float* src, int rows, int cols, int mn) {
    for (int idx = t; idx < rows * (cols / 4); idx += kThreads) {
      const int r = idx / (cols / 4), c4 = idx % (cols / 4);
      const float4 v = *reinterpret_cast<const float4*>(src + size_t(r) * cols + c4 * 4);
      *reinterpret_cast<float4*>(dst + (mn ? mn_off16(rows, r, c4) : km_off16(rows, r, c4))) = v;
    }
  };
  stage(As, Ag, a_rows, a_cols, A_MN);
  stage(Bs, Bg, b_rows, b_cols, B_MN);
  if (t == 0) { tc::mbar_init(tc::smem_u32(&bar), 1); tc::fence_barrier_init(); }
  if (t < 32) tc::tmem_alloc<256>(tc::smem_u32(&tmem_slot));
  tc::fence_proxy_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    issue_gemm<M, N, A_MN, B_MN>(tmem, tc::smem_u32(As), a_rows, tc::smem_u32(Bs), b_rows, K, false);
    tc::mma_commit(tc::smem_u32(&bar));
  }
  tc::mbar_wait(tc::smem_u32(&bar), 0);
  tc::fence_after_sync();
  const int warp = t >> 5;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tc::tmem_ld_32x32_issue(tmem + (uint32_t(warp * 32) << 16) + uint32_t(c0), r);
    tc::tmem_ld_wait(r);
    for (int j = 0; j < 32 && c0 + j < N; ++j) out[size_t(t) * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (t < 32) tc::tmem_dealloc<256>(tmem);
}

}  // namespace ntc

// mode = a_mn | (b_mn << 1); M in {64, 128}; N in {16, 32, 64, 128}.  Test hook (tests/test_gpu_tc.py).
extern "C" int brk_tc_selftest(brk_ctx* ctx, int32_t M, int32_t N, int32_t K, int32_t mode, const float* A, int32_t a_rows,
                               int32_t a_cols, const float* B, int32_t b_rows, int32_t b_cols, float* out, void* stream) {
  BRK_REQUIRE(ctx && A && B && out, BRK_E_ARG, "brk_tc_selftest: null argument");
  BRK_REQUIRE(a_cols % 32 == 0 && b_cols % 32 == 0 && a_rows % 8 == 0 && b_rows % 8 == 0 && K % 8 == 0, BRK_E_ARG,
              "brk_tc_selftest: shapes");
  const size_t smem = size_t(a_rows) * a_cols * 4 + size_t(b_rows) * b_cols * 4 + 2048;
  cudaStream_t st = (cudaStream_t)stream;
#define BRK_TC_CASE(M_, N_, AM_, BM_)                                                                         \
  if (M == M_ && N == N_ && mode == (AM_ | (BM_ << 1))) {                                                     \
    BRK_CUDA(cudaFuncSetAttribute(ntc::selftest_kernel<M_, N_, AM_, BM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); \
    ntc::selftest_kernel<M_, N_, AM_, BM_><<<1, ntc::kThreads, smem, st>>>(A, a_rows, a_cols, B, b_rows, b_cols, K, out); \
    BRK_LAUNCH_CHECK();                                                                                       \
    return 0;                                                                                                 \
  }
  BRK_TC_CASE(128, 64, 0, 0) BRK_TC_CASE(128, 32, 0, 0) BRK_TC_CASE(128, 16, 0, 0) BRK_TC_CASE(128, 128, 0, 0)
  BRK_TC_CASE(128, 64, 1, 1) BRK_TC_CASE(128, 32, 1, 1) BRK_TC_CASE(64, 32, 1, 1) BRK_TC_CASE(64, 16, 1, 1)
  BRK_TC_CASE(64, 32, 0, 0) BRK_TC_CASE(128, 64, 1, 0) BRK_TC_CASE(128, 64, 0, 1)
#undef BRK_TC_CASE
  brk_set_error("brk_tc_selftest: no instance for M=%d N=%d mode=%d", M, N, mode);
  return BRK_E_ARG;
}

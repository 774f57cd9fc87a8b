// Shared-memory operand tiles of the TF32 tensor-core NeuMF kernels (neumf_tc.cu, neumf_fused.cu): the two swizzled
// sample-major layouts, the descriptor walk of one product, and TMEM read helpers.
//
// Every activation / gradient tile of 128 samples is staged sample-major (rows of 32 floats = 128 B, one region of
// ROWS*128 B per block of 32 columns), in one of the two swizzles the tensor core reads fp32 (TF32) operands through:
//   "KM" K-major, SWIZZLE_128B        : 8-row groups of 1024 B, 16-byte chunk index ^ (row & 7)
//   "MN" MN-major, SWIZZLE_128B_BASE32B: 4-row groups of  512 B, 32-byte chunk index ^ (row & 3)
#pragma once
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

__host__ __device__ constexpr int pad32(int x) { return (x + 31) & ~31; }

// TMEM lane that holds row m of an M-row accumulator (M = 64 uses 16 lanes of each 32-lane quadrant)
template <int M> __device__ __forceinline__ int row_of_lane(int lane128) {
  if (M == 128) return lane128;
  return (lane128 & 31) < 16 ? (lane128 >> 5) * 16 + (lane128 & 15) : -1;
}

}  // namespace ntc

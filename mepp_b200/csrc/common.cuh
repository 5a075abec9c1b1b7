// Shared helpers for the mepp_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/mepp_b200.h"

namespace mepp {

void set_error(const std::string& msg);
int fail(const std::string& msg);

#define MEPP_CUDA_CHECK(expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      return ::mepp::fail(std::string(#expr) + ": " + cudaGetErrorString(_e));       \
    }                                                                                \
  } while (0)

#define MEPP_REQUIRE(cond, msg)                                                      \
  do {                                                                               \
    if (!(cond)) return ::mepp::fail(std::string("mepp_b200: ") + msg);              \
  } while (0)

// ---- operand tiling (see include/mepp_b200.h) -------------------------------
constexpr int kBlockM = 128;          // rows (task*L + position) per tile
constexpr int kBlockK = 32;           // sequences per k-block
constexpr int kChunks = kBlockK / 8;  // 16-byte chunks (8 halves) per k-block
constexpr int kATileBytes = 2 * kChunks * kBlockM * 16;  // 16384: hi+lo planes
// Behind the tiles the A operand carries a bit array: per m-tile FW words, bit (2*k_block + h) is set
// by the scan kernel when the tile holds a non-zero value among sequences [16h, 16h+16) of the k-block;
// the GEMM neither loads nor multiplies all-zero blocks.
__host__ __device__ inline int64_t flag_words(int64_t KB) { return ((2 * KB + 31) / 32 + 3) / 4 * 4; }
__host__ __device__ inline int64_t a_flags_offset(int64_t n_tiles_m, int64_t KB) {
  return n_tiles_m * KB * (int64_t)kATileBytes;
}

__host__ __device__ inline int64_t a_tile_offset(int64_t m_tile, int64_t k_block, int64_t KB) {
  return (m_tile * KB + k_block) * (int64_t)kATileBytes;
}
// byte offset inside an A tile of the 16-byte chunk (plane, kc, row)
__host__ __device__ inline int a_chunk_offset(int plane, int kc, int row) {
  return ((plane * kChunks + kc) * kBlockM + row) * 16;
}
__host__ __device__ inline int b_tile_bytes(int BN) { return 2 * kChunks * BN * 16; }
__host__ __device__ inline int64_t b_tile_offset(int64_t n_tile, int64_t k_block, int64_t KB, int BN) {
  return (n_tile * KB + k_block) * (int64_t)b_tile_bytes(BN);
}
__host__ __device__ inline int b_chunk_offset(int plane, int kc, int col, int BN) {
  return ((plane * kChunks + kc) * BN + col) * 16;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Split a float into fp16 hi + lo (hi + lo represents v to ~2^-22 relative).
__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

}  // namespace mepp

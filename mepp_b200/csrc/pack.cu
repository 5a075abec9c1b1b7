// K0: ASCII -> 3-bit symbol stream (2-bit code + N flag) + global GC ratio, and the Y (score / permuted
// score) GEMM operand.  Replaces mepp/onehot_dna.py:61-70, mepp/utils.py:141-194,
// mepp/utils.py:275-308 and mepp/permutations.py:5-58 of the reference.
#include "common.cuh"

namespace mepp {

// 3-bit symbol of one ASCII base: A,C,G,T = 0..3 (onehot_dna.py:3), everything else (N, lower case
// soft-masked bases, IUPAC codes) = 4, the reference's "N" rule (onehot_dna.py:23-30).
__device__ __forceinline__ uint32_t base_symbol(uint8_t ch) {
  return ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 4u;
}

// One thread per (stream word w, sequence n); consecutive threads = consecutive n so the
// transposed stores are coalesced.  Word w holds bits [32w, 32w+32) of the sequence's symbol
// stream (base u at bits [3u, 3u+3)), i.e. parts of up to 12 bases.
__global__ void pack_symbols_kernel(const uint8_t* __restrict__ ascii, int64_t N, int L,
                                    int64_t n_pad, uint32_t* __restrict__ sym_T) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int w = blockIdx.y;
  if (n >= n_pad) return;
  uint32_t word = 0;
  if (n < N) {
    const uint8_t* s = ascii + n * (int64_t)L;
    const int bit0 = 32 * w;
    const int u0 = bit0 / 3, u1 = (bit0 + 31) / 3;
    for (int u = u0; u <= u1; ++u) {
      if (u >= L) break;
      const uint32_t sym = base_symbol(s[u]);
      const int sh = 3 * u - bit0;
      word |= sh >= 0 ? (sym << sh) : (sym >> (-sh));
    }
  }
  sym_T[(int64_t)w * n_pad + n] = word;
}

// The GC covariate z[n] = (sum_l gc_l) / L with gc_l = 1 for C/G, 0.5 for N (= every non
// upper-case-ACGT char), 0 otherwise.
__global__ void pack_gc_kernel(const uint8_t* __restrict__ ascii, int64_t N, int L, int64_t n_pad,
                               float* __restrict__ gc) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_pad) return;
  int twice = 0;  // 2 * sum of gc_l, an exact integer
  if (n < N) {
    const uint8_t* s = ascii + n * (int64_t)L;
    for (int p = 0; p < L; ++p) {
      const uint32_t sym = base_symbol(s[p]);
      twice += sym == 4u ? 1 : ((sym == 1u || sym == 2u) ? 2 : 0);
    }
  }
  // float32 sum of {0, .5, 1} is exact; one rounding in the division (utils.py:262-266)
  gc[n] = n < N ? (0.5f * (float)twice) / (float)L : 0.0f;
}

// "Not this letter" bit planes of the symbol stream for the bit-parallel pre-filter of the scan (csrc/scan.cu): thread
// per (plane word w, sequence n).  Word w >= 1 of plane c holds bit k = [base 32 (w - 1) + k is a letter other than
// c]; N bases and positions past the end of the sequence are zero in all four planes (never a "bad" letter); word 0
// and the words behind the sequence are zero padding.  32 bases = exactly the symbol-stream words 3 (w - 1) .. + 2.
__global__ void pack_planes_kernel(const uint32_t* __restrict__ sym_T, int64_t N, int64_t n_pad, int L, int WB,
                                   uint32_t* __restrict__ planes_T) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = blockIdx.y;
  if (n >= n_pad) return;
  uint32_t pl[4] = {0u, 0u, 0u, 0u};
  const int x0 = 32 * (w - 1);
  if (w >= 1 && x0 < L && n < N) {   // (padding sequences stay zero)
    const uint32_t a = sym_T[(int64_t)(3 * (w - 1)) * n_pad + n], b = sym_T[(int64_t)(3 * (w - 1) + 1) * n_pad + n],
                   c = sym_T[(int64_t)(3 * (w - 1) + 2) * n_pad + n];
    const unsigned long long lo = (unsigned long long)a | ((unsigned long long)b << 32);   // stream bits 0..63: bases 0..20
    const unsigned long long hi = ((unsigned long long)b >> 31) | ((unsigned long long)c << 1);   // stream bits 63..95: bases 21..31
    for (int k = 0; k < 32; ++k) {
      const uint32_t sym = k < 21 ? (uint32_t)(lo >> (3 * k)) & 7u : (uint32_t)(hi >> (3 * (k - 21))) & 7u;
      if (sym < 4u && x0 + k < L) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) pl[cc] |= (cc != (int)sym ? 1u : 0u) << k;
      }
    }
  }
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) planes_T[((int64_t)w * 4 + cc) * n_pad + n] = pl[cc];
}

// ysplit[n] = fp16 hi | lo<<16 of yhat[n]; also accumulates sum and sum of squares
// of the split-rounded values in float64 (identical for every permuted column).
__global__ void y_split_kernel(const float* __restrict__ yhat, int64_t N, uint32_t* __restrict__ ysplit,
                               double* __restrict__ ysum) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (n < N) {
    __half h, l;
    split_half(yhat[n], h, l);
    ysplit[n] = (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16);
    v = (double)__half2float(h) + (double)__half2float(l);
  }
  double s = v, ss = v * v;
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0 && (s != 0.0 || ss != 0.0)) {
    atomicAdd(&ysum[0], s);
    atomicAdd(&ysum[1], ss);
  }
}

// One thread per (column c, 8 consecutive sequences): gathers the packed hi/lo pair
// and writes one 16-byte chunk to each plane of the tiled B operand.
__global__ void y_build_kernel(const uint32_t* __restrict__ ysplit, const int32_t* __restrict__ perm,
                               int64_t N, int64_t n_pad, int C, int BN, int n_tiles_n,
                               uint8_t* __restrict__ planes) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // group of 8 sequences
  int c = blockIdx.y;                                          // padded column index
  int64_t n_groups = n_pad / 8;
  if (g >= n_groups) return;
  int64_t n0 = g * 8;
  uint32_t v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t n = n0 + i;
    uint32_t x = 0;
    if (c < C && n < N) {
      int64_t src = c == 0 ? n : (int64_t)perm[(int64_t)(c - 1) * N + n];
      x = ysplit[src];
    }
    v[i] = x;
  }
  uint4 hi, lo;
  hi.x = (v[0] & 0xffffu) | (v[1] << 16);
  hi.y = (v[2] & 0xffffu) | (v[3] << 16);
  hi.z = (v[4] & 0xffffu) | (v[5] << 16);
  hi.w = (v[6] & 0xffffu) | (v[7] << 16);
  lo.x = (v[0] >> 16) | (v[1] & 0xffff0000u);
  lo.y = (v[2] >> 16) | (v[3] & 0xffff0000u);
  lo.z = (v[4] >> 16) | (v[5] & 0xffff0000u);
  lo.w = (v[6] >> 16) | (v[7] & 0xffff0000u);
  int64_t KB = n_pad / kBlockK;
  int64_t kb = n0 / kBlockK;
  int kc = (int)((n0 % kBlockK) / 8);
  int nt = c / BN, col = c % BN;
  uint8_t* tile = planes + b_tile_offset(nt, kb, KB, BN);
  *reinterpret_cast<uint4*>(tile + b_chunk_offset(0, kc, col, BN)) = hi;
  *reinterpret_cast<uint4*>(tile + b_chunk_offset(1, kc, col, BN)) = lo;
}

// col_syz[c] = sum_n y_c[n] * zhat[n] in float64; one block per column.
__global__ void y_colstats_kernel(const uint32_t* __restrict__ ysplit, const int32_t* __restrict__ perm,
                                  const float* __restrict__ zhat, int64_t N, int C,
                                  double* __restrict__ col_syz) {
  int c = blockIdx.x;
  double acc = 0.0;
  for (int64_t n = threadIdx.x; n < N; n += blockDim.x) {
    int64_t src = c == 0 ? n : (int64_t)perm[(int64_t)(c - 1) * N + n];
    uint32_t x = ysplit[src];
    double y = (double)__half2float(__ushort_as_half((unsigned short)(x & 0xffffu))) +
               (double)__half2float(__ushort_as_half((unsigned short)(x >> 16)));
    acc += y * (double)zhat[n];
  }
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) col_syz[c] = v;
  }
}

// Column statistics of the EXACT float32 operand (sparse profile path): sum and sum of squares of yhat in float64
// (identical for every permuted column) ...
__global__ void y_sum_kernel(const float* __restrict__ yhat, int64_t N, double* __restrict__ ysum) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double v = n < N ? (double)yhat[n] : 0.0;
  double s = v, ss = v * v;
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0 && (s != 0.0 || ss != 0.0)) {
    atomicAdd(&ysum[0], s);
    atomicAdd(&ysum[1], ss);
  }
}

// ... and col_syz[c] = sum_n y_c[n] * zhat[n] in float64; one block per column.
__global__ void y_colstats_exact_kernel(const float* __restrict__ yhat, const int32_t* __restrict__ perm,
                                        const float* __restrict__ zhat, int64_t N, int C,
                                        double* __restrict__ col_syz) {
  const int c = blockIdx.x;
  double acc = 0.0;
  for (int64_t n = threadIdx.x; n < N; n += blockDim.x) {
    const int64_t src = c == 0 ? n : (int64_t)perm[(int64_t)(c - 1) * N + n];
    acc += (double)yhat[src] * (double)zhat[n];
  }
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) col_syz[c] = v;
  }
}

// Row-major float32 copy of Y for the sparse profile path: y_rows[n][c] = the float32 score operand itself
// (the reference's dataset holds float32 scores, utils.py:188-192), zero in the padding rows / columns.
__global__ void y_rows_kernel(const float* __restrict__ yhat, const int32_t* __restrict__ perm, int64_t N,
                              int64_t n_pad, int C, int64_t ldY, float* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  const int64_t n = blockIdx.x;
  if (c >= ldY) return;
  float v = 0.0f;
  if (c < C && n < N) {
    const int64_t src = c == 0 ? n : (int64_t)perm[(c - 1) * N + n];
    v = yhat[src];
  }
  out[n * ldY + c] = v;
}

// Columns col0 .. col0 + n_cols - 1 of y_rows from a CHUNK of permutation rows (perm_chunk[j][n], j < n_cols; nullptr:
// the identity, i.e. the true scores of column 0): a 32 x 32 transpose through shared memory, so that the index rows
// are read and the operand rows written in full lines.
__global__ void __launch_bounds__(256) y_rows_cols_kernel(const float* __restrict__ yhat, const int32_t* __restrict__ perm_chunk,
                                                          int64_t N, int64_t n_pad, int col0, int n_cols, int64_t ldY,
                                                          float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int64_t n0 = (int64_t)blockIdx.x * 32;
  const int j0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  for (int jj = ty; jj < 32; jj += 8) {
    const int j = j0 + jj;
    const int64_t n = n0 + tx;
    float v = 0.0f;
    if (j < n_cols && n < N) v = yhat[perm_chunk ? (int64_t)perm_chunk[(int64_t)j * N + n] : n];
    tile[jj][tx] = v;
  }
  __syncthreads();
  for (int nn = ty; nn < 32; nn += 8) {
    const int64_t n = n0 + nn;
    const int j = j0 + tx;
    if (n < n_pad && j < n_cols) out[n * ldY + col0 + j] = tile[tx][nn];
  }
}

// col_syz[col0 + j] for a chunk of permutation rows (see y_colstats_exact_kernel)
__global__ void y_colstats_cols_kernel(const float* __restrict__ yhat, const int32_t* __restrict__ perm_chunk,
                                       const float* __restrict__ zhat, int64_t N, int col0, double* __restrict__ col_syz) {
  const int j = blockIdx.x;
  double acc = 0.0;
  for (int64_t n = threadIdx.x; n < N; n += blockDim.x) {
    const int64_t src = perm_chunk ? (int64_t)perm_chunk[(int64_t)j * N + n] : n;
    acc += (double)yhat[src] * (double)zhat[n];
  }
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) col_syz[col0 + j] = v;
  }
}

}  // namespace mepp

extern "C" {

int mepp_build_y_cols(const float* yhat_dev, const int32_t* perm_chunk_dev, const float* zhat_dev, int64_t N, int col0,
                      int n_cols, int C, float* y_rows_dev, double* col_syz_dev, double* ysum_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(yhat_dev && y_rows_dev && col_syz_dev && ysum_dev, "build_y_cols: null pointer");
  MEPP_REQUIRE(N > 0 && C > 0 && col0 >= 0 && n_cols > 0 && col0 + n_cols <= C && (perm_chunk_dev || (col0 == 0 && n_cols == 1)),
               "build_y_cols: bad arguments");
  MEPP_REQUIRE(mepp_device_ok(), "build_y_cols: no sm_100 CUDA device (no CPU fallback)");
  cudaStream_t st = as_stream(stream);
  const int64_t ldY = mepp_y_rows_ld(C), n_pad = mepp_pad32(N);
  MEPP_REQUIRE(n_pad / 32 < (1LL << 31) && (n_cols + 31) / 32 <= 65535, "build_y_cols: shape too large");
  if (col0 == 0) {
    // first call: zero the padding rows / columns, sum and sum of squares of the scores (the same for every column)
    MEPP_CUDA_CHECK(cudaMemsetAsync(y_rows_dev, 0, sizeof(float) * (size_t)n_pad * (size_t)ldY, st));
    MEPP_CUDA_CHECK(cudaMemsetAsync(ysum_dev, 0, 2 * sizeof(double), st));
    y_sum_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(yhat_dev, N, ysum_dev);
  }
  dim3 grid((unsigned)(n_pad / 32), (unsigned)((n_cols + 31) / 32));
  y_rows_cols_kernel<<<grid, 256, 0, st>>>(yhat_dev, perm_chunk_dev, N, n_pad, col0, n_cols, ldY, y_rows_dev);
  if (zhat_dev) {
    y_colstats_cols_kernel<<<(unsigned)n_cols, 256, 0, st>>>(yhat_dev, perm_chunk_dev, zhat_dev, N, col0, col_syz_dev);
  } else {
    MEPP_CUDA_CHECK(cudaMemsetAsync(col_syz_dev + col0, 0, sizeof(double) * (size_t)n_cols, st));
  }
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int mepp_build_y_rows(const float* yhat_dev, const int32_t* perm_dev, int64_t N, int P, float* y_rows_dev,
                      void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(yhat_dev && y_rows_dev, "build_y_rows: null pointer");
  MEPP_REQUIRE(N > 0 && P >= 0 && (P == 0 || perm_dev), "build_y_rows: bad arguments");
  MEPP_REQUIRE(mepp_device_ok(), "build_y_rows: no sm_100 CUDA device (no CPU fallback)");
  const int C = 1 + P;
  const int64_t ldY = mepp_y_rows_ld(C), n_pad = mepp_pad32(N);
  MEPP_REQUIRE(n_pad < (1LL << 31) && (ldY + 255) / 256 <= 65535, "build_y_rows: shape too large");
  dim3 grid((unsigned)n_pad, (unsigned)((ldY + 255) / 256));
  y_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(yhat_dev, perm_dev, N, n_pad, C, ldY, y_rows_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int mepp_y_stats(const float* yhat_dev, const int32_t* perm_dev, const float* zhat_dev, int64_t N, int P,
                 double* col_syz_dev, double* ysum_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(yhat_dev && col_syz_dev && ysum_dev, "y_stats: null pointer");
  MEPP_REQUIRE(N > 0 && P >= 0 && (P == 0 || perm_dev), "y_stats: bad arguments");
  MEPP_REQUIRE(mepp_device_ok(), "y_stats: no sm_100 CUDA device (no CPU fallback)");
  cudaStream_t st = as_stream(stream);
  const int C = 1 + P;
  MEPP_CUDA_CHECK(cudaMemsetAsync(ysum_dev, 0, 2 * sizeof(double), st));
  y_sum_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(yhat_dev, N, ysum_dev);
  if (zhat_dev) {
    y_colstats_exact_kernel<<<(unsigned)C, 256, 0, st>>>(yhat_dev, perm_dev, zhat_dev, N, C, col_syz_dev);
  } else {
    MEPP_CUDA_CHECK(cudaMemsetAsync(col_syz_dev, 0, sizeof(double) * (size_t)C, st));
  }
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int mepp_pack_dna(const uint8_t* ascii_dev, int64_t N, int L, uint32_t* sym_T_dev, float* gc_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(ascii_dev && sym_T_dev && gc_dev, "pack_dna: null pointer");
  MEPP_REQUIRE(N > 0 && L > 0, "pack_dna: empty input");
  MEPP_REQUIRE(mepp_device_ok(), "pack_dna: no sm_100 CUDA device (no CPU fallback)");
  int64_t n_pad = mepp_pad32(N);
  int W3 = (int)mepp_sym_words(L);
  cudaStream_t st = as_stream(stream);
  dim3 grid((unsigned)((n_pad + 255) / 256), (unsigned)W3);
  pack_symbols_kernel<<<grid, 256, 0, st>>>(ascii_dev, N, L, n_pad, sym_T_dev);
  pack_gc_kernel<<<(unsigned)((n_pad + 127) / 128), 128, 0, st>>>(ascii_dev, N, L, n_pad, gc_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int mepp_build_y(const float* yhat_dev, const int32_t* perm_dev, const float* zhat_dev, int64_t N,
                 int P, void* y_planes_dev, double* col_syz_dev, double* ysum_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(yhat_dev && y_planes_dev && col_syz_dev && ysum_dev, "build_y: null pointer");
  MEPP_REQUIRE(N > 0 && P >= 0 && (P == 0 || perm_dev), "build_y: bad arguments");
  MEPP_REQUIRE(mepp_device_ok(), "build_y: no sm_100 CUDA device (no CPU fallback)");
  cudaStream_t st = as_stream(stream);
  int C = 1 + P;
  int n_tiles_n = 0;
  int BN = mepp_choose_bn(C, &n_tiles_n);
  int64_t n_pad = mepp_pad32(N);
  // scratch behind the tiles (mepp_y_planes_bytes reserves it): no stream-ordered allocation per call, whose pool
  // trimming made single staging calls take 10-400 ms at random
  uint32_t* ysplit = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(y_planes_dev) +
                                                 (int64_t)n_tiles_n * (n_pad / kBlockK) * (int64_t)b_tile_bytes(BN));
  MEPP_CUDA_CHECK(cudaMemsetAsync(ysum_dev, 0, 2 * sizeof(double), st));
  y_split_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(yhat_dev, N, ysplit, ysum_dev);
  dim3 grid((unsigned)((n_pad / 8 + 127) / 128), (unsigned)(n_tiles_n * BN));
  y_build_kernel<<<grid, 128, 0, st>>>(ysplit, perm_dev, N, n_pad, C, BN, n_tiles_n,
                                       reinterpret_cast<uint8_t*>(y_planes_dev));
  if (zhat_dev) {
    y_colstats_kernel<<<(unsigned)C, 256, 0, st>>>(ysplit, perm_dev, zhat_dev, N, C, col_syz_dev);
  } else {
    MEPP_CUDA_CHECK(cudaMemsetAsync(col_syz_dev, 0, sizeof(double) * (size_t)C, st));
  }
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int64_t mepp_plane_words(int L) { return ((int64_t)L + 31) / 32 + 4; }

int mepp_pack_planes(const uint32_t* sym_T_dev, int64_t N, int L, uint32_t* planes_T_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(sym_T_dev && planes_T_dev, "pack_planes: null pointer");
  MEPP_REQUIRE(N > 0 && L > 0, "pack_planes: empty input");
  MEPP_REQUIRE(mepp_device_ok(), "pack_planes: no sm_100 CUDA device (no CPU fallback)");
  const int64_t n_pad = mepp_pad32(N);
  const int WB = (int)mepp_plane_words(L);
  dim3 grid((unsigned)((n_pad + 255) / 256), (unsigned)WB);
  pack_planes_kernel<<<grid, 256, 0, as_stream(stream)>>>(sym_T_dev, N, n_pad, L, WB, planes_T_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"

// K2s: the profile contraction S = X . Y for a SPARSE X, with the correlation epilogue fused.
//
// Replaces the same reference lines as the tcgen05 GEMM of gemm.cu (the (B, L, 1+P) x*y reduce_sum of
// mepp/core.py:250 and the r_ab / partial-correlation formulas of core.py:182-193,255-265, nan_to_num of
// core.py:602).  With the reference's default match threshold (p = 1e-4 per strand) the blurred score
// matrix X holds ~1e-3 non-zeros, so a row of S is the sum of a few dozen rows of Y:
//     S[r, :] = sum over the non-zeros (n, x) of row r of   x * Y[n, :]
// The scan kernel appends one record per non-zero; here they are binned by row (counting sort: offsets
// by prefix sum, fill with a per-row cursor) and every warp gathers the Y rows of one X row for a slab of
// 512 columns, accumulating exact fp32 x fp32 products in float64.  Y (n_pad x ldY float32, 80 MB at
// N = 20 000, P = 1000) stays L2-resident; the kernel is bound by L2 -> SM bandwidth (4 KB per non-zero).
#include "common.cuh"
#include <cfloat>

namespace mepp {

struct XEntry { int32_t row, seq; float x; uint32_t pad; };

constexpr int kScanThreads = 1024;

// offsets[0 .. m_rows] = exclusive prefix sum of row_nnz[0 .. m_rows) (one CTA; m_rows is at most a few 1e5)
// offsets[m_rows + 1] = 1 if a sub-list of the entry list overflowed (the kernels below then do nothing).
__global__ void __launch_bounds__(kScanThreads) csr_offsets_kernel(const int32_t* __restrict__ row_nnz, int64_t m_rows,
                                                                   const unsigned long long* __restrict__ entry_count,
                                                                   int64_t sub_cap, int32_t* __restrict__ offsets) {
  __shared__ int32_t s_warp[32];
  __shared__ int32_t s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { offsets[m_rows + 1] = 0; s_base = 0; }
  __syncthreads();
  if (tid < MEPP_ENTRY_LISTS && (int64_t)entry_count[tid * 16] > sub_cap) offsets[m_rows + 1] = 1;
  // tiles of 4096 rows: four consecutive rows per thread (coalesced), scan inside the thread, the warp, the block
  for (int64_t tile = 0; tile < m_rows; tile += 4 * kScanThreads) {
    const int64_t i0 = tile + 4 * tid;
    int32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = i0 + k < m_rows ? row_nnz[i0 + k] : 0;
    const int32_t mine = v[0] + v[1] + v[2] + v[3];
    int32_t inc = mine;
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int32_t w = s_warp[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      s_warp[lane] = w;                               // inclusive scan of the warp totals
    }
    __syncthreads();
    int32_t run = s_base + (warp ? s_warp[warp - 1] : 0) + inc - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < m_rows) offsets[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (tid == kScanThreads - 1) s_base = run;
    __syncthreads();
  }
  if (tid == 0) offsets[m_rows] = s_base;
}

// csr[offsets[row] + k] = (seq, x) for the k-th record of the row that arrives (cursor = row_nnz counted down)
__global__ void csr_fill_kernel(const XEntry* __restrict__ entries, const unsigned long long* __restrict__ entry_count,
                                int64_t sub_cap, int64_t m_rows, const int32_t* __restrict__ offsets,
                                int32_t* __restrict__ row_nnz, uint2* __restrict__ csr) {
  if (offsets[m_rows + 1] != 0) return;               // overflow: the caller falls back to the tiled path
  const int list = blockIdx.y;
  const int64_t n = (int64_t)entry_count[list * 16];
  const XEntry* src = entries + (int64_t)list * sub_cap;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src + i));
    const int row = (int)raw.x;
    const int k = atomicSub(row_nnz + row, 1) - 1;
    csr[offsets[row] + k] = make_uint2(raw.y, raw.z);
  }
}

struct SparseArgs {
  const uint2* csr; const int32_t* offsets;
  const float* y_rows; int64_t ldY;
  const double* rowstats; const double* colconst;
  float* r; int64_t ldr;
  int64_t m_rows; int C; int n_slabs;
};

// float -> double without the F2F pipe (16 lanes / clk / SM on B200, a quarter of the DFMA rate; measured with
// tools/micro/fp64_rate.cu): rebias the exponent with integer ops.  Exact for normal floats; +-0 maps to
// +-2^-127 (Y holds no denormals, and a zero of Y then contributes |x| * 6e-39 to a sum of magnitude >> 1).
__device__ __forceinline__ double f2d(float f) {
  const uint32_t u = __float_as_uint(f);
  const uint32_t hi = ((((u >> 3) & 0x0fffffffu) + 0x38000000u)) | (u & 0x80000000u);
  return __hiloint2double((int)hi, (int)(u << 29));
}

constexpr int kSlabCols = 512;   // columns of S per warp: lane holds slab*512 + j*128 + lane*4 + {0..3}, j < 4

__device__ __forceinline__ void fma4(double (&acc)[16], double x, const float4 (&v)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    acc[4 * j + 0] = fma(x, f2d(v[j].x), acc[4 * j + 0]);
    acc[4 * j + 1] = fma(x, f2d(v[j].y), acc[4 * j + 1]);
    acc[4 * j + 2] = fma(x, f2d(v[j].z), acc[4 * j + 2]);
    acc[4 * j + 3] = fma(x, f2d(v[j].w), acc[4 * j + 3]);
  }
}

// One warp = one (row, slab of 512 columns); four non-zeros (16 x 512 B of Y) in flight per warp.
__global__ void __launch_bounds__(256, 3) profile_sparse_kernel(const __grid_constant__ SparseArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_items = ((a.m_rows + 7) / 8) * 8 * a.n_slabs;
  if (a.offsets[a.m_rows + 1] != 0) return;            // overflow: results come from the tiled path instead
  const int64_t ld4 = a.ldY >> 2;
  const double* cc = a.colconst;
  const double n = cc[0], sy = cc[1], ay = cc[2], sz = cc[3], az = cc[4];
  const bool use_z = cc[5] != 0.0;
  for (int64_t item = warp0; item < n_items; item += n_warps) {
    // items are ordered so that the eight warps of a CTA take eight consecutive rows of one slab: the 2 * margin + 1
    // rows a match touches gather the same row of Y, which then comes from L1 for all but the first of them
    const int64_t blk = item / (8 * a.n_slabs);
    const int rem = (int)(item - blk * 8 * a.n_slabs);
    const int slab = rem >> 3;
    const int64_t row = blk * 8 + (rem & 7);
    if (row >= a.m_rows) continue;
    const int beg = a.offsets[row], end = a.offsets[row + 1];
    const float4* yb = reinterpret_cast<const float4*>(a.y_rows) + (int64_t)slab * (kSlabCols / 4) + lane;
    double acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.0;
    auto rec = [&](int e) { return e < end ? __ldg(a.csr + e) : make_uint2(0u, 0u); };   // (row 0, x = 0) past the end
    uint2 c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k] = rec(beg + k);
#pragma unroll 1
    for (int e = beg; e < end; e += 4) {
      float4 v[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4* y = yb + (int64_t)c[k].x * ld4;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[k][j] = __ldg(y + j * 32);
      }
      uint2 nx[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) nx[k] = rec(e + 4 + k);
#pragma unroll
      for (int k = 0; k < 4; ++k) fma4(acc, (double)__uint_as_float(c[k].y), v[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) c[k] = nx[k];
    }
    // correlation epilogue: core.py:182-193,255-265 in float64 + np.nan_to_num (core.py:602)
    const double sx = a.rowstats[row * 3 + 0], sxx = a.rowstats[row * 3 + 1], sxz = a.rowstats[row * 3 + 2];
    const double ax = n * sxx - sx * sx;
    const double inv_xy = 1.0 / sqrt(ax * ay);
    const double sxsy = sx * sy;
    double r_xz = 0.0, row_den = 1.0;
    if (use_z) {
      r_xz = (n * sxz - sx * sz) / sqrt(ax * az);
      row_den = 1.0 / sqrt(1.0 - r_xz * r_xz);
    }
    float* dst = a.r + row * a.ldr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c0 = slab * kSlabCols + j * 128 + lane * 4;
      if (c0 < a.ldr) {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int col = c0 + k;
          float out = 0.0f;
          if (col < a.C) {
            double rr = (n * acc[4 * j + k] - sxsy) * inv_xy;
            if (use_z) {
              const double2 cz = __ldg(reinterpret_cast<const double2*>(cc + 8 + 2 * col));
              rr = (rr - r_xz * cz.x) * row_den * cz.y;
            }
            out = (float)rr;
            if (out != out) out = 0.0f;
            else if (isinf(out)) out = out > 0 ? FLT_MAX : -FLT_MAX;
          }
          o[k] = out;
        }
        *reinterpret_cast<float4*>(dst + c0) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

}  // namespace mepp

extern "C" {

int64_t mepp_y_rows_ld(int C) { return ((int64_t)C + mepp::kSlabCols - 1) / mepp::kSlabCols * mepp::kSlabCols; }

int64_t mepp_profile_sparse_work_bytes(int64_t m_rows, int64_t entry_cap) {
  return ((m_rows + 2) * 4 + 255) / 256 * 256 + entry_cap * 8;
}

int mepp_profile_sparse(const void* entries_dev, const unsigned long long* entry_count_dev, int64_t entry_cap,
                        int32_t* row_nnz_dev, const float* y_rows_dev, const double* rowstats_dev,
                        const double* colconst_dev, float* r_dev, int64_t ldr, int64_t m_rows, int C, void* work_dev,
                        void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(entries_dev && entry_count_dev && row_nnz_dev && y_rows_dev && rowstats_dev && colconst_dev && r_dev &&
               work_dev, "profile_sparse: null pointer");
  MEPP_REQUIRE(m_rows > 0 && C > 0 && entry_cap >= MEPP_ENTRY_LISTS && entry_cap % MEPP_ENTRY_LISTS == 0 && ldr >= C &&
               ldr % 4 == 0, "profile_sparse: bad arguments");
  MEPP_REQUIRE(mepp_device_ok(), "profile_sparse: no sm_100 CUDA device (no CPU fallback)");
  cudaStream_t st = as_stream(stream);
  uint8_t* w = reinterpret_cast<uint8_t*>(work_dev);
  int32_t* offsets = reinterpret_cast<int32_t*>(w);
  uint2* csr = reinterpret_cast<uint2*>(w + ((m_rows + 2) * 4 + 255) / 256 * 256);
  const int64_t sub_cap = entry_cap / MEPP_ENTRY_LISTS;
  csr_offsets_kernel<<<1, kScanThreads, 0, st>>>(row_nnz_dev, m_rows, entry_count_dev, sub_cap, offsets);
  csr_fill_kernel<<<dim3(32, MEPP_ENTRY_LISTS), 256, 0, st>>>(reinterpret_cast<const XEntry*>(entries_dev), entry_count_dev,
                                                              sub_cap, m_rows, offsets, row_nnz_dev, csr);
  SparseArgs a;
  a.csr = csr; a.offsets = offsets;
  a.y_rows = y_rows_dev; a.ldY = mepp_y_rows_ld(C);
  a.rowstats = rowstats_dev; a.colconst = colconst_dev; a.r = r_dev; a.ldr = ldr;
  a.m_rows = m_rows; a.C = C; a.n_slabs = (int)(a.ldY / kSlabCols);
  profile_sparse_kernel<<<148 * 16, 256, 0, st>>>(a);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"

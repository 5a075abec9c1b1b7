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
#include <stdlib.h>

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


// ------------------------------------------------------------------------------------------------------------------
// Match-driven form.  The blur is linear and acts along the positions, the contraction along the sequences, so they
// commute:   sum_n xbar[n, p] y[n, c]  =  (1 / k) sum_{|d| <= margin} ( sum_n x0[n, p + d] y[n, c] ).
// Gathering one row of Y per MATCH instead of one per non-zero of the blurred matrix moves 2 margin + 1 = 5 times
// fewer bytes and converts 5 times fewer floats -- at cfg 3 Y is 410 MB and comes from HBM (65 GB per step for the
// entry form).  Two passes, each with the simplest possible inner loop:
//   match_gather_kernel: S0[row, :] = sum over the matches (n, x0) of row = task * L + position of (x0 / k) * Y[n, :]
//                        (float64, exact products; the loop of profile_sparse_kernel without its epilogue; rows
//                        without a match are neither computed nor written),
//   match_blur_kernel:   window sums of S0 along the positions of a task (each S0 row is read once: a CTA walks a block
//                        of positions with the last k rows in a shared-memory ring) and the correlation epilogue.
// S0 costs 8 bytes per (row, column) written and read once (2 x 2.6 GB per cfg 3 step), against the 52 GB of gathers it
// saves.  x0 / k is rounded to float32 like the blurred value it stands for (match_fill_kernel), so the only difference
// to the entry form are the windows that hold two matches of one sequence (the row statistics sum(xbar), sum(xbar^2),
// sum(xbar z) always come from the float32 blur).
struct HitRec4 { int32_t task, seq, pos; float x0; };

// offsets[0 .. m_rows] = exclusive prefix sum of count; offsets[m_rows + 1] = 1 when the match list overflowed;
// cursor[0 .. m_rows) = 0
__global__ void __launch_bounds__(kScanThreads) match_offsets_kernel(const int32_t* __restrict__ count, int64_t m_rows,
                                                                     const unsigned long long* __restrict__ hit_count,
                                                                     int64_t hits_cap, int32_t* __restrict__ offsets,
                                                                     int32_t* __restrict__ cursor) {
  __shared__ int32_t s_warp[32];
  __shared__ int32_t s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { offsets[m_rows + 1] = (int64_t)*hit_count > hits_cap ? 1 : 0; s_base = 0; }
  __syncthreads();
  for (int64_t tile = 0; tile < m_rows; tile += 4 * kScanThreads) {
    const int64_t i0 = tile + 4 * tid;
    int32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = i0 + k < m_rows ? count[i0 + k] : 0;
      if (i0 + k < m_rows) cursor[i0 + k] = 0;
    }
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
      s_warp[lane] = w;
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

// csr[offsets[row] + j] = (seq, x0 / k in float32) for the j-th match of row = task * L + pos that arrives.  x0 / k
// rounded to float32 IS the blurred value of every row of the window of a match that is alone in it (the float32 sum of
// the window is x0 itself), so the products below are the very numbers the entry form multiplies; only windows that
// hold two matches of one sequence differ, by the rounding of one float32 addition.
__global__ void match_fill_kernel(const HitRec4* __restrict__ hits, const unsigned long long* __restrict__ hit_count, int L,
                                  int64_t m_rows, const int32_t* __restrict__ offsets, int32_t* __restrict__ cursor,
                                  float kf, uint2* __restrict__ csr) {
  if (offsets[m_rows + 1] != 0) return;
  const int64_t n = (int64_t)*hit_count;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(hits + i));
    const int64_t row = (int64_t)(int)raw.x * L + (int)raw.z;
    const int j = atomicAdd(cursor + row, 1);
    const float x0 = __uint_as_float(raw.w);
    csr[offsets[row] + j] = make_uint2(raw.y, __float_as_uint(kf > 1.0f ? __fdiv_rn(x0, kf) : x0));
  }
}

constexpr int kGatherMixDefault = 2;
struct GatherArgs {
  const uint2* csr; const int32_t* offsets;
  const float* y_rows; int64_t ldY;
  double* S0;                      // [m_rows][ldY]
  int64_t m_rows; int n_slabs;
};

// One warp = one (row with matches, slab of 128 * J columns); four matches in flight per warp.  Items run slab-major: all
// warps work on the same slab of Y at about the same time.  J = 4 (512 columns) when Y fits the L2 anyway; J = 1 when it
// does not (cfg 3: 410 MB): a 128-column slab of Y is N * 512 bytes (51 MB), stays L2-resident while the rows of the
// group pass over it, and the gather then reads HBM once per slab instead of once per match.
// MIX: two of the four conversions of a float4 go through the F2F pipe, two through the integer pipe (f2d): the two
// pipes work side by side (MEPP_GATHER_MIX).
template <int J, int MIX>      // MIX = conversions per float4 that take the F2F pipe (0 .. 4)
__global__ void __launch_bounds__(256, J == 1 ? 4 : 3) match_gather_kernel(const __grid_constant__ GatherArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_items = a.m_rows * a.n_slabs;
  if (a.offsets[a.m_rows + 1] != 0) return;            // match list overflow: the caller redoes the group
  const int64_t ld4 = a.ldY >> 2;
  for (int64_t item = warp0; item < n_items; item += n_warps) {
    const int slab = (int)(item / a.m_rows);
    const int64_t row = item - (int64_t)slab * a.m_rows;
    const int beg = __ldg(a.offsets + row), end = __ldg(a.offsets + row + 1);
    if (beg == end) continue;
    const float4* yb = reinterpret_cast<const float4*>(a.y_rows) + (int64_t)slab * (32 * J) + lane;
    double acc[4 * J];
#pragma unroll
    for (int j = 0; j < 4 * J; ++j) acc[j] = 0.0;
    auto rec = [&](int e) { return e < end ? __ldg(a.csr + e) : make_uint2(0u, 0u); };   // (row 0, x = 0) past the end
    uint2 c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k] = rec(beg + k);
#pragma unroll 1
    for (int e = beg; e < end; e += 4) {
      float4 v[4][J];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4* y = yb + (int64_t)c[k].x * ld4;
#pragma unroll
        for (int j = 0; j < J; ++j) v[k][j] = __ldg(y + j * 32);
      }
      uint2 nx[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) nx[k] = rec(e + 4 + k);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double x = (double)__uint_as_float(c[k].y);
#pragma unroll
        for (int j = 0; j < J; ++j) {
          acc[4 * j + 0] = fma(x, MIX >= 4 ? (double)v[k][j].x : f2d(v[k][j].x), acc[4 * j + 0]);
          acc[4 * j + 1] = fma(x, MIX >= 3 ? (double)v[k][j].y : f2d(v[k][j].y), acc[4 * j + 1]);
          acc[4 * j + 2] = fma(x, MIX >= 2 ? (double)v[k][j].z : f2d(v[k][j].z), acc[4 * j + 2]);
          acc[4 * j + 3] = fma(x, MIX >= 1 ? (double)v[k][j].w : f2d(v[k][j].w), acc[4 * j + 3]);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) c[k] = nx[k];
    }
    double2* dst = reinterpret_cast<double2*>(a.S0 + row * a.ldY + (int64_t)slab * (128 * J) + lane * 4);
#pragma unroll
    for (int j = 0; j < J; ++j) {                       // streaming stores: S0 must not evict the slab of Y from the L2
      __stcs(dst + j * 64, make_double2(acc[4 * j + 0], acc[4 * j + 1]));
      __stcs(dst + j * 64 + 1, make_double2(acc[4 * j + 2], acc[4 * j + 3]));
    }
  }
}

struct BlurArgs {
  const double* S0; int64_t ldS0; const int32_t* offsets;
  const double* rowstats; const double* colconst;
  float* r; int64_t ldr;
  int T, L, C, margin, pb, n_pb, n_slabs;
};

constexpr int kBlurThreads = 128;                 // one thread = two adjacent columns
constexpr int kBlurCols = 2 * kBlurThreads;       // columns of r per CTA
constexpr int kBlurMaxPb = 128;                   // positions per CTA, at most

// CTA = (task, block of `pb` positions, slab of 256 columns).  It walks the positions p0 - margin .. p1 + margin - 1
// once: row q of S0 enters a ring of the last k rows (shared memory), and row p = q - margin of r is the sum of the
// ring in position order, divided by k, through the correlation epilogue (core.py:182-193,255-265 in float64 +
// np.nan_to_num, core.py:602).  Rows of S0 without a match were never written: they are known from the CSR offsets.
__global__ void __launch_bounds__(kBlurThreads) match_blur_kernel(const __grid_constant__ BlurArgs a) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  __shared__ double s_rc[kBlurMaxPb][4];            // per row of the block: n / sqrt(ax ay) ..., see below
  __shared__ uint8_t s_has[kBlurMaxPb + 2 * MEPP_MAX_MARGIN];
  if (a.offsets[(int64_t)a.T * a.L + 1] != 0) return;           // match list overflow: the caller redoes the group
  const int tid = threadIdx.x;
  const int k = 2 * a.margin + 1;
  double2* ring = reinterpret_cast<double2*>(s_dyn) + tid;      // ring[j * kBlurThreads]
  int64_t b = blockIdx.x;
  const int slab = (int)(b % a.n_slabs); b /= a.n_slabs;
  const int pbi = (int)(b % a.n_pb);
  const int t = (int)(b / a.n_pb);
  const int p0 = pbi * a.pb, p1 = p0 + a.pb < a.L ? p0 + a.pb : a.L;
  const int q0 = p0 - a.margin, q1 = p1 + a.margin;             // rows of S0 this CTA walks
  const double* cc = a.colconst;
  const double n = cc[0], sy = cc[1], ay = cc[2], sz = cc[3], az = cc[4];
  const bool use_z = cc[5] != 0.0;
  const int64_t row_t = (int64_t)t * a.L;
  for (int i = tid; i < q1 - q0; i += kBlurThreads) {
    const int q = q0 + i;
    s_has[i] = (q >= 0 && q < a.L) ? (__ldg(a.offsets + row_t + q + 1) != __ldg(a.offsets + row_t + q)) : 0;
  }
  for (int i = tid; i < p1 - p0; i += kBlurThreads) {
    const double* rs = a.rowstats + (row_t + p0 + i) * 3;
    const double sx = rs[0], sxx = rs[1], sxz = rs[2];
    const double ax = n * sxx - sx * sx;
    double r_xz = 0.0, row_den = 1.0;
    if (use_z) {
      r_xz = (n * sxz - sx * sz) / sqrt(ax * az);
      row_den = 1.0 / sqrt(1.0 - r_xz * r_xz);
    }
    s_rc[i][0] = 1.0 / sqrt(ax * ay); s_rc[i][1] = sx * sy; s_rc[i][2] = r_xz; s_rc[i][3] = row_den;
  }
  for (int j = 0; j < k; ++j) ring[j * kBlurThreads] = make_double2(0.0, 0.0);
  __syncthreads();
  const int c0 = slab * kBlurCols + 2 * tid;
  const bool store = c0 < a.ldr;
  double2 cz0 = make_double2(0.0, 1.0), cz1 = make_double2(0.0, 1.0);
  if (use_z) {
    if (c0 < a.C) cz0 = __ldg(reinterpret_cast<const double2*>(cc + 8 + 2 * c0));
    if (c0 + 1 < a.C) cz1 = __ldg(reinterpret_cast<const double2*>(cc + 8 + 2 * (c0 + 1)));
  }
  const double* src = a.S0 + row_t * a.ldS0 + c0;
  int slot = 0;
  int live_rows = 0;                      // rows with a match among the k rows of the ring
  constexpr int kAhead = 8;               // rows requested before the first of them is used
  for (int qb = q0; qb < q1; qb += kAhead) {
    double2 v[kAhead];
#pragma unroll
    for (int u = 0; u < kAhead; ++u) {
      const int q = qb + u;
      v[u] = make_double2(0.0, 0.0);
      if (q < q1 && s_has[q - q0]) v[u] = __ldg(reinterpret_cast<const double2*>(src + (int64_t)q * a.ldS0));
    }
#pragma unroll
    for (int u = 0; u < kAhead; ++u) {
      const int q = qb + u;
      if (q >= q1) break;
      // row q enters, row q - k leaves
      live_rows += (int)s_has[q - q0] - (q - k >= q0 ? (int)s_has[q - k - q0] : 0);
      ring[slot * kBlurThreads] = v[u];
      if (++slot == k) slot = 0;
      const int p = q - a.margin;
      if (p < p0) continue;
      float o0 = 0.0f, o1 = 0.0f;
      // core.py:79-91 zeroes the first / last `margin` rows of the blurred matrix
      const bool live = p >= a.margin && p < a.L - a.margin;
      double w0 = 0.0, w1 = 0.0;
      if (live && live_rows > 0) {
        int j = slot;                                           // oldest row first: position order
        for (int i = 0; i < k; ++i) {
          const double2 x = ring[j * kBlurThreads];
          w0 += x.x; w1 += x.y;
          if (++j == k) j = 0;
        }
      }
      const double inv_xy = s_rc[p - p0][0], sxsy = s_rc[p - p0][1];
      double r0 = (n * w0 - sxsy) * inv_xy, r1 = (n * w1 - sxsy) * inv_xy;
      if (use_z) {
        const double r_xz = s_rc[p - p0][2], row_den = s_rc[p - p0][3];
        r0 = (r0 - r_xz * cz0.x) * row_den * cz0.y;
        r1 = (r1 - r_xz * cz1.x) * row_den * cz1.y;
      }
      if (c0 < a.C) {
        o0 = (float)r0;
        if (o0 != o0) o0 = 0.0f;
        else if (isinf(o0)) o0 = o0 > 0 ? FLT_MAX : -FLT_MAX;
      }
      if (c0 + 1 < a.C) {
        o1 = (float)r1;
        if (o1 != o1) o1 = 0.0f;
        else if (isinf(o1)) o1 = o1 > 0 ? FLT_MAX : -FLT_MAX;
      }
      if (store) *reinterpret_cast<float2*>(a.r + (row_t + p) * a.ldr + c0) = make_float2(o0, o1);
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

int64_t mepp_profile_matches_work_bytes(int64_t m_rows, int64_t hits_cap, int C) {
  return ((m_rows + 2) * 4 + 255) / 256 * 256 + (m_rows * 4 + 255) / 256 * 256 + (hits_cap * 8 + 255) / 256 * 256 +
         m_rows * mepp_y_rows_ld(C) * 8;
}

int mepp_profile_matches(const void* hits_dev, const unsigned long long* hit_count_dev, int64_t hits_cap,
                         const int32_t* count_dev, int T, int L, int margin, const float* y_rows_dev, int64_t n_pad_rows,
                         const double* rowstats_dev, const double* colconst_dev, float* r_dev, int64_t ldr, int C,
                         void* work_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(hits_dev && hit_count_dev && count_dev && y_rows_dev && rowstats_dev && colconst_dev && r_dev && work_dev,
               "profile_matches: null pointer");
  MEPP_REQUIRE(T > 0 && L > 0 && C > 0 && hits_cap > 0 && ldr >= C && ldr % 4 == 0 && margin >= 0 &&
               margin <= MEPP_MAX_MARGIN, "profile_matches: bad arguments");
  MEPP_REQUIRE(mepp_device_ok(), "profile_matches: no sm_100 CUDA device (no CPU fallback)");
  cudaStream_t st = as_stream(stream);
  const int64_t m_rows = (int64_t)T * L;
  MEPP_REQUIRE(m_rows < ((int64_t)1 << 31) - 2, "profile_matches: too many rows in one group");
  uint8_t* w = reinterpret_cast<uint8_t*>(work_dev);
  int32_t* offsets = reinterpret_cast<int32_t*>(w);
  w += ((m_rows + 2) * 4 + 255) / 256 * 256;
  int32_t* cursor = reinterpret_cast<int32_t*>(w);
  w += (m_rows * 4 + 255) / 256 * 256;
  uint2* csr = reinterpret_cast<uint2*>(w);
  w += (hits_cap * 8 + 255) / 256 * 256;
  double* S0 = reinterpret_cast<double*>(w);
  match_offsets_kernel<<<1, kScanThreads, 0, st>>>(count_dev, m_rows, hit_count_dev, hits_cap, offsets, cursor);
  match_fill_kernel<<<148 * 8, 256, 0, st>>>(reinterpret_cast<const HitRec4*>(hits_dev), hit_count_dev, L, m_rows, offsets,
                                             cursor, (float)(2 * margin + 1), csr);
  GatherArgs g;
  g.csr = csr; g.offsets = offsets; g.y_rows = y_rows_dev; g.ldY = mepp_y_rows_ld(C); g.S0 = S0;
  g.m_rows = m_rows;
  // narrow slabs when the operand does not fit the L2 (MEPP_GATHER_SLAB = 128 / 512 overrides)
  static const int slab_env = [] { const char* e = getenv("MEPP_GATHER_SLAB"); return e ? atoi(e) : 0; }();
  const bool narrow = slab_env == 128 || (slab_env != 512 && 4.0 * (double)n_pad_rows * (double)g.ldY > 96e6);
  // (grids: measured.  The narrow form reads 2.5 GB of HBM per cfg 3 launch for the 0.4 GB of Y -- its later waves of
  // CTAs start again at slab 0 -- but a grid of exactly the resident CTAs, one sweep, was no faster (4.46 vs 4.37 ms
  // per step: the rows are too uneven for a static deal over 4.7 k warps); the wide form gained 3 % from it.  A dynamic
  // deal -- warps taking 16 consecutive rows at a time from one counter, slab-major -- was much slower (7.5 ms), and so
  // was one launch per slab (every wave of a launch on the same 51 MB of Y: 5.4 vs 4.4 ms -- eight short launches with
  // their tails cost more than the re-reads).)
  static const int sms = [] { int d = 0, n = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
                              return n > 0 ? n : 148; }();
  static const int mix = [] { const char* e = getenv("MEPP_GATHER_MIX"); return e ? atoi(e) : kGatherMixDefault; }();
  if (narrow) {
    g.n_slabs = (int)(g.ldY / 128);
    if (mix == 1) match_gather_kernel<1, 1><<<sms * 16, 256, 0, st>>>(g);
    else if (mix == 2) match_gather_kernel<1, 2><<<sms * 16, 256, 0, st>>>(g);
    else if (mix == 3) match_gather_kernel<1, 3><<<sms * 16, 256, 0, st>>>(g);
    else match_gather_kernel<1, 0><<<sms * 16, 256, 0, st>>>(g);
  } else {
    g.n_slabs = (int)(g.ldY / kSlabCols);
    if (mix == 1) match_gather_kernel<4, 1><<<sms * 3, 256, 0, st>>>(g);
    else if (mix == 2) match_gather_kernel<4, 2><<<sms * 3, 256, 0, st>>>(g);
    else if (mix == 3) match_gather_kernel<4, 3><<<sms * 3, 256, 0, st>>>(g);
    else match_gather_kernel<4, 0><<<sms * 3, 256, 0, st>>>(g);
  }
  BlurArgs a;
  a.S0 = S0; a.ldS0 = g.ldY; a.offsets = offsets; a.rowstats = rowstats_dev; a.colconst = colconst_dev;
  a.r = r_dev; a.ldr = ldr; a.T = T; a.L = L; a.C = C; a.margin = margin;
  // positions per CTA: long enough that the 2 * margin extra rows a block reads stay a small share, short enough for
  // a few thousand CTAs
  int pb = 8 * (2 * margin + 1);
  pb = pb < 32 ? 32 : pb > kBlurMaxPb ? kBlurMaxPb : pb;
  a.pb = pb; a.n_pb = (L + pb - 1) / pb;
  a.n_slabs = (int)((ldr + kBlurCols - 1) / kBlurCols);
  const size_t smem = (size_t)(2 * margin + 1) * kBlurThreads * sizeof(double2);
  MEPP_CUDA_CHECK(cudaFuncSetAttribute(match_blur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_ctas = (int64_t)T * a.n_pb * a.n_slabs;
  MEPP_REQUIRE(n_ctas < ((int64_t)1 << 31), "profile_matches: too many blocks in one group");
  match_blur_kernel<<<(unsigned)n_ctas, kBlurThreads, smem, st>>>(a);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"

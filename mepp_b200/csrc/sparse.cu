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


// ------------------------------------------------------------------------------------------------------------------
// Match-driven form (default).  The blur is linear and acts along the positions, the contraction along the sequences,
// so they commute:   sum_n xbar[n, p] y[n, c]  =  (1 / k) sum_{|d| <= margin} ( sum_n x0[n, p + d] y[n, c] ).
// Gathering one row of Y per MATCH instead of one per non-zero of the blurred matrix moves 2 margin + 1 = 5 times
// fewer bytes -- and those bytes are what bounds the kernel (at cfg 3 Y is 410 MB and comes from HBM: 65 GB per step
// for the entry form).  A warp owns 32 consecutive positions of one task and 256 columns: it walks the positions once,
// S0[q] = sum over the matches at q of x0 * Y[seq] (float64, exact products), keeps the last k rows of S0 in a
// shared-memory ring and a running window sum, and applies the correlation epilogue to sum / k.  The only difference
// to the entry form is that xbar is not rounded to float32 before it multiplies y (<= 6e-8 relative per term; the row
// statistics sum(xbar), sum(xbar^2), sum(xbar z) still come from the float32 blur, bucket_entries_kernel / the scan).
struct HitRec4 { int32_t task, seq, pos; float x0; };

constexpr int kMCols = 256;      // columns of S per warp: lane holds slab * 256 + j * 128 + lane * 4 + {0..3}, j < 2
constexpr int kMRows = 32;       // output positions per work item, at most
constexpr int kMBlockRecs = 512; // matches per work item, about (a single position may hold more)

// offsets[0 .. m_rows] = exclusive prefix sum of count; offsets[m_rows + 1] = 1 when the match list overflowed;
// cursor[0 .. m_rows) = 0
__global__ void __launch_bounds__(kScanThreads) match_offsets_kernel(const int32_t* __restrict__ count, int64_t m_rows,
                                                                     const unsigned long long* __restrict__ hit_count,
                                                                     int64_t hits_cap, int32_t* __restrict__ offsets,
                                                                     int32_t* __restrict__ cursor, int L,
                                                                     int32_t* __restrict__ blk_start) {
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
  __syncthreads();
  // Work items of the gather kernel: blocks of consecutive positions of one task, at most kMRows positions and about
  // kMBlockRecs matches each.  Real profiles are peaked (that is what MEPP looks for): a position under a peak holds
  // 50 times the average number of matches, and equal-length blocks would leave one warp with most of a task's work.
  // A row starts a block when it is a multiple of kMRows inside its task or when the running match count crosses a
  // multiple of kMBlockRecs -- a local rule, compacted with one more scan.  blk_start[0] = number of blocks,
  // blk_start[1 + b] = first row of block b, blk_start[1 + n] = m_rows.
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int64_t tile = 0; tile < m_rows; tile += 4 * kScanThreads) {
    const int64_t i0 = tile + 4 * tid;
    int32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t row = i0 + k;
      int f = 0;
      if (row < m_rows) {
        const int pp = (int)(row % L);
        f = (pp % kMRows) == 0 || (offsets[row] / kMBlockRecs) != (offsets[row - 1] / kMBlockRecs);   // (pp > 0 => row > 0)
      }
      v[k] = f;
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
      if (v[k]) blk_start[1 + run] = (int32_t)(i0 + k);
      run += v[k];
    }
    __syncthreads();
    if (tid == kScanThreads - 1) s_base = run;
    __syncthreads();
  }
  if (tid == 0) { blk_start[0] = s_base; blk_start[1 + s_base] = (int32_t)m_rows; }
}

// csr[offsets[row] + k] = (seq, x0) for the k-th match of row = task * L + pos that arrives
__global__ void match_fill_kernel(const HitRec4* __restrict__ hits, const unsigned long long* __restrict__ hit_count, int L,
                                  int64_t m_rows, const int32_t* __restrict__ offsets, int32_t* __restrict__ cursor,
                                  uint2* __restrict__ csr) {
  if (offsets[m_rows + 1] != 0) return;
  const int64_t n = (int64_t)*hit_count;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(hits + i));
    const int64_t row = (int64_t)(int)raw.x * L + (int)raw.z;
    const int k = atomicAdd(cursor + row, 1);
    csr[offsets[row] + k] = make_uint2(raw.y, raw.w);
  }
}

struct MatchArgs {
  const uint2* csr; const int32_t* offsets; const int32_t* blk_start;
  const float* y_rows; int64_t ldY;
  const double* rowstats; const double* colconst;
  float* r; int64_t ldr;
  int T, L, C, margin, n_slabs, warps;
};

constexpr int kMBatch = 4;       // records per cp.async group
constexpr int kMBatches = 4;     // batches in the gather ring: three in flight while one is consumed
constexpr int kMDepth = 16;      // slots of the gather ring (power of two >= kMBatch * kMBatches), kMCols * 4 bytes each
constexpr bool kMF32 = false;    // true: sum up to kMFlush products in float32 before they move to the float64 accumulators
constexpr int kMFlush = 8;       // (cheaper per element, but the result then depends on the order of the match list)
constexpr int kMV = kMCols / 128;   // float4 pieces of a gathered row per lane

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Work item = (block of positions of one task, slab of kMCols columns), one warp each.  The matches of the positions
// [p0 - margin, p1 + margin) of a task are ONE contiguous run of the CSR (rows are positions), so the warp streams that
// run through a ring of asynchronous copies (the latency of a Y row -- HBM at cfg 3 -- is paid once per ring, not once
// per row) and finishes a position whenever the stream crosses a row boundary.
// Arithmetic: exact float32 x float32 products accumulated in float64 (kMF32 = false), the window sums and the
// correlation epilogue in float64.  The kernel is bound by instruction issue -- ~100 instructions per record and 128
// columns, 20 of them the float -> double conversions (ncu: profiles/r02_profile_matches_ncu_summary.txt); summing a few
// products in float32 first (kMF32) is cheaper per element but makes r depend on the order of the match list, which
// the atomics of match_fill_kernel do not fix, so it stays off.
__global__ void __launch_bounds__(128, 2) profile_matches_kernel(const __grid_constant__ MatchArgs a) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (a.offsets[(int64_t)a.T * a.L + 1] != 0) return;           // match list overflow: the caller redoes the group
  if (warp >= a.warps) return;
  const int k = 2 * a.margin + 1;
  // per warp: [k + 1][4 kMV][32] float64 (window ring + the running window sum)
  const size_t per_warp = (size_t)(k + 1) * kMCols * 8;
  double* ring = reinterpret_cast<double*>(s_raw + warp * per_warp) + lane;
  double* runp = ring + (size_t)k * kMCols;
  const int64_t n_items = (int64_t)a.blk_start[0] * a.n_slabs;
  const int64_t w0 = (int64_t)blockIdx.x * a.warps + warp, n_warps = (int64_t)gridDim.x * a.warps;
  const double* cc = a.colconst;
  const double n = cc[0], sy = cc[1], ay = cc[2], sz = cc[3], az = cc[4];
  const bool use_z = cc[5] != 0.0;
  const double inv_k = 1.0 / (double)k;
  const unsigned full = 0xffffffffu;
  for (int64_t item = w0; item < n_items; item += n_warps) {
    // consecutive items = the slabs of one block of positions: the warps that gather the same rows of Y run together
    const int slab = (int)(item % a.n_slabs);
    const int64_t blk = item / a.n_slabs;
    const int row0 = __ldg(a.blk_start + 1 + blk), row1 = __ldg(a.blk_start + 2 + blk);
    const int t = row0 / a.L, p0 = row0 - t * a.L, p1 = p0 + (row1 - row0);
    const int qa = p0 - a.margin > 0 ? p0 - a.margin : 0, qb = p1 + a.margin < a.L ? p1 + a.margin : a.L;
    const float* yb = a.y_rows + (int64_t)slab * kMCols + lane * 4;
    const int32_t* offs = a.offsets + (int64_t)t * a.L;
    // Nothing in the loops below may wait for a dependent global load (a warp issues in order: one exposed L2 round
    // trip per record or per position would serialise the whole stream).  The row boundaries of the item's positions
    // are fetched once (lane i: boundary qa + i, + 32, + 64), the CSR records in coalesced chunks of 32 one chunk
    // ahead, the row statistics one position ahead.
    int32_t bnd[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int qq = qa + lane + 32 * j;
      bnd[j] = __ldg(offs + (qq < qb ? qq : qb));
    }
    auto boundary = [&](int qq) {           // offs[qq] for qa <= qq <= qb (warp uniform)
      const int i = qq - qa;
      const int32_t b0 = __shfl_sync(full, bnd[0], i & 31), b1 = __shfl_sync(full, bnd[1], i & 31),
                    b2 = __shfl_sync(full, bnd[2], i & 31);
      return i < 32 ? b0 : i < 64 ? b1 : b2;
    };
    const int e_begin = boundary(qa), e_end = boundary(qb);
    float part[4 * kMV];                    // float32 partial sums of the current position (at most kMFlush records)
    double s0[4 * kMV];                     // float64 sums of the current position
#pragma unroll
    for (int j = 0; j < 4 * kMV; ++j) { part[j] = 0.0f; s0[j] = 0.0; }
    for (int i = 0; i < (k + 1) * 4 * kMV; ++i) ring[i * 32] = 0.0;
    __syncwarp();
    int slot = 0, n_part = 0;
    int q = p0 - a.margin;                 // the position whose matches are being summed (rows before 0 are empty)
    int next_bnd = q >= 0 ? boundary(q + 1) : e_begin;        // first record that does NOT belong to position q
    // record chunks: chunk c = records e_begin + 32 c + lane
    auto load_chunk = [&](int c) {
      const int e = e_begin + 32 * c + lane;
      return e < e_end ? __ldg(a.csr + e) : make_uint2(0u, 0u);
    };
    uint2 ch_prev = make_uint2(0u, 0u), ch_cur = load_chunk(0), ch_next = load_chunk(1);
    int ch_idx = 0;                        // chunk index of ch_cur = chunk of the issue pointer
    // one batch = kMBatch records starting at e (e - e_begin is a multiple of kMBatch: a batch never straddles a chunk);
    // the rows of Y go straight to registers, the next batch is requested before the current one is summed
    auto load_batch = [&](int e, float4 (&v)[kMBatch][kMV]) {
      if (e < e_end) {
        const int i = e - e_begin;
        if ((i >> 5) != ch_idx) { ch_prev = ch_cur; ch_cur = ch_next; ++ch_idx; ch_next = load_chunk(ch_idx + 1); }
#pragma unroll
        for (int u = 0; u < kMBatch; ++u) {
          const uint32_t seq = __shfl_sync(full, ch_cur.x, (i + u) & 31);     // (padding records: row 0, x = 0)
          const float4* src = reinterpret_cast<const float4*>(yb + (int64_t)seq * a.ldY);
#pragma unroll
          for (int j = 0; j < kMV; ++j) v[u][j] = __ldg(src + 32 * j);
        }
      }
    };
    auto flush_part = [&]() {
#pragma unroll
      for (int j = 0; j < 4 * kMV; ++j) { s0[j] += f2d(part[j]); part[j] = 0.0f; }
      n_part = 0;
    };
    // row statistics of the next row to be written, fetched one position ahead
    double st_sx = 0.0, st_sxx = 0.0, st_sxz = 0.0;
    auto fetch_stats = [&](int p) {
      if (p >= p0 && p < p1) {
        const double* rsx = a.rowstats + ((int64_t)t * a.L + p) * 3;
        st_sx = __ldg(rsx); st_sxx = __ldg(rsx + 1); st_sxz = __ldg(rsx + 2);
      }
    };
    fetch_stats(p0);
    // position q is complete: it enters the window, the oldest row leaves, and row q - margin can be written
    auto finish_position = [&]() {
      if (n_part) flush_part();
      double* rs = ring + (size_t)slot * kMCols;
      if (++slot == k) slot = 0;
      const int p = q - a.margin;
      ++q;
      next_bnd = q < 0 ? e_begin : q < qb ? boundary(q + 1) : e_end;
      const bool emit = p >= p0;
      // core.py:79-91 zeroes the first / last `margin` rows of the blurred matrix
      const bool live = p >= a.margin && p < a.L - a.margin;
      const int64_t row = (int64_t)t * a.L + p;
      double inv_xy = 0.0, sxsy = 0.0, r_xz = 0.0, row_den = 1.0;
      if (emit) {
        const double sx = st_sx, sxx = st_sxx, sxz = st_sxz;
        fetch_stats(p + 1);
        const double ax = n * sxx - sx * sx;
        inv_xy = 1.0 / sqrt(ax * ay);
        sxsy = sx * sy;
        if (use_z) {
          r_xz = (n * sxz - sx * sz) / sqrt(ax * az);
          row_den = 1.0 / sqrt(1.0 - r_xz * r_xz);
        }
      }
#pragma unroll
      for (int j = 0; j < kMV; ++j) {
        float o[4];
        const int c0 = slab * kMCols + j * 128 + lane * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = (4 * j + u) * 32;
          const double old = rs[idx];
          const double w = runp[idx] + (s0[4 * j + u] - old);       // the window sum after position q - 1 entered
          rs[idx] = s0[4 * j + u];
          runp[idx] = w;
          s0[4 * j + u] = 0.0;
          float out = 0.0f;
          const int col = c0 + u;
          if (emit && col < a.C) {
            const double sxy = live ? w * inv_k : 0.0;
            double rr = (n * sxy - sxsy) * inv_xy;
            if (use_z) {
              const double2 cz = __ldg(reinterpret_cast<const double2*>(cc + 8 + 2 * col));
              rr = (rr - r_xz * cz.x) * row_den * cz.y;
            }
            out = (float)rr;
            if (out != out) out = 0.0f;
            else if (isinf(out)) out = out > 0 ? FLT_MAX : -FLT_MAX;
          }
          o[u] = out;
        }
        if (emit && c0 < a.ldr) *reinterpret_cast<float4*>(a.r + row * a.ldr + c0) = make_float4(o[0], o[1], o[2], o[3]);
      }
    };
    auto sum_batch = [&](int e, const float4 (&v)[kMBatch][kMV]) {
      // the load pointer is less than 32 records ahead: the batch is in the current chunk or the previous one
      const int i = e - e_begin;
      const bool in_cur = (i >> 5) == ch_idx;
#pragma unroll
      for (int u = 0; u < kMBatch; ++u) {
        const uint32_t xc = __shfl_sync(full, ch_cur.y, (i + u) & 31), xp = __shfl_sync(full, ch_prev.y, (i + u) & 31);
        if (e + u < e_end) {
          while (e + u >= next_bnd) finish_position();
          const double xd = f2d(__uint_as_float(in_cur ? xc : xp));
#pragma unroll
          for (int j = 0; j < kMV; ++j) {
            s0[4 * j + 0] = fma(xd, f2d(v[u][j].x), s0[4 * j + 0]);
            s0[4 * j + 1] = fma(xd, f2d(v[u][j].y), s0[4 * j + 1]);
            s0[4 * j + 2] = fma(xd, f2d(v[u][j].z), s0[4 * j + 2]);
            s0[4 * j + 3] = fma(xd, f2d(v[u][j].w), s0[4 * j + 3]);
          }
        }
      }
    };
    float4 va[kMBatch][kMV], vb[kMBatch][kMV];
    load_batch(e_begin, va);
#pragma unroll 1
    for (int e = e_begin; e < e_end; e += 2 * kMBatch) {
      load_batch(e + kMBatch, vb);
      sum_batch(e, va);
      if (e + kMBatch < e_end) {
        load_batch(e + 2 * kMBatch, va);
        sum_batch(e + kMBatch, vb);
      }
    }
    while (q < p1 + a.margin) finish_position();
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

int64_t mepp_profile_matches_work_bytes(int64_t m_rows, int64_t hits_cap) {
  return 2 * (((m_rows + 2) * 4 + 255) / 256 * 256) + (m_rows * 4 + 255) / 256 * 256 + hits_cap * 8;
}

int mepp_profile_matches(const void* hits_dev, const unsigned long long* hit_count_dev, int64_t hits_cap,
                         const int32_t* count_dev, int T, int L, int margin, const float* y_rows_dev,
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
  int32_t* blk_start = reinterpret_cast<int32_t*>(w);
  w += ((m_rows + 2) * 4 + 255) / 256 * 256;
  uint2* csr = reinterpret_cast<uint2*>(w);
  match_offsets_kernel<<<1, kScanThreads, 0, st>>>(count_dev, m_rows, hit_count_dev, hits_cap, offsets, cursor, L,
                                                   blk_start);
  match_fill_kernel<<<148 * 8, 256, 0, st>>>(reinterpret_cast<const HitRec4*>(hits_dev), hit_count_dev, L, m_rows, offsets,
                                             cursor, csr);
  MatchArgs a;
  a.csr = csr; a.offsets = offsets; a.blk_start = blk_start;
  a.y_rows = y_rows_dev; a.ldY = mepp_y_rows_ld(C);
  a.rowstats = rowstats_dev; a.colconst = colconst_dev; a.r = r_dev; a.ldr = ldr;
  a.T = T; a.L = L; a.C = C; a.margin = margin; a.n_slabs = (int)(a.ldY / kMCols);
  // per warp: window ring + running sum ((k + 1) rows x kMCols float64) and the gather ring; as many warps per CTA
  // (<= 4) as fit ~106 KB, two CTAs per SM
  const int k = 2 * margin + 1;
  const size_t per_warp = (size_t)(k + 1) * kMCols * 8;
  int warps = (int)(106 * 1024 / per_warp);
  warps = warps > 4 ? 4 : warps < 1 ? 1 : warps;
  a.warps = warps;
  const size_t smem = (size_t)warps * per_warp;
  MEPP_CUDA_CHECK(cudaFuncSetAttribute(profile_matches_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  profile_matches_kernel<<<148 * 2, 128, smem, st>>>(a);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"

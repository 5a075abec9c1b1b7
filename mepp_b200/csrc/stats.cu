// Correlation epilogue and K3 permutation statistics.
//
// Replaces (reference npdeloss/mepp): the r_ab / partial-correlation formulas of
// mepp/core.py:182-193,255-265 plus np.nan_to_num (core.py:602), and the
// per-position percentiles / permutation p-values / pooled ("semi") null /
// integral statistic of mepp/core.py:269-376.
#include "common.cuh"
#include <cfloat>

namespace mepp {

// ------------------------------------------------------------------ correlation
struct CorrArgs {
  const float* S; int64_t ldS;
  const double* rowstats;   // [m_rows][3]  sum x, sum x^2, sum x*zhat   (x unscaled)
  const double* col_syz;    // [C]
  const double* ysum;       // {sum y, sum y^2}
  double sz, szz, n;
  int64_t m_rows; int C; int use_z;
  float* r; int64_t ldr;
};

__device__ __forceinline__ float nan_to_num_f32(double v) {
  // np.nan_to_num on the float32 array: NaN -> 0, +-inf -> +-FLT_MAX (core.py:602)
  float f = (float)v;
  if (f != f) return 0.0f;
  if (isinf(f)) return f > 0 ? FLT_MAX : -FLT_MAX;
  return f;
}

__global__ void correlation_kernel(const CorrArgs a) {
  const int64_t row = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const double n = a.n;
  const double sx = a.rowstats[row * 3 + 0], sxx = a.rowstats[row * 3 + 1], sxz = a.rowstats[row * 3 + 2];
  const double sy = a.ysum[0], syy = a.ysum[1];
  const double sxy = (double)a.S[row * a.ldS + c] * (1.0 / (double)MEPP_X_SCALE);
  const double ax = n * sxx - sx * sx;
  const double ay = n * syy - sy * sy;
  const double r_xy = (n * sxy - sx * sy) / sqrt(ax * ay);
  double r = r_xy;
  if (a.use_z) {
    const double az = n * a.szz - a.sz * a.sz;
    const double r_xz = (n * sxz - sx * a.sz) / sqrt(ax * az);
    const double r_yz = (n * a.col_syz[c] - sy * a.sz) / sqrt(ay * az);
    r = (r_xy - r_xz * r_yz) / sqrt((1.0 - r_xz * r_xz) * (1.0 - r_yz * r_yz));
  }
  a.r[row * a.ldr + c] = nan_to_num_f32(r);
}

// Per-dataset constants of the correlation epilogue fused into the GEMM (gemm.cu): header
// {n, sy, ay, sz, az, use_z, -, -} then per column {r_yz[c], 1 / sqrt(1 - r_yz[c]^2)}.
__global__ void col_constants_kernel(const double* __restrict__ col_syz, const double* __restrict__ ysum,
                                     double sz, double szz, double n, int C, int use_z, double* __restrict__ cc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const double sy = ysum[0], syy = ysum[1];
  const double ay = n * syy - sy * sy, az = n * szz - sz * sz;
  if (c == 0) { cc[0] = n; cc[1] = sy; cc[2] = ay; cc[3] = sz; cc[4] = az; cc[5] = (double)use_z; cc[6] = 0.0; cc[7] = 0.0; }
  if (c >= C) return;
  double r_yz = 0.0, cden = 1.0;
  if (use_z) {
    r_yz = (n * col_syz[c] - sy * sz) / sqrt(ay * az);
    cden = 1.0 / sqrt(1.0 - r_yz * r_yz);
  }
  cc[8 + 2 * c] = r_yz;
  cc[9 + 2 * c] = cden;
}

// ------------------------------------------------------------------ K3 helpers
// Monotone float -> uint key (total order; -0 < +0).
__device__ __forceinline__ uint32_t f2key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// In-place ascending bitonic sort of n_pow2 keys in shared memory by the whole block.
__device__ void bitonic_sort_keys(uint32_t* s, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint32_t a = s[i], b = s[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s[i] = b; s[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// Per (task, position): order statistics of the P permuted r values, the count of
// |r_p| >= |r_0|, and the sorted row (as keys) for the pooled quantiles.
__global__ void row_stats_kernel(const float* __restrict__ r, int64_t ldr, int L, int C, int P_pow2, int k_lo, int k_hi,
                                 float* __restrict__ full_os, int32_t* __restrict__ full_cnt,
                                 uint32_t* __restrict__ sorted_keys) {
  extern __shared__ uint32_t s_keys[];
  __shared__ int s_cnt;
  const int P = C - 1;
  const int64_t row = (int64_t)blockIdx.y * L + blockIdx.x;
  const float* rr = r + row * ldr;
  const float t0 = fabsf(rr[0]);
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int cnt = 0;
  for (int i = threadIdx.x; i < P_pow2; i += blockDim.x) {
    uint32_t key = 0xffffffffu;  // pads sort to the end
    if (i < P) {
      const float v = rr[1 + i];
      key = f2key(v);
      cnt += fabsf(v) >= t0 ? 1 : 0;
    }
    s_keys[i] = key;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  bitonic_sort_keys(s_keys, P_pow2);
  for (int i = threadIdx.x; i < P; i += blockDim.x) sorted_keys[row * P + i] = s_keys[i];
  if (threadIdx.x == 0) {
    const int k1 = min(k_lo + 1, P - 1), k3 = min(k_hi + 1, P - 1);
    float* o = full_os + row * 4;
    o[0] = key2f(s_keys[k_lo]); o[1] = key2f(s_keys[k1]);
    o[2] = key2f(s_keys[k_hi]); o[3] = key2f(s_keys[k3]);
    full_cnt[row] = s_cnt;
  }
}

// One bitonic stage (block size K, partner distance J) on a row held as Q keys per lane
// (element e = lane + 32 q), then the next stage; compile-time recursion keeps every register
// index static.
template <int Q, int K, int J>
__device__ __forceinline__ void bitonic_stage(uint32_t (&key)[Q], const int lane) {
  if constexpr (J >= 32) {
    constexpr int jq = J >> 5;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int pq = q ^ jq;
      if (pq > q) {
        const bool up = ((32 * q) & K) == 0;     // K >= 64 here: direction depends on q only
        const uint32_t a = key[q], b = key[pq];
        const uint32_t lo = min(a, b), hi = max(a, b);
        key[q] = up ? lo : hi;
        key[pq] = up ? hi : lo;
      }
    }
  } else {
    const bool lower = (lane & J) == 0;          // this lane holds the smaller index of the pair
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const uint32_t mine = key[q];
      const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, J);
      const bool up = ((lane + 32 * q) & K) == 0;
      key[q] = (lower == up) ? min(mine, other) : max(mine, other);
    }
  }
  if constexpr (J > 1) bitonic_stage<Q, K, J / 2>(key, lane);
  else if constexpr (K < 32 * Q) bitonic_stage<Q, 2 * K, K>(key, lane);
}

// Warp-per-row variant for P <= 1024: the whole row lives in registers (Q = P_pow2 / 32 keys per
// lane, element e = lane + 32 q), bitonic stages with partner distance >= 32 are register-to-register
// compare-exchanges, smaller distances are warp shuffles; no shared memory, no block barriers.
template <int Q>
__global__ void __launch_bounds__(256) row_stats_warp_kernel(const float* __restrict__ r, int64_t ldr, int64_t n_rows, int C,
                                                             int k_lo, int k_hi, float* __restrict__ full_os,
                                                             int32_t* __restrict__ full_cnt,
                                                             uint32_t* __restrict__ sorted_keys) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int P = C - 1;
  const float* rr = r + row * ldr;
  const float t0 = fabsf(rr[0]);
  uint32_t key[Q];
  int cnt = 0;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int e = lane + 32 * q;
    uint32_t kk = 0xffffffffu;  // pads sort to the end
    if (e < P) {
      const float v = rr[1 + e];
      kk = f2key(v);
      cnt += fabsf(v) >= t0 ? 1 : 0;
    }
    key[q] = kk;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(full, cnt, o);
  bitonic_stage<Q, 2, 1>(key, lane);
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int e = lane + 32 * q;
    if (e < P) sorted_keys[row * P + e] = key[q];
  }
  const int ranks[4] = {k_lo, min(k_lo + 1, P - 1), k_hi, min(k_hi + 1, P - 1)};
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int e = ranks[w];
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < Q; ++q)
      if (q == (e >> 5)) v = key[q];
    v = __shfl_sync(full, v, e & 31);
    if (lane == 0) full_os[row * 4 + w] = key2f(v);
  }
  if (lane == 0) full_cnt[row] = cnt;
}

// Per task: sorted thresholds |r[l,0]| (ascending) for the pooled p-values.
__global__ void thresholds_kernel(const float* __restrict__ r, int64_t ldr, int L, int C, int L_pow2,
                                  float* __restrict__ ts_sorted) {
  extern __shared__ uint32_t s_keys[];
  const int t = blockIdx.x;
  for (int i = threadIdx.x; i < L_pow2; i += blockDim.x)
    s_keys[i] = i < L ? f2key(fabsf(r[((int64_t)t * L + i) * ldr])) : 0xffffffffu;
  __syncthreads();
  bitonic_sort_keys(s_keys, L_pow2);
  for (int i = threadIdx.x; i < L; i += blockDim.x) ts_sorted[(int64_t)t * L + i] = key2f(s_keys[i]);
}

// Histogram of the pooled |r_perm| values against the task's sorted thresholds:
// hist[idx] counts values with exactly idx thresholds <= |v|.
__global__ void semi_hist_kernel(const float* __restrict__ r, int64_t ldr, int L, int C, const float* __restrict__ ts_sorted,
                                 unsigned long long* __restrict__ hist) {
  extern __shared__ uint32_t s_raw[];
  float* s_ts = reinterpret_cast<float*>(s_raw);        // [L]
  uint32_t* s_hist = s_raw + L;                         // [L + 1]
  const int t = blockIdx.y;
  for (int i = threadIdx.x; i < L; i += blockDim.x) s_ts[i] = ts_sorted[(int64_t)t * L + i];
  for (int i = threadIdx.x; i <= L; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const int rows_per = (L + gridDim.x - 1) / gridDim.x;
  const int l0 = blockIdx.x * rows_per, l1 = min(L, l0 + rows_per);
  const int P = C - 1;
  for (int l = l0; l < l1; ++l) {
    const float* rr = r + ((int64_t)t * L + l) * ldr + 1;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
      const float a = fabsf(rr[i]);
      int lo = 0, hi = L;  // upper_bound: number of thresholds <= a
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_ts[mid] <= a) lo = mid + 1; else hi = mid;
      }
      atomicAdd(&s_hist[lo], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= L; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&hist[(int64_t)t * (L + 1) + i], (unsigned long long)s_hist[i]);
}

// Per task: semi_cnt[l] = #{pooled |v| >= |r[l,0]|} from the histogram (suffix sums).
__global__ void semi_count_kernel(const float* __restrict__ r, int64_t ldr, int L, int C, const float* __restrict__ ts_sorted,
                                  unsigned long long* __restrict__ hist, long long* __restrict__ semi_cnt) {
  const int t = blockIdx.x;
  unsigned long long* h = hist + (int64_t)t * (L + 1);
  // in-place suffix sum by one thread (L is small); h[j] becomes sum_{idx >= j} hist[idx]
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int j = L; j >= 0; --j) { run += h[j]; h[j] = run; }
  }
  __syncthreads();
  const float* ts = ts_sorted + (int64_t)t * L;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float a = fabsf(r[((int64_t)t * L + l) * ldr]);
    int lo = 0, hi = L;  // lower_bound: first j with ts[j] >= a  (ts[j] == a)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (ts[mid] < a) lo = mid + 1; else hi = mid;
    }
    // values with at least lo+1 thresholds <= |v| are exactly those with |v| >= ts[lo] = a
    semi_cnt[(int64_t)t * L + l] = (long long)h[lo + 1];
  }
}

// Pooled order statistics by bisection on the key space over the sorted rows.
// grid = (4 targets, T); smallest key K with #{keys <= K} >= rank+1.
__global__ void semi_select_kernel(const uint32_t* __restrict__ sorted_keys, int L, int P,
                                   long long ks_lo, long long ks_hi, float* __restrict__ semi_os) {
  __shared__ unsigned long long s_red[32];
  __shared__ unsigned long long s_total;
  const int t = blockIdx.y;
  const long long M = (long long)L * P;
  long long rank = blockIdx.x == 0 ? ks_lo : blockIdx.x == 1 ? ks_lo + 1 : blockIdx.x == 2 ? ks_hi : ks_hi + 1;
  if (rank > M - 1) rank = M - 1;
  const uint32_t* base = sorted_keys + (int64_t)t * L * P;
  uint32_t klo = 0u, khi = 0xffffffffu;
  while (klo < khi) {
    const uint32_t mid = klo + ((khi - klo) >> 1);
    unsigned long long cnt = 0;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      const uint32_t* row = base + (int64_t)l * P;
      int lo = 0, hi = P;  // upper_bound(row, mid)
      while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (row[m] <= mid) lo = m + 1; else hi = m;
      }
      cnt += (unsigned long long)lo;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w];
      s_total = tot;
    }
    __syncthreads();
    if ((long long)s_total >= rank + 1) khi = mid; else klo = mid + 1;
    __syncthreads();
  }
  if (threadIdx.x == 0) semi_os[t * 4 + blockIdx.x] = key2f(klo);
}

// integral[t][c] = trapezoid over positions of |r[:, c]| (np.trapz, dx = 1).
__global__ void integral_kernel(const float* __restrict__ r, int64_t ldr, int L, int C, double* __restrict__ integral) {
  const int t = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* rr = r + (int64_t)t * L * ldr + c;
  double acc = 0.0;
  float prev = fabsf(rr[0]);
  for (int l = 1; l < L; ++l) {
    const float cur = fabsf(rr[(int64_t)l * ldr]);
    acc += 0.5 * ((double)prev + (double)cur);
    prev = cur;
  }
  integral[(int64_t)t * C + c] = acc;
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace mepp

extern "C" {

int mepp_correlation(const float* S_dev, int64_t ldS, const double* rowstats_dev, const double* col_syz_dev,
                     const double* ysum_dev, const double* zsum_host2, int64_t N, int64_t m_rows, int C,
                     int use_z, float* r_dev, int64_t ldr, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(S_dev && rowstats_dev && ysum_dev && r_dev, "correlation: null pointer");
  MEPP_REQUIRE(!use_z || (col_syz_dev && zsum_host2), "correlation: z statistics missing");
  MEPP_REQUIRE(m_rows > 0 && m_rows < (1LL << 31) && C > 0, "correlation: bad shape");
  MEPP_REQUIRE(mepp_device_ok(), "correlation: no sm_100 CUDA device (no CPU fallback)");
  CorrArgs a;
  a.S = S_dev; a.ldS = ldS; a.rowstats = rowstats_dev; a.col_syz = col_syz_dev; a.ysum = ysum_dev;
  a.sz = use_z ? zsum_host2[0] : 0.0; a.szz = use_z ? zsum_host2[1] : 0.0; a.n = (double)N;
  a.m_rows = m_rows; a.C = C; a.use_z = use_z; a.r = r_dev; a.ldr = ldr;
  MEPP_REQUIRE(ldr >= C, "correlation: ldr must be >= C");
  // rows on grid.y (<= 65535): split large row counts over several launches
  const int64_t max_rows = 65535;
  for (int64_t r0 = 0; r0 < m_rows; r0 += max_rows) {
    CorrArgs b = a;
    const int64_t nr = m_rows - r0 < max_rows ? m_rows - r0 : max_rows;
    b.S = a.S + r0 * ldS; b.rowstats = a.rowstats + r0 * 3; b.r = a.r + r0 * ldr; b.m_rows = nr;
    dim3 grid((unsigned)((C + 255) / 256), (unsigned)nr);
    correlation_kernel<<<grid, 256, 0, as_stream(stream)>>>(b);
  }
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int mepp_col_constants(const double* col_syz_dev, const double* ysum_dev, const double* zsum_host2, int64_t N, int C,
                       int use_z, double* colconst_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(ysum_dev && colconst_dev && C > 0, "col_constants: bad arguments");
  MEPP_REQUIRE(!use_z || (col_syz_dev && zsum_host2), "col_constants: z statistics missing");
  MEPP_REQUIRE(mepp_device_ok(), "col_constants: no sm_100 CUDA device (no CPU fallback)");
  col_constants_kernel<<<(unsigned)((C + 255) / 256), 256, 0, as_stream(stream)>>>(
      col_syz_dev, ysum_dev, use_z ? zsum_host2[0] : 0.0, use_z ? zsum_host2[1] : 0.0, (double)N, C, use_z, colconst_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int64_t mepp_perm_stats_work_bytes(int T, int L, int C) {
  const int64_t P = C - 1;
  int64_t keys = (int64_t)T * L * (P > 0 ? P : 1) * 4;
  int64_t ts = (int64_t)T * L * 4;
  int64_t hist = (int64_t)T * (L + 1) * 8;
  return ((keys + 255) / 256 + (ts + 255) / 256 + (hist + 255) / 256) * 256;
}

int mepp_perm_stats(const float* r_dev, int64_t ldr, int T, int L, int C, int k_lo, int k_hi, int64_t ks_lo, int64_t ks_hi,
                    float* full_os_dev, int32_t* full_cnt_dev, float* semi_os_dev, long long* semi_cnt_dev,
                    double* integral_dev, void* work_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(r_dev && full_os_dev && full_cnt_dev && semi_os_dev && semi_cnt_dev && integral_dev && work_dev,
               "perm_stats: null pointer");
  const int P = C - 1;
  MEPP_REQUIRE(T > 0 && L > 0 && P >= 1, "perm_stats: needs at least one permutation");
  MEPP_REQUIRE(ldr >= C, "perm_stats: ldr must be >= C");
  MEPP_REQUIRE(T <= 65535 && L <= 65535, "perm_stats: T and L must be <= 65535");
  MEPP_REQUIRE(k_lo >= 0 && k_lo < P && k_hi >= 0 && k_hi < P, "perm_stats: percentile index out of range");
  MEPP_REQUIRE(ks_lo >= 0 && ks_hi < (int64_t)L * P, "perm_stats: pooled percentile index out of range");
  MEPP_REQUIRE(mepp_device_ok(), "perm_stats: no sm_100 CUDA device (no CPU fallback)");
  const int P_pow2 = next_pow2(P), L_pow2 = next_pow2(L);
  MEPP_REQUIRE(P_pow2 <= 32768, "perm_stats: at most 32768 permutations per call");
  MEPP_REQUIRE(L_pow2 <= 16384, "perm_stats: at most 16384 positions");
  cudaStream_t st = as_stream(stream);
  uint8_t* w = reinterpret_cast<uint8_t*>(work_dev);
  const int64_t keys_b = ((int64_t)T * L * P * 4 + 255) / 256 * 256;
  const int64_t ts_b = ((int64_t)T * L * 4 + 255) / 256 * 256;
  uint32_t* sorted_keys = reinterpret_cast<uint32_t*>(w);
  float* ts_sorted = reinterpret_cast<float*>(w + keys_b);
  unsigned long long* hist = reinterpret_cast<unsigned long long*>(w + keys_b + ts_b);
  MEPP_CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * (size_t)T * (L + 1), st));

  if (P_pow2 <= 1024) {
    // rows in registers, one warp per row
    const int64_t n_rows = (int64_t)T * L;
    const unsigned blocks = (unsigned)((n_rows + 7) / 8);
    const int Q = P_pow2 <= 32 ? 1 : P_pow2 / 32;
#define MEPP_ROW_WARP(QQ) row_stats_warp_kernel<QQ><<<blocks, 256, 0, st>>>(r_dev, ldr, n_rows, C, k_lo, k_hi, full_os_dev, \
                                                                           full_cnt_dev, sorted_keys)
    switch (Q) {
      case 1: MEPP_ROW_WARP(1); break;
      case 2: MEPP_ROW_WARP(2); break;
      case 4: MEPP_ROW_WARP(4); break;
      case 8: MEPP_ROW_WARP(8); break;
      case 16: MEPP_ROW_WARP(16); break;
      default: MEPP_ROW_WARP(32); break;
    }
#undef MEPP_ROW_WARP
  } else {
    const size_t row_smem = sizeof(uint32_t) * (size_t)P_pow2;
    MEPP_CUDA_CHECK(cudaFuncSetAttribute(row_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem));
    row_stats_kernel<<<dim3((unsigned)L, (unsigned)T), 256, row_smem, st>>>(
        r_dev, ldr, L, C, P_pow2, k_lo, k_hi, full_os_dev, full_cnt_dev, sorted_keys);
  }

  const size_t thr_smem = sizeof(uint32_t) * (size_t)L_pow2;
  MEPP_CUDA_CHECK(cudaFuncSetAttribute(thresholds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)thr_smem));
  thresholds_kernel<<<(unsigned)T, 256, thr_smem, st>>>(r_dev, ldr, L, C, L_pow2, ts_sorted);

  const size_t hist_smem = sizeof(uint32_t) * (size_t)(2 * L + 1);
  MEPP_CUDA_CHECK(cudaFuncSetAttribute(semi_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem));
  int chunks = (int)(((int64_t)L * P + 65535) / 65536);
  if (chunks < 1) chunks = 1;
  if (chunks > L) chunks = L;
  semi_hist_kernel<<<dim3((unsigned)chunks, (unsigned)T), 256, hist_smem, st>>>(r_dev, ldr, L, C, ts_sorted, hist);
  semi_count_kernel<<<(unsigned)T, 256, 0, st>>>(r_dev, ldr, L, C, ts_sorted, hist, semi_cnt_dev);
  semi_select_kernel<<<dim3(4, (unsigned)T), 256, 0, st>>>(sorted_keys, L, P, ks_lo, ks_hi, semi_os_dev);
  integral_kernel<<<dim3((unsigned)((C + 127) / 128), (unsigned)T), 128, 0, st>>>(r_dev, ldr, L, C, integral_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"

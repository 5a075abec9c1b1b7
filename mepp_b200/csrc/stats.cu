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
                                     double sz, double szz, const double* __restrict__ zsum_dev, double n, int C,
                                     int use_z, double* __restrict__ cc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (zsum_dev) { sz = zsum_dev[0]; szz = zsum_dev[1]; }
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
constexpr int kRadixBins = 2048;   // 11-bit digits of the pooled radix select
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
                                 uint32_t* __restrict__ hist1) {
  extern __shared__ uint32_t s_keys[];
  __shared__ int s_cnt;
  __shared__ uint32_t s_h[kRadixBins];
  for (int i = threadIdx.x; i < kRadixBins; i += blockDim.x) s_h[i] = 0u;
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
  // first digit of the pooled radix select (see radix_pass_kernel)
  for (int i = threadIdx.x; i < P; i += blockDim.x) atomicAdd(&s_h[s_keys[i] >> 21], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < kRadixBins; i += blockDim.x)
    if (s_h[i]) atomicAdd(&hist1[(int64_t)blockIdx.y * kRadixBins + i], s_h[i]);
  bitonic_sort_keys(s_keys, P_pow2);
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

// Warp-per-row kernel for P <= 1024.  A row is staged in shared memory (cp.async, the next row of the warp in
// flight while this one is processed) and turned into order-preserving keys in place.  Only the few smallest /
// largest values matter (the interval ranks sit in the tails), so instead of sorting the row the warp bounds each
// tail with a pivot -- the m-th smallest of the 64 minima of the half-lane groups is >= the m-th smallest element
// -- compacts the <= 64 values beyond the pivot and sorts those in registers.  Rows with more than 64 candidates
// (heavy ties) and intervals that need more than 40 values per tail take a bitonic sort of the staged row.
// grid = (row chunks, T): a CTA also accumulates the first digit of the pooled radix select for its task.
constexpr int kRowWarps = 4;

__device__ __forceinline__ void warp_sort_smem(uint32_t* s, int n_pow2, int lane) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n_pow2; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint32_t a = s[i], b = s[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s[i] = b; s[ixj] = a; }
        }
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(kRowWarps * 32) row_select_kernel(const float* __restrict__ r, int64_t ldr, int L, int C,
                                                                    int P_pow2, int rows_per_cta, int k_lo, int k_hi,
                                                                    int fast, float* __restrict__ full_os,
                                                                    int32_t* __restrict__ full_cnt,
                                                                    uint32_t* __restrict__ hist1) {
  const unsigned full = 0xffffffffu;
  extern __shared__ __align__(16) uint32_t s_dyn[];     // [warps][2][P_pow2] row buffers
  __shared__ uint32_t s_h[kRadixBins];
  __shared__ uint32_t s_cand[kRowWarps][2][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.y;
  const int P = C - 1;
  const int Q = P_pow2 >> 5;                            // >= 1
  for (int i = threadIdx.x; i < kRadixBins; i += blockDim.x) s_h[i] = 0u;
  __syncthreads();
  const int l0 = blockIdx.x * rows_per_cta, l1 = min(L, l0 + rows_per_cta);
  const int r1 = min(k_lo + 1, P - 1), r3 = min(k_hi + 1, P - 1);
  const int m_lo = r1 + 1, m_hi = P - k_hi;             // values needed from either tail
  const unsigned lt = (1u << lane) - 1u;
  uint32_t* buf0 = s_dyn + (size_t)warp * 2 * P_pow2;
  auto prefetch = [&](int l, uint32_t* dst) {
    const float* rr = r + ((int64_t)t * L + l) * ldr + 1;
    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
    for (int e = lane; e < P; e += 32)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 4u * (uint32_t)e), "l"(rr + e) : "memory");
  };
  int cur = 0;
  if (l0 + warp < l1) prefetch(l0 + warp, buf0);
#pragma unroll 1
  for (int l = l0 + warp; l < l1; l += kRowWarps) {
    const int64_t row = (int64_t)t * L + l;
    uint32_t* key = buf0 + cur * P_pow2;
    const float t0 = fabsf(r[row * ldr]);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    if (l + kRowWarps < l1) prefetch(l + kRowWarps, buf0 + (cur ^ 1) * P_pow2);
    cur ^= 1;
    // keys in place, count of |v| >= |r0|, first radix digit, minima / maxima of the two half-lane groups
    int cnt = 0;
    uint32_t gmin0 = 0xffffffffu, gmin1 = 0xffffffffu, gmax0 = 0u, gmax1 = 0u;
    const int half = Q > 1 ? Q >> 1 : 1;
#pragma unroll 4
    for (int q = 0; q < Q; ++q) {
      const int e = lane + 32 * q;
      uint32_t kk = 0xffffffffu;                        // pads sort to the end
      if (e < P) {
        const float v = __uint_as_float(key[e]);
        kk = f2key(v);
        cnt += fabsf(v) >= t0 ? 1 : 0;
        atomicAdd(&s_h[kk >> 21], 1u);
        if (q >= half) { gmin1 = min(gmin1, kk); gmax1 = max(gmax1, kk); }
        else { gmin0 = min(gmin0, kk); gmax0 = max(gmax0, kk); }
      }
      key[e] = kk;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(full, cnt, o);
    if (lane == 0) full_cnt[row] = cnt;
    __syncwarp();
    bool done = false;
    if (fast) {
      uint32_t gmin[2] = {gmin0, gmin1}, gmax[2] = {gmax0, gmax1};
      bitonic_stage<2, 2, 1>(gmin, lane);
      bitonic_stage<2, 2, 1>(gmax, lane);
      const uint32_t row_min = __shfl_sync(full, gmin[0], 0), row_max = __shfl_sync(full, gmax[1], 31);
      if (row_min == row_max) {                         // a constant row (e.g. r = 0 where nothing matched)
        if (lane == 0) {
          const float v = key2f(row_min);
          *reinterpret_cast<float4*>(full_os + row * 4) = make_float4(v, v, v, v);
        }
        done = true;
      } else {
        const int e_lo = m_lo - 1, e_hi = 64 - m_hi;
        const uint32_t pivot_lo = __shfl_sync(full, (e_lo >> 5) ? gmin[1] : gmin[0], e_lo & 31);
        const uint32_t pivot_hi = __shfl_sync(full, (e_hi >> 5) ? gmax[1] : gmax[0], e_hi & 31);
        int n_lo = 0, n_hi = 0;
        uint32_t* c_lo = s_cand[warp][0];
        uint32_t* c_hi = s_cand[warp][1];
#pragma unroll 4
        for (int q = 0; q < Q; ++q) {
          const int e = lane + 32 * q;
          const uint32_t kk = key[e];
          const bool real = e < P;
          const bool is_lo = kk <= pivot_lo && real;
          const bool is_hi = kk >= pivot_hi && real;
          const unsigned b_lo = __ballot_sync(full, is_lo), b_hi = __ballot_sync(full, is_hi);
          if (b_lo) {
            const int pos = n_lo + __popc(b_lo & lt);
            if (is_lo && pos < 64) c_lo[pos] = kk;
            n_lo += __popc(b_lo);
          }
          if (b_hi) {
            const int pos = n_hi + __popc(b_hi & lt);
            if (is_hi && pos < 64) c_hi[pos] = kk;
            n_hi += __popc(b_hi);
          }
        }
        if (n_lo <= 64 && n_hi <= 64) {
          __syncwarp();
          uint32_t a[2], b[2];
          a[0] = lane < n_lo ? c_lo[lane] : 0xffffffffu;  a[1] = lane + 32 < n_lo ? c_lo[lane + 32] : 0xffffffffu;
          b[0] = lane < n_hi ? c_hi[lane] : 0u;           b[1] = lane + 32 < n_hi ? c_hi[lane + 32] : 0u;
          bitonic_stage<2, 2, 1>(a, lane);
          bitonic_stage<2, 2, 1>(b, lane);
          // a: the n_lo smallest values ascending; b: (64 - n_hi) zero pads, then the n_hi largest ascending
          const int i2 = k_hi - P + 64, i3 = r3 - P + 64;
          const uint32_t v0 = __shfl_sync(full, (k_lo >> 5) ? a[1] : a[0], k_lo & 31);
          const uint32_t v1 = __shfl_sync(full, (r1 >> 5) ? a[1] : a[0], r1 & 31);
          const uint32_t v2 = __shfl_sync(full, (i2 >> 5) ? b[1] : b[0], i2 & 31);
          const uint32_t v3 = __shfl_sync(full, (i3 >> 5) ? b[1] : b[0], i3 & 31);
          if (lane == 0)
            *reinterpret_cast<float4*>(full_os + row * 4) = make_float4(key2f(v0), key2f(v1), key2f(v2), key2f(v3));
          done = true;
        }
      }
    }
    if (!done) {
      warp_sort_smem(key, P_pow2, lane);
      if (lane == 0)
        *reinterpret_cast<float4*>(full_os + row * 4) =
            make_float4(key2f(key[k_lo]), key2f(key[r1]), key2f(key[k_hi]), key2f(key[r3]));
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRadixBins; i += blockDim.x)
    if (s_h[i]) atomicAdd(&hist1[(int64_t)t * kRadixBins + i], s_h[i]);
}

// ---- pooled order statistics: exact radix select (11 + 11 + 10 bits of the order-preserving key) over the L * P
// permuted values of a task.  The first digit is histogrammed by the row kernels above; radix_pick_kernel turns a
// histogram into (bin holding the rank, rank inside the bin); radix_pass_kernel histograms the next digit of the
// values whose prefix matches; four independent targets per task (ranks ks_lo, ks_lo+1, ks_hi, ks_hi+1).
struct RadixState { uint32_t prefix[4]; unsigned long long rank[4]; };

// grid = T, block = 128 (one warp per target).  digit_bits: width of the digit just histogrammed; shift: its position.
__global__ void radix_pick_kernel(const uint32_t* __restrict__ hist, int hists_per_task, int n_bins, int shift,
                                  RadixState* __restrict__ state, long long ks_lo, long long ks_hi, long long M, int first,
                                  float* __restrict__ semi_os) {
  const int t = blockIdx.x, j = threadIdx.x >> 5, lane = threadIdx.x & 31;
  RadixState* s = state + t;
  unsigned long long rank;
  uint32_t prefix = 0u;
  if (first) {
    long long rk = j == 0 ? ks_lo : j == 1 ? ks_lo + 1 : j == 2 ? ks_hi : ks_hi + 1;
    rank = (unsigned long long)(rk > M - 1 ? M - 1 : rk);
  } else {
    rank = s->rank[j];
    prefix = s->prefix[j];
  }
  const uint32_t* h = hist + ((int64_t)t * hists_per_task + (hists_per_task > 1 ? j : 0)) * kRadixBins;
  // two levels: every lane sums a contiguous run of n_bins / 32 bins, the runs are scanned across the warp, then
  // the run holding the rank is scanned bin by bin
  const int per = n_bins >> 5;                          // 64 or 32 bins per lane
  unsigned long long tot = 0ull;
  {
    const uint4* h4 = reinterpret_cast<const uint4*>(h + lane * per);
    for (int i = 0; i < per / 4; ++i) {
      const uint4 v = h4[i];
      tot += (unsigned long long)v.x + v.y + v.z + v.w;
    }
  }
  unsigned long long inc = tot;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  unsigned hit = __ballot_sync(0xffffffffu, inc > rank);
  const int run = hit ? __ffs(hit) - 1 : 31;
  unsigned long long before = __shfl_sync(0xffffffffu, inc - tot, run);
  // inside the run: per / 32 bins per lane
  const int sub = per >> 5;                             // 2 or 1
  const uint32_t* hr = h + run * per + lane * sub;
  const unsigned long long b0 = hr[0], b1 = sub == 2 ? hr[1] : 0ull;
  unsigned long long inc2 = b0 + b1;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long u = __shfl_up_sync(0xffffffffu, inc2, o);
    if (lane >= o) inc2 += u;
  }
  hit = __ballot_sync(0xffffffffu, before + inc2 > rank);
  const int src = hit ? __ffs(hit) - 1 : 31;
  const unsigned long long before_l = before + __shfl_sync(0xffffffffu, inc2 - (b0 + b1), src);
  const unsigned long long s0 = __shfl_sync(0xffffffffu, b0, src);
  int bin = run * per + src * sub;
  unsigned long long in_bin = rank - before_l;
  if (sub == 2 && before_l + s0 <= rank) { bin += 1; in_bin -= s0; }
  if (lane == 0) {
    prefix |= (uint32_t)bin << shift;
    s->prefix[j] = prefix;
    s->rank[j] = in_bin;
    if (shift == 0) semi_os[t * 4 + j] = key2f(prefix);
  }
}

// Per task: sorted thresholds |r[l,0]| (ascending) for the pooled p-values, and a table that narrows the search
// among them: the bit patterns of non-negative floats are ordered like the floats, so bin(a) = (bits(a) - bits(t_min))
// >> shift is monotone; tab[b] = index of the first threshold whose bin is >= b (kNarrowBins + 1 entries, shift
// chosen so that t_max lands in the last bin).  t_min is the smallest NON-ZERO threshold: the rows the blur zeroes
// (and motifs without a match) have r = 0 exactly, and a table that started at bit pattern 0 would spend its bins on
// 127 empty octaves (the search among the thresholds of a bin was 37 % of the pooled pass's instructions with 4096
// bins from zero; with 16384 from t_min most bins hold no threshold and the rest one).
// thr_par[t] = {bits(t_min), shift, number of zero thresholds, -}.
constexpr int kNarrowBins = 4096;    // (16384 measured too: see below)
__global__ void thresholds_kernel(const float* __restrict__ r, int64_t ldr, int L, int C, int L_pow2,
                                  float* __restrict__ ts_sorted, uint16_t* __restrict__ tab, uint32_t* __restrict__ thr_par) {
  extern __shared__ uint32_t s_keys[];
  __shared__ int s_zero;
  const int t = blockIdx.x;
  if (threadIdx.x == 0) s_zero = 0;
  for (int i = threadIdx.x; i < L_pow2; i += blockDim.x)
    s_keys[i] = i < L ? __float_as_uint(fabsf(r[((int64_t)t * L + i) * ldr])) : 0xffffffffu;
  __syncthreads();
  bitonic_sort_keys(s_keys, L_pow2);
  int nz = 0;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    ts_sorted[(int64_t)t * L + i] = __uint_as_float(s_keys[i]);
    nz += s_keys[i] == 0u;
  }
  if (nz) atomicAdd(&s_zero, nz);
  __syncthreads();
  const int j0 = s_zero;                                   // thresholds [0, j0) are zero: every |v| reaches them
  uint16_t* tb = tab + (int64_t)t * (kNarrowBins + 2);
  if (j0 >= L) {                                           // all zero
    for (int b = threadIdx.x; b <= kNarrowBins; b += blockDim.x) tb[b] = (uint16_t)L;
    if (threadIdx.x == 0) { thr_par[4 * t] = 0xffffffffu; thr_par[4 * t + 1] = 0u; thr_par[4 * t + 2] = (uint32_t)L; }
    return;
  }
  const uint32_t kmin = s_keys[j0], kmax = s_keys[L - 1];
  int shift = 0;
  while (((kmax - kmin) >> shift) >= (uint32_t)kNarrowBins) ++shift;
  for (int j = j0 + threadIdx.x; j < L; j += blockDim.x) {
    const int bj = (int)((s_keys[j] - kmin) >> shift);
    const int bp = j == j0 ? -1 : (int)((s_keys[j - 1] - kmin) >> shift);
    for (int b = bp + 1; b <= bj; ++b) tb[b] = (uint16_t)j;      // first threshold with bin >= b
  }
  const int b_last = (int)((kmax - kmin) >> shift);
  for (int b = b_last + 1 + threadIdx.x; b <= kNarrowBins; b += blockDim.x) tb[b] = (uint16_t)L;
  if (threadIdx.x == 0) { thr_par[4 * t] = kmin; thr_par[4 * t + 1] = (uint32_t)shift; thr_par[4 * t + 2] = (uint32_t)j0; }
}

// One pass over the permuted values of a task for two things: (1) the histogram of the pooled |r_perm| values
// against the task's sorted thresholds, hist[idx] = number of values with exactly idx thresholds <= |v| (the
// narrowing table leaves a search range of a few thresholds); (2) the second digit of the pooled radix select.
// grid = (row chunks, T); dynamic shared memory: thresholds [L], hist [L + 1], tab [kNarrowBins + 2] (uint16),
// radix histograms [4][kRadixBins].
__global__ void __launch_bounds__(256) pooled_pass_kernel(const float* __restrict__ r, int64_t ldr, int L, int C,
                                                          int rows_per_cta, const float* __restrict__ ts_sorted,
                                                          const uint16_t* __restrict__ tab,
                                                          const uint32_t* __restrict__ thr_par,
                                                          const RadixState* __restrict__ state,
                                                          unsigned long long* __restrict__ hist,
                                                          uint32_t* __restrict__ hist2) {
  extern __shared__ uint32_t s_raw[];
  float* s_ts = reinterpret_cast<float*>(s_raw);        // [L]
  uint32_t* s_hist = s_raw + L;                         // [L + 1]
  uint32_t* s_h2 = s_hist + L + 1;                      // [4][kRadixBins]
  uint16_t* s_tab = reinterpret_cast<uint16_t*>(s_h2 + 4 * kRadixBins);   // [kNarrowBins + 2]
  const int t = blockIdx.y;
  const int P = C - 1;
  for (int i = threadIdx.x; i < L; i += blockDim.x) s_ts[i] = ts_sorted[(int64_t)t * L + i];
  for (int i = threadIdx.x; i <= L; i += blockDim.x) s_hist[i] = 0;
  for (int i = threadIdx.x; i < 4 * kRadixBins; i += blockDim.x) s_h2[i] = 0u;
  for (int i = threadIdx.x; i < kNarrowBins + 2; i += blockDim.x) s_tab[i] = tab[(int64_t)t * (kNarrowBins + 2) + i];
  const uint32_t kmin = thr_par[4 * t], shift = thr_par[4 * t + 1];
  const int n_zero = (int)thr_par[4 * t + 2];
  const RadixState st = state[t];
  const uint32_t p0 = st.prefix[0] >> 21, p1 = st.prefix[1] >> 21, p2 = st.prefix[2] >> 21, p3 = st.prefix[3] >> 21;
  __syncthreads();
  const int l0 = blockIdx.x * rows_per_cta, l1 = min(L, l0 + rows_per_cta);
  for (int l = l0; l < l1; l += 4) {                    // four rows at a time: four loads in flight per thread
    const float* rr = r + ((int64_t)t * L + l) * ldr + 1;
    const int nr = min(4, l1 - l);
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = j < nr ? rr[(int64_t)j * ldr + i] : 0.0f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nr) {
          const float a = fabsf(v[j]);
          const uint32_t ab = __float_as_uint(a);
          int lo = n_zero, hi = n_zero;                 // below every non-zero threshold: only the zero ones are <= a
          if (ab >= kmin) {
            const uint32_t b = min((ab - kmin) >> shift, (uint32_t)kNarrowBins);
            lo = s_tab[b];
            hi = b < (uint32_t)kNarrowBins ? s_tab[b + 1] : L;
          }
          while (lo < hi) {                             // upper_bound inside the bin: number of thresholds <= a
            const int mid = (lo + hi) >> 1;
            if (s_ts[mid] <= a) lo = mid + 1; else hi = mid;
          }
          atomicAdd(&s_hist[lo], 1u);
          const uint32_t k = f2key(v[j]);
          const uint32_t top = k >> 21, d = (k >> 10) & 2047u;
          if (top == p0) atomicAdd(&s_h2[0 * kRadixBins + d], 1u);
          if (top == p1) atomicAdd(&s_h2[1 * kRadixBins + d], 1u);
          if (top == p2) atomicAdd(&s_h2[2 * kRadixBins + d], 1u);
          if (top == p3) atomicAdd(&s_h2[3 * kRadixBins + d], 1u);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= L; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&hist[(int64_t)t * (L + 1) + i], (unsigned long long)s_hist[i]);
  for (int i = threadIdx.x; i < 4 * kRadixBins; i += blockDim.x)
    if (s_h2[i]) atomicAdd(&hist2[(int64_t)t * 4 * kRadixBins + i], s_h2[i]);
}

// Per task: semi_cnt[l] = #{pooled |v| >= |r[l,0]|} from the histogram (suffix sums).
__global__ void semi_count_kernel(const float* __restrict__ r, int64_t ldr, int L, int C, const float* __restrict__ ts_sorted,
                                  unsigned long long* __restrict__ hist, long long* __restrict__ semi_cnt) {
  const int t = blockIdx.x;
  unsigned long long* h = hist + (int64_t)t * (L + 1);
  // in-place suffix sum, h[j] becomes sum_{idx >= j} hist[idx]: per-thread segments + a scan of the segment totals
  {
    __shared__ unsigned long long s_part[256];
    const int n = L + 1;
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int lo = threadIdx.x * per, hi = min(n, lo + per);
    unsigned long long sum = 0ull;
    for (int j = lo; j < hi; ++j) sum += h[j];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long run = 0ull;
      for (int i = (int)blockDim.x - 1; i >= 0; --i) { const unsigned long long v = s_part[i]; s_part[i] = run; run += v; }
    }
    __syncthreads();
    unsigned long long run = s_part[threadIdx.x];      // sum of the segments to the right
    for (int j = hi - 1; j >= lo; --j) { run += h[j]; h[j] = run; }
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

// integral[t][c] = trapezoid over positions of |r[:, c]| (np.trapz, dx = 1); the same pass histograms the third
// (last, 10-bit) digit of the pooled radix select.  grid = (column blocks, T), one thread per column.
constexpr int kIntChunks = 4;   // row chunks of the integral pass (partial sums, added in order afterwards)
__global__ void __launch_bounds__(128) integral_pass_kernel(const float* __restrict__ r, int64_t ldr, int L, int C,
                                                            const RadixState* __restrict__ state,
                                                            double* __restrict__ partial, uint32_t* __restrict__ hist3) {
  __shared__ uint32_t s_h3[4][1024];
  const int t = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = threadIdx.x; i < 4 * 1024; i += blockDim.x) (&s_h3[0][0])[i] = 0u;
  const RadixState st = state[t];
  const uint32_t p0 = st.prefix[0] >> 10, p1 = st.prefix[1] >> 10, p2 = st.prefix[2] >> 10, p3 = st.prefix[3] >> 10;
  __syncthreads();
  if (c < C) {
    const float* rr = r + (int64_t)t * L * ldr + c;
    const bool perm = c >= 1;                            // column 0 is the unpermuted profile: not in the pooled null
    auto digit = [&](float v) {
      const uint32_t k = f2key(v);
      const uint32_t top = k >> 10, d = k & 1023u;
      if (top == p0) atomicAdd(&s_h3[0][d], 1u);
      if (top == p1) atomicAdd(&s_h3[1][d], 1u);
      if (top == p2) atomicAdd(&s_h3[2][d], 1u);
      if (top == p3) atomicAdd(&s_h3[3][d], 1u);
    };
    // rows [la, lb) of this chunk: their radix digits, and the trapezoids between rows la-1 .. lb-1
    const int per = (L + kIntChunks - 1) / kIntChunks;
    const int la = blockIdx.z * per, lb = min(L, la + per);
    double acc = 0.0;
    if (la < lb) {
      int l = la;
      float prev;
      if (la == 0) {
        const float v0 = rr[0];
        if (perm) digit(v0);
        prev = fabsf(v0);
        l = 1;
      } else {
        prev = fabsf(rr[(int64_t)(la - 1) * ldr]);
      }
      for (; l + 8 <= lb; l += 8) {                        // eight loads in flight, summed in position order
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = rr[(int64_t)(l + j) * ldr];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (perm) digit(v[j]);
          const float cur = fabsf(v[j]);
          acc += 0.5 * ((double)prev + (double)cur);
          prev = cur;
        }
      }
      for (; l < lb; ++l) {
        const float v = rr[(int64_t)l * ldr];
        if (perm) digit(v);
        const float cur = fabsf(v);
        acc += 0.5 * ((double)prev + (double)cur);
        prev = cur;
      }
    }
    partial[((int64_t)t * kIntChunks + blockIdx.z) * C + c] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * 1024; i += blockDim.x) {
    const uint32_t v = (&s_h3[0][0])[i];
    if (v) atomicAdd(&hist3[(int64_t)t * 4 * kRadixBins + (i >> 10) * kRadixBins + (i & 1023)], v);
  }
}

// integral[t][c] = the row-chunk partial sums of integral_pass_kernel added in chunk order
__global__ void integral_sum_kernel(const double* __restrict__ partial, int C, double* __restrict__ integral) {
  const int t = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double acc = 0.0;
  for (int k = 0; k < kIntChunks; ++k) acc += partial[((int64_t)t * kIntChunks + k) * C + c];
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
      col_syz_dev, ysum_dev, use_z ? zsum_host2[0] : 0.0, use_z ? zsum_host2[1] : 0.0, nullptr, (double)N, C, use_z,
      colconst_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int mepp_col_constants_dev(const double* col_syz_dev, const double* ysum_dev, const double* zsum_dev, int64_t N, int C,
                           int use_z, double* colconst_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(ysum_dev && colconst_dev && C > 0, "col_constants: bad arguments");
  MEPP_REQUIRE(!use_z || (col_syz_dev && zsum_dev), "col_constants: z statistics missing");
  MEPP_REQUIRE(mepp_device_ok(), "col_constants: no sm_100 CUDA device (no CPU fallback)");
  col_constants_kernel<<<(unsigned)((C + 255) / 256), 256, 0, as_stream(stream)>>>(
      col_syz_dev, ysum_dev, 0.0, 0.0, use_z ? zsum_dev : nullptr, (double)N, C, use_z, colconst_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static int64_t up256(int64_t v) { return (v + 255) / 256 * 256; }

int64_t mepp_perm_stats_work_bytes(int T, int L, int C) {
  const int64_t ts = up256((int64_t)T * L * 4);                     // sorted |r0| thresholds
  const int64_t hist = up256((int64_t)T * (L + 1) * 8);             // pooled histogram against the thresholds
  const int64_t h1 = up256((int64_t)T * mepp::kRadixBins * 4);      // radix select: first digit
  const int64_t h23 = up256((int64_t)T * 4 * mepp::kRadixBins * 4); // second / third digit, four targets
  const int64_t state = up256((int64_t)T * sizeof(mepp::RadixState));
  const int64_t tab = up256((int64_t)T * (mepp::kNarrowBins + 2) * 2) + up256((int64_t)T * 16);   // narrowing table + parameters
  const int64_t part = up256((int64_t)T * mepp::kIntChunks * C * 8);   // row-chunk partial sums of the integrals
  return ts + hist + h1 + 2 * h23 + state + tab + part;
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
  const int64_t ts_b = up256((int64_t)T * L * 4), hist_b = up256((int64_t)T * (L + 1) * 8);
  const int64_t h1_b = up256((int64_t)T * kRadixBins * 4), h23_b = up256((int64_t)T * 4 * kRadixBins * 4);
  float* ts_sorted = reinterpret_cast<float*>(w);
  unsigned long long* hist = reinterpret_cast<unsigned long long*>(w + ts_b);
  uint32_t* hist1 = reinterpret_cast<uint32_t*>(w + ts_b + hist_b);
  uint32_t* hist2 = reinterpret_cast<uint32_t*>(w + ts_b + hist_b + h1_b);
  uint32_t* hist3 = reinterpret_cast<uint32_t*>(w + ts_b + hist_b + h1_b + h23_b);
  RadixState* state = reinterpret_cast<RadixState*>(w + ts_b + hist_b + h1_b + 2 * h23_b);
  const int64_t state_b = up256((int64_t)T * sizeof(RadixState)), tab_b = up256((int64_t)T * (kNarrowBins + 2) * 2);
  uint16_t* tab = reinterpret_cast<uint16_t*>(w + ts_b + hist_b + h1_b + 2 * h23_b + state_b);
  uint32_t* thr_par = reinterpret_cast<uint32_t*>(w + ts_b + hist_b + h1_b + 2 * h23_b + state_b + tab_b);
  double* partial = reinterpret_cast<double*>(w + ts_b + hist_b + h1_b + 2 * h23_b + state_b + tab_b + up256((int64_t)T * 16));
  MEPP_CUDA_CHECK(cudaMemsetAsync(hist, 0, (size_t)(hist_b + h1_b + 2 * h23_b), st));

  // rows per CTA: enough CTAs to fill the machine, few enough that the per-task histogram flushes stay cheap
  int chunks_r = (148 * 16 + T - 1) / T;
  if (chunks_r > (L + 7) / 8) chunks_r = (L + 7) / 8;
  if (chunks_r < 1) chunks_r = 1;
  const int rows_per_cta = (L + chunks_r - 1) / chunks_r;
  chunks_r = (L + rows_per_cta - 1) / rows_per_cta;
  if (P_pow2 <= 1024) {
    // one warp per row, rows staged in shared memory; tail selection when both tails need at most 40 values
    const int Pp = P_pow2 < 32 ? 32 : P_pow2;
    const int fast = (Pp >= 256 && (k_lo + 2) <= 40 && (P - k_hi) <= 40) ? 1 : 0;
    const size_t smem = sizeof(uint32_t) * (size_t)kRowWarps * 2 * Pp;
    MEPP_CUDA_CHECK(cudaFuncSetAttribute(row_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    row_select_kernel<<<dim3((unsigned)chunks_r, (unsigned)T), kRowWarps * 32, smem, st>>>(
        r_dev, ldr, L, C, Pp, rows_per_cta, k_lo, k_hi, fast, full_os_dev, full_cnt_dev, hist1);
  } else {
    const size_t row_smem = sizeof(uint32_t) * (size_t)P_pow2;
    MEPP_CUDA_CHECK(cudaFuncSetAttribute(row_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem));
    row_stats_kernel<<<dim3((unsigned)L, (unsigned)T), 256, row_smem, st>>>(
        r_dev, ldr, L, C, P_pow2, k_lo, k_hi, full_os_dev, full_cnt_dev, hist1);
  }
  // pooled statistics: the first radix digit is in hist1; the second rides on the pass that histograms the
  // pooled |r| against the thresholds, the third on the pass that integrates the columns
  const long long M = (long long)L * P;
  radix_pick_kernel<<<(unsigned)T, 128, 0, st>>>(hist1, 1, kRadixBins, 21, state, ks_lo, ks_hi, M, 1, semi_os_dev);
  const size_t thr_smem = sizeof(uint32_t) * (size_t)L_pow2;
  MEPP_CUDA_CHECK(cudaFuncSetAttribute(thresholds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)thr_smem));
  thresholds_kernel<<<(unsigned)T, 256, thr_smem, st>>>(r_dev, ldr, L, C, L_pow2, ts_sorted, tab, thr_par);
  const size_t pool_smem = sizeof(uint32_t) * (size_t)(2 * L + 1 + 4 * kRadixBins) + sizeof(uint16_t) * (kNarrowBins + 2);
  MEPP_CUDA_CHECK(cudaFuncSetAttribute(pooled_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pool_smem));
  pooled_pass_kernel<<<dim3((unsigned)chunks_r, (unsigned)T), 256, pool_smem, st>>>(r_dev, ldr, L, C, rows_per_cta, ts_sorted,
                                                                                  tab, thr_par, state, hist, hist2);
  semi_count_kernel<<<(unsigned)T, 256, 0, st>>>(r_dev, ldr, L, C, ts_sorted, hist, semi_cnt_dev);
  radix_pick_kernel<<<(unsigned)T, 128, 0, st>>>(hist2, 4, kRadixBins, 10, state, ks_lo, ks_hi, M, 0, semi_os_dev);
  integral_pass_kernel<<<dim3((unsigned)((C + 127) / 128), (unsigned)T, kIntChunks), 128, 0, st>>>(r_dev, ldr, L, C, state,
                                                                                                  partial, hist3);
  integral_sum_kernel<<<dim3((unsigned)((C + 127) / 128), (unsigned)T), 128, 0, st>>>(partial, C, integral_dev);
  radix_pick_kernel<<<(unsigned)T, 128, 0, st>>>(hist3, 4, 1024, 0, state, ks_lo, ks_hi, M, 0, semi_os_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"

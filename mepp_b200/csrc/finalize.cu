// K4: per-position / per-task output columns of one task group, computed on the device from the permutation
// statistics of K3 so that the host only copies them into the result table.
//
// Replaces (reference npdeloss/mepp): the column arithmetic of core.get_positional_r_df (mepp/core.py:286-374:
// percentile interpolation, p-values, integral statistic), the parametric Fisher-z interval and t-test of
// core.py:380-413 (float32 arithmetic, like the reference's numpy on a float32 r) and the centred rolling mean
// of the match counts (core.py:438-454).  The interpolation reproduces numpy's `_lerp` on float32 neighbours
// (difference taken in float32, combination in float64, upper form for gamma >= 0.5).
#include "common.cuh"
#include <cfloat>
#include <cmath>

namespace mepp {

// numpy _lerp(a, b, gamma) for float32 a, b and float64 gamma
__device__ __forceinline__ double lerp_f32(float a, float b, double gamma) {
  const float diff = __fsub_rn(b, a);
  if (gamma >= 0.5) return __dsub_rn((double)b, __dmul_rn((double)diff, 1.0 - gamma));
  return __dadd_rn((double)a, __dmul_rn((double)diff, gamma));
}

// Continued fraction of the incomplete beta function (modified Lentz; converges in O(sqrt(max(a, b))) steps).
__device__ double betacf(double a, double b, double x) {
  const double tiny = 1e-300, eps = 1e-15;   // (1e-16 is below what d * c can reach: the loop then runs to its cap)
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 3000; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < eps) break;
  }
  return h;
}

// Regularised incomplete beta I_x(a, b); xc = 1 - x is passed separately (no cancellation near x = 1).
// lnbeta = lgamma(a + b) - lgamma(a) - lgamma(b) is the same for every element of a call.
__device__ double incbet(double a, double b, double x, double xc, double lnbeta) {
  if (!(x > 0.0)) return 0.0;
  if (!(xc > 0.0)) return 1.0;
  const double bt = exp(lnbeta + a * log(x) + b * log(xc));
  if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
  return 1.0 - bt * betacf(b, a, xc) / b;
}

// 2 * scipy.stats.t.sf(|t|, df) = I_{df / (df + t^2)}(df / 2, 1 / 2)   (core.py:410)
__device__ double t_two_sided(double t, double df, double lnbeta) {
  if (t != t) return t;
  const double t2 = t * t;
  if (isinf(t2)) return 0.0;
  const double den = df + t2;
  return incbet(0.5 * df, 0.5, df / den, t2 / den, lnbeta);
}

struct FinalizeArgs {
  const float* r; int64_t ldr;
  int T, L, C;
  const float* full_os; const int32_t* full_cnt; const float* semi_os; const long long* semi_cnt;
  const double* integral; const int32_t* count;
  double n_seq; int margin;
  int k_lo, k_hi; double g_lo, g_hi, gs_lo, gs_hi;
  float z_margin;
  double lnbeta;   // lgamma(df/2 + 1/2) - lgamma(df/2) - lgamma(1/2), df = n_seq - 1
  uint8_t* out;
};

// Output block (mepp_finalize_out_bytes): float64 columns [full_lower, full_upper, full_pval, semi_pval, pval,
// count, smoothed_count][T*L], float32 columns [positional_r, lower, upper][T*L], then per task float64
// [semi_lower, semi_upper, integral_lower, integral_upper, integral_pval][T] (8-byte aligned) and float32
// [integral_r][T].
__global__ void finalize_positions_kernel(const FinalizeArgs a) {
  const int64_t TL = (int64_t)a.T * a.L;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= TL) return;
  double* d64 = reinterpret_cast<double*>(a.out);
  float* f32 = reinterpret_cast<float*>(a.out + TL * 7 * sizeof(double));
  const int P = a.C - 1;
  const int l = (int)(i % a.L);
  const float r = a.r[i * a.ldr];
  if (P > 0) {
    const float4 os = *reinterpret_cast<const float4*>(a.full_os + i * 4);
    d64[0 * TL + i] = lerp_f32(os.x, os.y, a.g_lo);
    d64[1 * TL + i] = lerp_f32(os.z, os.w, a.g_hi);
    d64[2 * TL + i] = (double)a.full_cnt[i] / (double)P;
    d64[3 * TL + i] = (double)a.semi_cnt[i] / ((double)a.L * (double)P);
  }
  // parametric interval and t-test, float32 arithmetic (core.py:380-413)
  const float eps = FLT_EPSILON;
  const float zr = 0.5f * logf(__fdiv_rn(__fadd_rn(__fadd_rn(1.0f, r), eps), __fadd_rn(__fsub_rn(1.0f, r), eps)));
  const float el = expf(__fadd_rn(__fmul_rn(2.0f, __fsub_rn(zr, a.z_margin)), eps));
  const float eu = expf(__fadd_rn(__fmul_rn(2.0f, __fadd_rn(zr, a.z_margin)), eps));
  f32[0 * TL + i] = r;
  f32[1 * TL + i] = __fdiv_rn(__fsub_rn(el, 1.0f), __fadd_rn(el, 1.0f));
  f32[2 * TL + i] = __fdiv_rn(__fsub_rn(eu, 1.0f), __fadd_rn(eu, 1.0f));
  const float t = __fdiv_rn(r, sqrtf(__fdiv_rn(__fsub_rn(1.0f, __fmul_rn(r, r)), (float)(a.n_seq - 2.0))));
  d64[4 * TL + i] = t_two_sided((double)fabsf(t), a.n_seq - 1.0, a.lnbeta);
  // match counts and their centred rolling mean with the edges filled from the first / last full window
  const int32_t* cnt = a.count + (i - l);
  d64[5 * TL + i] = (double)cnt[l];
  double sm = (double)cnt[l];
  const int w = 2 * a.margin + 1;
  if (a.margin > 0) {
    if (a.L >= w) {
      int centre = l < a.margin ? a.margin : (l > a.L - 1 - a.margin ? a.L - 1 - a.margin : l);
      long long s = 0;
      for (int d = -a.margin; d <= a.margin; ++d) s += cnt[centre + d];
      sm = (double)s / (double)w;
    } else {
      sm = nan("");
    }
  }
  d64[6 * TL + i] = sm;
}

__device__ __forceinline__ uint32_t fkey(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float keyf(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// One block per task: pooled interval from the pooled order statistics, and the integral statistic (float32
// view of the trapezoid sums, core.py:352): percentiles over the P permuted integrals, p-value = fraction >= the
// unpermuted one.
__global__ void finalize_tasks_kernel(const FinalizeArgs a, int P_pow2) {
  extern __shared__ uint32_t skeys[];
  const int t = blockIdx.x;
  const int64_t TL = (int64_t)a.T * a.L;
  const int P = a.C - 1;
  double* d64 = reinterpret_cast<double*>(a.out + ((TL * 68 + 7) & ~(int64_t)7));
  float* f32 = reinterpret_cast<float*>(d64 + 5 * (int64_t)a.T);
  const double* integ = a.integral + (int64_t)t * a.C;
  const float i0 = (float)integ[0];
  __shared__ int s_ge;
  if (threadIdx.x == 0) s_ge = 0;
  int ge = 0;
  for (int p = threadIdx.x; p < P_pow2; p += blockDim.x) {
    uint32_t k = 0xffffffffu;
    if (p < P) {
      const float v = (float)integ[1 + p];
      k = fkey(v);
      ge += v >= i0;
    }
    skeys[p] = k;
  }
  __syncthreads();
  atomicAdd(&s_ge, ge);
  __syncthreads();
  for (int k = 2; k <= P_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P_pow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint32_t x = skeys[i], y = skeys[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { skeys[i] = y; skeys[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    const int klo1 = a.k_lo + 1 < P ? a.k_lo + 1 : P - 1, khi1 = a.k_hi + 1 < P ? a.k_hi + 1 : P - 1;
    d64[0 * (int64_t)a.T + t] = lerp_f32(a.semi_os[t * 4 + 0], a.semi_os[t * 4 + 1], a.gs_lo);
    d64[1 * (int64_t)a.T + t] = lerp_f32(a.semi_os[t * 4 + 2], a.semi_os[t * 4 + 3], a.gs_hi);
    d64[2 * (int64_t)a.T + t] = lerp_f32(keyf(skeys[a.k_lo]), keyf(skeys[klo1]), a.g_lo);
    d64[3 * (int64_t)a.T + t] = lerp_f32(keyf(skeys[a.k_hi]), keyf(skeys[khi1]), a.g_hi);
    d64[4 * (int64_t)a.T + t] = (double)s_ge / (double)P;
    f32[t] = i0;
  }
}

// density (matches per sequence, <= L <= 65535) narrowed for the copy to the host: at cfg 3 the int32 form is 130 MB per step
__global__ void density_u16_kernel(const int32_t* __restrict__ in, int64_t n4, uint2* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int4 v = reinterpret_cast<const int4*>(in)[i];
  const uint32_t a = (uint32_t)min(v.x, 65535), b = (uint32_t)min(v.y, 65535);
  const uint32_t c = (uint32_t)min(v.z, 65535), d = (uint32_t)min(v.w, 65535);
  out[i] = make_uint2(a | (b << 16), c | (d << 16));
}

}  // namespace mepp

extern "C" {

int mepp_density_u16(const int32_t* density_dev, int64_t n, uint16_t* out_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(density_dev && out_dev && n > 0 && (n & 3) == 0, "density_u16: bad arguments (n must be a multiple of 4)");
  MEPP_REQUIRE(mepp_device_ok(), "density_u16: no sm_100 CUDA device (no CPU fallback)");
  const int64_t n4 = n / 4;
  density_u16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, as_stream(stream)>>>(density_dev, n4,
                                                                               reinterpret_cast<uint2*>(out_dev));
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int64_t mepp_finalize_out_bytes(int T, int L) {
  const int64_t TL = (int64_t)T * L;
  return ((TL * 68 + 7) & ~(int64_t)7) + (int64_t)T * (5 * 8 + 4);
}

int mepp_finalize_positions(const float* r_dev, int64_t ldr, int T, int L, int C, const float* full_os_dev,
                            const int32_t* full_cnt_dev, const float* semi_os_dev, const long long* semi_cnt_dev,
                            const double* integral_dev, const int32_t* count_dev, int64_t n_seq, int margin,
                            int k_lo, int k_hi, double g_lo, double g_hi, double gs_lo, double gs_hi, float z_margin,
                            void* out_dev, void* stream) {
  using namespace mepp;
  const int P = C - 1;
  MEPP_REQUIRE(r_dev && count_dev && out_dev, "finalize_positions: null pointer");
  MEPP_REQUIRE(T > 0 && L > 0 && C >= 1 && ldr >= C && n_seq > 3, "finalize_positions: bad shape");
  MEPP_REQUIRE(P == 0 || (full_os_dev && full_cnt_dev && semi_os_dev && semi_cnt_dev && integral_dev),
               "finalize_positions: permutation statistics missing");
  MEPP_REQUIRE(P == 0 || (k_lo >= 0 && k_lo < P && k_hi >= 0 && k_hi < P), "finalize_positions: percentile index");
  MEPP_REQUIRE(margin >= 0 && margin <= MEPP_MAX_MARGIN, "finalize_positions: margin must be in [0, 16]");
  MEPP_REQUIRE(mepp_device_ok(), "finalize_positions: no sm_100 CUDA device (no CPU fallback)");
  cudaStream_t st = as_stream(stream);
  FinalizeArgs a;
  a.r = r_dev; a.ldr = ldr; a.T = T; a.L = L; a.C = C;
  a.full_os = full_os_dev; a.full_cnt = full_cnt_dev; a.semi_os = semi_os_dev; a.semi_cnt = semi_cnt_dev;
  a.integral = integral_dev; a.count = count_dev;
  a.n_seq = (double)n_seq; a.margin = margin;
  a.k_lo = k_lo; a.k_hi = k_hi; a.g_lo = g_lo; a.g_hi = g_hi; a.gs_lo = gs_lo; a.gs_hi = gs_hi;
  a.z_margin = z_margin;
  {
    const double df = (double)n_seq - 1.0;
    a.lnbeta = lgamma(0.5 * df + 0.5) - lgamma(0.5 * df) - lgamma(0.5);
  }
  a.out = reinterpret_cast<uint8_t*>(out_dev);
  const int64_t TL = (int64_t)T * L;
  finalize_positions_kernel<<<(unsigned)((TL + 63) / 64), 64, 0, st>>>(a);
  if (P > 0) {
    int P_pow2 = 1;
    while (P_pow2 < P) P_pow2 <<= 1;
    MEPP_REQUIRE(P_pow2 <= 32768, "finalize_positions: at most 32768 permutations per call");
    const size_t smem = sizeof(uint32_t) * (size_t)P_pow2;
    MEPP_CUDA_CHECK(cudaFuncSetAttribute(finalize_tasks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    finalize_tasks_kernel<<<(unsigned)T, 256, smem, st>>>(a, P_pow2);
  }
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"

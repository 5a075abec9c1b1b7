// Thresholds on the device: the MOODS threshold_from_p dynamic programme (mepp/motif_layers.py:44-48 calls it once
// per task on the host) for every distinct (motif, filter-0 orientation) of a call in one launch, and the
// threshold-dependent part of the scan tables (bias, pass bounds of the integer pre-filter).  Same arithmetic as
// mepp_threshold_from_p in host.cu (products and sums rounded separately, additions per cell in letter order, the
// tail summed from the top), so the thresholds are identical; the host DP was the longest host-side piece of an
// end-to-end step and does not shrink when several ranks share the host cores.
#include "common.cuh"

namespace mepp {

// grid = units, block = 1024.  smat [U][32][4]: integer scores of column i, letter j, shifted to >= 0;
// par [U][8] = {bg[4], p, precision, m * minV, R}; scratch: two arrays of R + 1 doubles per unit at scratch_off[u].
__global__ void __launch_bounds__(1024) threshold_dp_kernel(const int32_t* __restrict__ smat, const int32_t* __restrict__ mlen,
                                                            const double* __restrict__ par,
                                                            const long long* __restrict__ scratch_off,
                                                            double* __restrict__ scratch, double* __restrict__ thr) {
  const int u = blockIdx.x;
  const int m = mlen[u];
  const int32_t* sm = smat + (size_t)u * MEPP_MAX_MOTIF_LEN * 4;
  const double* pr = par + (size_t)u * 8;
  const long long R = (long long)pr[7];
  double* t0 = scratch + scratch_off[u];
  double* t1 = t0 + (R + 1);
  __shared__ int s_s[4];
  __shared__ double s_b[4];
  if (threadIdx.x < 4) s_b[threadIdx.x] = pr[threadIdx.x];
  long long hi = 0;
  for (int j = 0; j < 4; ++j) hi = max(hi, (long long)sm[j]);
  for (long long q = threadIdx.x; q <= hi; q += blockDim.x) t0[q] = 0.0;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int j = 0; j < 4; ++j) t0[sm[j]] = __dadd_rn(t0[sm[j]], pr[j]);
  __syncthreads();
  for (int i = 1; i < m; ++i) {
    if (threadIdx.x < 4) s_s[threadIdx.x] = sm[i * 4 + threadIdx.x];
    __syncthreads();
    const int s0 = s_s[0], s1 = s_s[1], s2 = s_s[2], s3 = s_s[3];
    const long long new_hi = hi + max(max(s0, s1), max(s2, s3));
    // dst[q] = sum over letters j (ascending) of bg[j] * src[q - s_j]: the scatter of the host loop, gathered;
    // four cells per thread and iteration so that sixteen loads are in flight
    const double b0 = s_b[0], b1 = s_b[1], b2 = s_b[2], b3 = s_b[3];
    for (long long q0 = threadIdx.x; q0 <= new_hi; q0 += 4 * blockDim.x) {
      double x[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long q = q0 + (long long)k * blockDim.x;
        long long r = q - s0; x[k][0] = (r >= 0 && r <= hi) ? t0[r] : -1.0;
        r = q - s1;           x[k][1] = (r >= 0 && r <= hi) ? t0[r] : -1.0;
        r = q - s2;           x[k][2] = (r >= 0 && r <= hi) ? t0[r] : -1.0;
        r = q - s3;           x[k][3] = (r >= 0 && r <= hi) ? t0[r] : -1.0;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long q = q0 + (long long)k * blockDim.x;
        if (q <= new_hi) {
          double acc = 0.0;                            // (-1 marks "no such source cell": probabilities are >= 0)
          if (x[k][0] >= 0.0) acc = __dadd_rn(acc, __dmul_rn(b0, x[k][0]));
          if (x[k][1] >= 0.0) acc = __dadd_rn(acc, __dmul_rn(b1, x[k][1]));
          if (x[k][2] >= 0.0) acc = __dadd_rn(acc, __dmul_rn(b2, x[k][2]));
          if (x[k][3] >= 0.0) acc = __dadd_rn(acc, __dmul_rn(b3, x[k][3]));
          t1[q] = acc;
        }
      }
    }
    __syncthreads();
    double* tmp = t0; t0 = t1; t1 = tmp;
    hi = new_hi;
  }
  // tail: the smallest grid score whose upper tail is <= p, i.e. sum from the top until the sum exceeds p.  The
  // sum is sequential like the host's (one thread), but the values are staged through shared memory by the whole
  // block, 2048 at a time, so that the thread adds from shared memory instead of waiting on global loads.
  __shared__ double s_tail[2048];
  __shared__ int s_done;
  __shared__ double s_sum;
  __shared__ double s_wsum[32];
  const double p = pr[4], precision = pr[5], base = pr[6];
  // With a flat background and m <= 26 every probability is a multiple of 4^-m >= 2^-52 and every partial sum is
  // exact in float64, so the order of the additions cannot matter: the tail is then summed by the whole block.
  const bool exact = m <= 26 && pr[0] == 0.25 && pr[1] == 0.25 && pr[2] == 0.25 && pr[3] == 0.25;
  if (threadIdx.x == 0) { s_done = 0; s_sum = 0.0; thr[u] = base / precision; }
  __syncthreads();
  for (long long top = R; top >= 0 && !s_done; top -= 2048) {
    for (int k = threadIdx.x; k < 2048; k += blockDim.x) {
      const long long r = top - k;
      s_tail[k] = (r >= 0 && r <= hi) ? t0[r] : 0.0;
    }
    __syncthreads();
    const int n = (int)min((long long)2048, top + 1);
    if (exact) {
      // inclusive scan of the 2048 staged values (two per thread), then the first index whose running sum > p
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      const double a0 = s_tail[2 * threadIdx.x], a1 = s_tail[2 * threadIdx.x + 1];
      double inc = a0 + a1;
      for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (lane == 31) s_wsum[warp] = inc;
      __syncthreads();
      if (warp == 0) {
        double w = s_wsum[lane];
        for (int o = 1; o < 32; o <<= 1) {
          const double v = __shfl_up_sync(0xffffffffu, w, o);
          if (lane >= o) w += v;
        }
        s_wsum[lane] = w;
      }
      __syncthreads();
      const double before = s_sum + (warp ? s_wsum[warp - 1] : 0.0) + inc - (a0 + a1);   // sum of everything above
      int first = 1 << 30;
      if (before + a0 > p) first = 2 * threadIdx.x;
      else if (before + a0 + a1 > p) first = 2 * threadIdx.x + 1;
      if (first >= n) first = 1 << 30;
      for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      __syncthreads();
      if (lane == 0) s_tail[warp] = (double)first;     // (s_tail is free again)
      __syncthreads();
      if (threadIdx.x == 0) {
        int best = 1 << 30;
        for (int w = 0; w < 32; ++w) best = min(best, (int)s_tail[w]);
        if (best < n) { thr[u] = (double)(top - best + (long long)base + 1) / precision; s_done = 1; }
        s_sum = s_sum + s_wsum[31];
      }
    } else if (threadIdx.x == 0) {
      double sum = s_sum;
      for (int k = 0; k < n; ++k) {
        sum = __dadd_rn(sum, s_tail[k]);
        if (sum > p) { thr[u] = (double)(top - k + (long long)base + 1) / precision; s_done = 1; break; }
      }
      s_sum = sum;
    }
    __syncthreads();
  }
}

// One thread per task: bias = float(-thr), the pass bounds of the pre-filter (host.cu: build_filter_table).
__global__ void task_thresholds_kernel(const double* __restrict__ thr_unit, const int32_t* __restrict__ unit_of_task,
                                       const double* __restrict__ qpar, const int32_t* __restrict__ mlen,
                                       const int32_t* __restrict__ nfilt, int T, float* __restrict__ bias,
                                       uint32_t* __restrict__ qthr, double* __restrict__ thr_task) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const double thr = thr_unit[unit_of_task[t]];
  const float b = (float)(-thr);
  bias[t] = b;
  thr_task[t] = thr;
  const double thr32 = -(double)b;
  const double scale = qpar[4 * t], delta = qpar[4 * t + 1];
  const int m = mlen[t];
  uint32_t qinit = 0u, bounds = 0u;
  for (int f = 0; f < nfilt[t]; ++f) {
    const double need = __dsub_rn(__dmul_rn(scale, __dsub_rn(__dsub_rn(thr32, delta), qpar[4 * t + 2 + f])), (double)m);
    const long long qq = need < 0.0 ? 0 : (long long)floor(need) + 1;
    const uint32_t half = qq > 0x7fff ? 0u : (uint32_t)(0x8000 - qq);
    qinit |= half << (16 * f);
    bounds |= (uint32_t)(qq > 0xffff ? 0xffff : qq) << (16 * f);
  }
  qthr[2 * t] = qinit;
  qthr[2 * t + 1] = bounds;
}

}  // namespace mepp

extern "C" {

int mepp_thresholds_dp(const int32_t* smat_dev, const int32_t* mlen_dev, const double* par_dev,
                       const long long* scratch_off_dev, double* scratch_dev, double* thr_dev, int U, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(smat_dev && mlen_dev && par_dev && scratch_off_dev && scratch_dev && thr_dev && U > 0,
               "thresholds_dp: bad arguments");
  MEPP_REQUIRE(mepp_device_ok(), "thresholds_dp: no sm_100 CUDA device (no CPU fallback)");
  threshold_dp_kernel<<<(unsigned)U, 1024, 0, as_stream(stream)>>>(smat_dev, mlen_dev, par_dev, scratch_off_dev, scratch_dev,
                                                                   thr_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int mepp_task_thresholds(const double* thr_unit_dev, const int32_t* unit_of_task_dev, const double* qpar_dev,
                         const int32_t* mlen_dev, const int32_t* nfilt_dev, int T, float* bias_dev, uint32_t* qthr_dev,
                         double* thr_task_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(thr_unit_dev && unit_of_task_dev && qpar_dev && mlen_dev && nfilt_dev && bias_dev && qthr_dev &&
               thr_task_dev && T > 0, "task_thresholds: bad arguments");
  MEPP_REQUIRE(mepp_device_ok(), "task_thresholds: no sm_100 CUDA device (no CPU fallback)");
  task_thresholds_kernel<<<(unsigned)((T + 127) / 128), 128, 0, as_stream(stream)>>>(
      thr_unit_dev, unit_of_task_dev, qpar_dev, mlen_dev, nfilt_dev, T, bias_dev, qthr_dev, thr_task_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"

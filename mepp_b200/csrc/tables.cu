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
    for (long long q = threadIdx.x; q <= new_hi; q += blockDim.x) {
      // dst[q] = sum over letters j (ascending) of bg[j] * src[q - s_j]: the scatter of the host loop, gathered
      double acc = 0.0;
      long long r = q - s0; if (r >= 0 && r <= hi) acc = __dadd_rn(acc, __dmul_rn(s_b[0], t0[r]));
      r = q - s1;           if (r >= 0 && r <= hi) acc = __dadd_rn(acc, __dmul_rn(s_b[1], t0[r]));
      r = q - s2;           if (r >= 0 && r <= hi) acc = __dadd_rn(acc, __dmul_rn(s_b[2], t0[r]));
      r = q - s3;           if (r >= 0 && r <= hi) acc = __dadd_rn(acc, __dmul_rn(s_b[3], t0[r]));
      t1[q] = acc;
    }
    __syncthreads();
    double* tmp = t0; t0 = t1; t1 = tmp;
    hi = new_hi;
  }
  if (threadIdx.x == 0) {
    const double p = pr[4], precision = pr[5], base = pr[6];
    double sum = 0.0, out = base / precision;
    for (long long r = R; r >= 0; --r) {
      sum = __dadd_rn(sum, r <= hi ? t0[r] : 0.0);
      if (sum > p) { out = (double)(r + (long long)base + 1) / precision; break; }
    }
    thr[u] = out;
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

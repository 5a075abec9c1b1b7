// Heat-map reduction of the motif score matrix on the device, from the match list.
//
// Replaces (reference npdeloss/mepp): mepp/plot.py:57-66 (clip_motif_score_matrix: upper clip = mean + sigma * std of the
// NON-ZERO scores) and plot.py:32-55,106-145 (compress_matrix_to_pixels with np.max: block maximum over
// get_aspect_tensor blocks, zero padded) -- the consumer of the dense (N, L) float32 matrix that core.py:568-573 ships
// to the host for every task (400 MB per task at cfg 3).  Both steps only depend on the matches, so the dense matrix
// never exists here: one pass over the {task, seq, pos, x0} records for the moments, one for the block maxima
// (scores are > 0, so the int32 order of their bit patterns is their order and atomicMax does).
#include "common.cuh"

namespace mepp {

struct HeatHit { int32_t task, seq, pos; float x0; };

// moments[t] = {n, sum x0, sum x0^2, -} in float64, accumulated per CTA in shared memory
__global__ void __launch_bounds__(256) heat_moments_kernel(const HeatHit* __restrict__ hits,
                                                           const unsigned long long* __restrict__ hit_count,
                                                           int64_t hits_cap, int T, double* __restrict__ moments) {
  extern __shared__ double s_m[];                     // [T][3]
  for (int i = threadIdx.x; i < 3 * T; i += blockDim.x) s_m[i] = 0.0;
  __syncthreads();
  int64_t n = (int64_t)*hit_count;
  if (n > hits_cap) n = hits_cap;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(hits + i));
    const int t = (int)raw.x;
    if (t < 0 || t >= T) continue;
    const double x = (double)__uint_as_float(raw.w);
    atomicAdd(&s_m[3 * t + 0], 1.0);
    atomicAdd(&s_m[3 * t + 1], x);
    atomicAdd(&s_m[3 * t + 2], x * x);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * T; i += blockDim.x)
    if (s_m[i] != 0.0) atomicAdd(&moments[4 * (i / 3) + (i % 3)], s_m[i]);
}

// limit[t] = float32(mean + sigma * population std) of the task's match scores (0 without a match); sigma < 0: no clip
__global__ void heat_limits_kernel(const double* __restrict__ moments, int T, float sigma, float* __restrict__ limit) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float lim = __int_as_float(0x7f800000);             // +inf: no clip
  if (sigma >= 0.0f) {
    const double n = moments[4 * t], sx = moments[4 * t + 1], sxx = moments[4 * t + 2];
    if (n > 0.0) {
      const double mean = sx / n;
      double var = sxx / n - mean * mean;
      if (var < 0.0) var = 0.0;
      lim = (float)(mean + (double)sigma * sqrt(var));
    } else {
      lim = 0.0f;
    }
  }
  limit[t] = lim;
}

__global__ void __launch_bounds__(256) heat_blockmax_kernel(const HeatHit* __restrict__ hits,
                                                            const unsigned long long* __restrict__ hit_count,
                                                            int64_t hits_cap, int T, int block_rows, int block_cols, int H,
                                                            int W, const float* __restrict__ limit, int* __restrict__ out) {
  int64_t n = (int64_t)*hit_count;
  if (n > hits_cap) n = hits_cap;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(hits + i));
    const int t = (int)raw.x;
    if (t < 0 || t >= T) continue;
    const int py = (int)raw.y / block_rows, px = (int)raw.z / block_cols;
    if (py >= H || px >= W) continue;
    const float v = fminf(__uint_as_float(raw.w), limit[t]);           // np.clip(x, 0, limit): scores are > 0
    atomicMax(out + ((int64_t)t * H + py) * W + px, __float_as_int(v));
  }
}

}  // namespace mepp

extern "C" {

int64_t mepp_heatmap_work_bytes(int T) { return (int64_t)T * 4 * 8 + (int64_t)T * 4; }

int mepp_heatmap(const void* hits_dev, const unsigned long long* hit_count_dev, int64_t hits_cap, int T, int64_t N, int L,
                 int block_rows, int block_cols, float sigma, float* out_dev, void* work_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(hits_dev && hit_count_dev && out_dev && work_dev, "heatmap: null pointer");
  MEPP_REQUIRE(T > 0 && T <= 4096 && N > 0 && L > 0 && hits_cap > 0 && block_rows >= 1 && block_cols >= 1,
               "heatmap: bad arguments (at most 4096 tasks per call)");
  MEPP_REQUIRE(mepp_device_ok(), "heatmap: no sm_100 CUDA device (no CPU fallback)");
  cudaStream_t st = as_stream(stream);
  const int H = (int)((N + block_rows - 1) / block_rows), W = (L + block_cols - 1) / block_cols;
  double* moments = reinterpret_cast<double*>(work_dev);
  float* limit = reinterpret_cast<float*>(moments + 4 * (size_t)T);
  MEPP_CUDA_CHECK(cudaMemsetAsync(moments, 0, sizeof(double) * 4 * (size_t)T, st));
  MEPP_CUDA_CHECK(cudaMemsetAsync(out_dev, 0, sizeof(float) * (size_t)T * H * W, st));
  const size_t smem = sizeof(double) * 3 * (size_t)T;
  MEPP_CUDA_CHECK(cudaFuncSetAttribute(heat_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  heat_moments_kernel<<<148 * 2, 256, smem, st>>>(reinterpret_cast<const HeatHit*>(hits_dev), hit_count_dev, hits_cap, T,
                                                  moments);
  heat_limits_kernel<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(moments, T, sigma, limit);
  heat_blockmax_kernel<<<148 * 4, 256, 0, st>>>(reinterpret_cast<const HeatHit*>(hits_dev), hit_count_dev, hits_cap, T,
                                                block_rows, block_cols, H, W, limit, reinterpret_cast<int*>(out_dev));
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"

// K1: PWM scan of 2-bit packed DNA, both Conv1D filters in one pass, threshold +
// ReLU, +-margin blur, fp16 hi/lo split of the blurred score written straight into
// the GEMM's tiled A operand, plus the per-position / per-sequence side outputs.
//
// Replaces (reference npdeloss/mepp): the Keras conv model of
// mepp/motif_layers.py:80-165 (normalise one-hot, Conv1D, reduce_max over the two
// filters, ZeroPadding1D((m-1)//2, ...), ReLU), the AveragePooling1D blur of
// mepp/core.py:59-145, the raw-moment sums over x of mepp/core.py:238-253 and the
// count / density reductions of mepp/core.py:415-470.
//
// Float32 summation order (declared canonical order, identical to oracle/profile.py):
//   acc = 0; for tap j ascending: acc = fl32(acc + W[base_j][j]);  an N base adds the four
//   quarter weights .25*W[A..T][j] one after the other;  s = fl32(acc + fl32(-thr)).
//
// Work decomposition: one CTA = one k-block (32 sequences) x 4 tasks; one warp = one
// task; one lane = one sequence, walking the positions left to right.  Lanes of a
// warp always look up the same tap, so a LUT read touches at most 4 distinct
// addresses in 4 distinct banks (conflict free).
#include "common.cuh"

namespace mepp {

struct HitRec { int32_t task, seq, pos; float x0; };

constexpr int kScanWarps = 4;
constexpr int kRowStride = 36;  // floats per position row of the x0 staging buffer (32 + 4 pad)

struct ScanArgs {
  const uint32_t* codes_T; const uint32_t* nmask_T; const float* zhat;
  int64_t N, n_pad, KB;
  int L, T, margin, Wc, Wm, strideC, strideM;
  const float* lut; const float* bias; const int32_t* mlen; const int32_t* nfilt;
  uint8_t* x_planes; double* rowstats; int32_t* count; int32_t* density;
  HitRec* hits; int64_t hits_cap; unsigned long long* hit_count;
};

template <bool DUAL>
__device__ __forceinline__ void scan_task(const ScanArgs& a, const int t, const int kb,
                                          const uint32_t* __restrict__ my_codes,
                                          const uint32_t* __restrict__ my_mask,
                                          float2* __restrict__ s_lut, float* __restrict__ s_buf,
                                          const float* __restrict__ s_z, const int lane) {
  const unsigned full = 0xffffffffu;
  const int L = a.L, mg = a.margin;
  const int m = a.mlen[t];
  const int padl = (m - 1) / 2;          // motif_layers.py:99-101
  const int n_valid = L - m + 1;         // 'valid' convolution outputs
  const float bias = a.bias[t];
  const int64_t n = (int64_t)kb * 32 + lane;
  const bool real = n < a.N;

  const float2* g_lut = reinterpret_cast<const float2*>(a.lut) + (size_t)t * (MEPP_MAX_MOTIF_LEN * 4);
  for (int i = lane; i < MEPP_MAX_MOTIF_LEN * 4; i += 32) s_lut[i] = g_lut[i];
  __syncwarp();

  const int m4 = (m + 3) & ~3;           // taps padded with zero weights (adding +0 is exact)
  const uint32_t mbits = m >= 32 ? 0xffffffffu : ((1u << m) - 1u);
  const int k = 1 + 2 * mg;
  const float kf = (float)k;
  const int nbuf = 32 + 2 * mg;
  const int kc = lane & 3, rsub = lane >> 2;

  int dens = 0;
  int loaded_cw = -1, loaded_mw = -1;
  uint32_t w0 = 0, w1 = 0, w2 = 0, mw0 = 0, mw1 = 0;

  for (int base = 0; base < L; base += 32) {
    int b_start = 0;
    if (base != 0) {
      // carry the last 2*margin rows of the previous chunk to the front
      for (int b = 0; b < 2 * mg; ++b) s_buf[b * kRowStride + lane] = s_buf[(32 + b) * kRowStride + lane];
      b_start = 2 * mg;
    }
    // ---- phase A: unsmoothed thresholded scores x0 for positions base-mg+b ----
    for (int b = b_start; b < nbuf; ++b) {
      const int p = base - mg + b;
      const int i = p - padl;            // window start
      float x0 = 0.0f;
      if (i >= 0 && i < n_valid) {       // warp uniform
        const int cw = i >> 4, mw = i >> 5;
        if (cw != loaded_cw) { w0 = my_codes[cw]; w1 = my_codes[cw + 1]; w2 = my_codes[cw + 2]; loaded_cw = cw; }
        if (mw != loaded_mw) { mw0 = my_mask[mw]; mw1 = my_mask[mw + 1]; loaded_mw = mw; }
        const uint32_t lo = __funnelshift_r(w0, w1, 2 * (i & 15));   // bases i .. i+15
        const uint32_t hi = __funnelshift_r(w1, w2, 2 * (i & 15));   // bases i+16 .. i+31
        const uint32_t nwin = __funnelshift_r(mw0, mw1, i & 31) & mbits;
        float a0 = 0.0f, a1 = 0.0f;
        for (int j0 = 0; j0 < m4; j0 += 4) {
          const uint32_t v = j0 < 16 ? (lo >> (2 * j0)) : (hi >> (2 * (j0 - 16)));
          const float2* l4 = s_lut + j0 * 4;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint32_t c = (v >> (2 * jj)) & 3u;
            const float2 w = l4[jj * 4 + c];
            a0 = __fadd_rn(a0, w.x);
            if (DUAL) a1 = __fadd_rn(a1, w.y);
          }
        }
        if (nwin != 0u) {
          // window contains N: redo in the canonical order with four quarter-weight adds per N tap
          a0 = 0.0f; a1 = 0.0f;
          for (int j = 0; j < m; ++j) {
            const uint32_t c = (j < 16 ? (lo >> (2 * j)) : (hi >> (2 * (j - 16)))) & 3u;
            const float2* l4 = s_lut + j * 4;
            if ((nwin >> j) & 1u) {
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) {
                a0 = __fadd_rn(a0, __fmul_rn(0.25f, l4[cc].x));
                if (DUAL) a1 = __fadd_rn(a1, __fmul_rn(0.25f, l4[cc].y));
              }
            } else {
              a0 = __fadd_rn(a0, l4[c].x);
              if (DUAL) a1 = __fadd_rn(a1, l4[c].y);
            }
          }
        }
        float s = __fadd_rn(a0, bias);
        if (DUAL) s = fmaxf(s, __fadd_rn(a1, bias));
        x0 = real ? fmaxf(s, 0.0f) : 0.0f;
        const bool hit = x0 > 0.0f;
        const unsigned bal = __ballot_sync(full, hit);
        if (bal != 0u) {
          const int nh = __popc(bal);
          unsigned long long slot0 = 0;
          if (lane == 0) {
            atomicAdd(&a.count[(int64_t)t * L + p], nh);
            if (a.hit_count) slot0 = atomicAdd(a.hit_count, (unsigned long long)nh);
          }
          if (a.hits) {
            slot0 = __shfl_sync(full, slot0, 0);
            if (hit) {
              const unsigned long long slot = slot0 + (unsigned)__popc(bal & ((1u << lane) - 1u));
              if ((int64_t)slot < a.hits_cap) {
                HitRec rec; rec.task = t; rec.seq = (int32_t)n; rec.pos = p; rec.x0 = x0;
                a.hits[slot] = rec;
              }
            }
          }
          dens += hit ? 1 : 0;
        }
      }
      s_buf[b * kRowStride + lane] = x0;
    }
    __syncwarp();
    // ---- phase B: blur, row sums, fp16 split, tiled store; lane = (kc, rsub) ----
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
      const int rl = rsub + 8 * it;
      const int p = base + rl;
      const bool valid = p < L;
      const bool interior = valid && p >= mg && p <= L - 1 - mg;   // core.py:79-91 zero edges
      float xb[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) xb[q] = 0.0f;
      if (interior) {
        const float* row = s_buf + rl * kRowStride + kc * 8;  // buffer row rl+d <-> position p-mg+d
        float4 u0 = *reinterpret_cast<const float4*>(row);
        float4 u1 = *reinterpret_cast<const float4*>(row + 4);
        xb[0] = u0.x; xb[1] = u0.y; xb[2] = u0.z; xb[3] = u0.w;
        xb[4] = u1.x; xb[5] = u1.y; xb[6] = u1.z; xb[7] = u1.w;
        for (int d = 1; d < k; ++d) {
          u0 = *reinterpret_cast<const float4*>(row + d * kRowStride);
          u1 = *reinterpret_cast<const float4*>(row + d * kRowStride + 4);
          xb[0] = __fadd_rn(xb[0], u0.x); xb[1] = __fadd_rn(xb[1], u0.y);
          xb[2] = __fadd_rn(xb[2], u0.z); xb[3] = __fadd_rn(xb[3], u0.w);
          xb[4] = __fadd_rn(xb[4], u1.x); xb[5] = __fadd_rn(xb[5], u1.y);
          xb[6] = __fadd_rn(xb[6], u1.z); xb[7] = __fadd_rn(xb[7], u1.w);
        }
        if (k > 1) {
#pragma unroll
          for (int q = 0; q < 8; ++q) xb[q] = __fdiv_rn(xb[q], kf);
        }
      }
      // row sums over this k-block's 32 sequences (float64), reduced over the 4 kc lanes
      double sx = 0.0, sxx = 0.0, sxz = 0.0;
      bool nz = false;
#pragma unroll
      for (int q = 0; q < 8; ++q) nz |= xb[q] != 0.0f;
      if (nz) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const double v = (double)xb[q];
          sx += v; sxx += v * v; sxz += v * (double)s_z[kc * 8 + q];
        }
      }
      if (__any_sync(full, nz)) {
        sx += __shfl_xor_sync(full, sx, 1);   sx += __shfl_xor_sync(full, sx, 2);
        sxx += __shfl_xor_sync(full, sxx, 1); sxx += __shfl_xor_sync(full, sxx, 2);
        sxz += __shfl_xor_sync(full, sxz, 1); sxz += __shfl_xor_sync(full, sxz, 2);
        if (kc == 0 && valid && sxx != 0.0) {
          double* rs = a.rowstats + ((int64_t)t * L + p) * 3;
          atomicAdd(rs + 0, sx); atomicAdd(rs + 1, sxx); atomicAdd(rs + 2, sxz);
        }
      }
      if (valid) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          __half h0, l0, h1, l1;
          split_half(xb[2 * q] * MEPP_X_SCALE, h0, l0);
          split_half(xb[2 * q + 1] * MEPP_X_SCALE, h1, l1);
          h[q] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          l[q] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        const int64_t r = (int64_t)t * L + p;
        uint8_t* tile = a.x_planes + a_tile_offset(r >> 7, kb, a.KB);
        const int row = (int)(r & 127);
        *reinterpret_cast<uint4*>(tile + a_chunk_offset(0, kc, row)) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(tile + a_chunk_offset(1, kc, row)) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
    __syncwarp();
  }
  a.density[(int64_t)t * a.n_pad + n] = real ? dens : 0;
}

__global__ void __launch_bounds__(kScanWarps * 32) scan_kernel(const ScanArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = blockIdx.x;
  const int nbuf = 32 + 2 * a.margin;
  // carve shared memory
  float* s_bufs = reinterpret_cast<float*>(smem_raw);                                   // [warps][nbuf][36]
  float2* s_luts = reinterpret_cast<float2*>(s_bufs + kScanWarps * nbuf * kRowStride);  // [warps][128]
  uint32_t* s_codes = reinterpret_cast<uint32_t*>(s_luts + kScanWarps * MEPP_MAX_MOTIF_LEN * 4);
  uint32_t* s_mask = s_codes + 32 * a.strideC;
  float* s_z = reinterpret_cast<float*>(s_mask + 32 * a.strideM);                       // [32]

  for (int idx = threadIdx.x; idx < a.Wc * 32; idx += blockDim.x) {
    const int s = idx & 31, w = idx >> 5;
    s_codes[s * a.strideC + w] = a.codes_T[(int64_t)w * a.n_pad + (int64_t)kb * 32 + s];
  }
  for (int idx = threadIdx.x; idx < a.Wm * 32; idx += blockDim.x) {
    const int s = idx & 31, w = idx >> 5;
    s_mask[s * a.strideM + w] = a.nmask_T[(int64_t)w * a.n_pad + (int64_t)kb * 32 + s];
  }
  if (threadIdx.x < 32) {
    const int64_t n = (int64_t)kb * 32 + threadIdx.x;
    s_z[threadIdx.x] = (a.zhat && n < a.N) ? a.zhat[n] : 0.0f;
  }
  __syncthreads();

  const int t = blockIdx.y * kScanWarps + warp;
  if (t >= a.T) return;
  float* s_buf = s_bufs + warp * nbuf * kRowStride;
  float2* s_lut = s_luts + warp * MEPP_MAX_MOTIF_LEN * 4;
  const uint32_t* my_codes = s_codes + lane * a.strideC;
  const uint32_t* my_mask = s_mask + lane * a.strideM;
  if (a.nfilt[t] == 2) scan_task<true>(a, t, kb, my_codes, my_mask, s_lut, s_buf, s_z, lane);
  else                 scan_task<false>(a, t, kb, my_codes, my_mask, s_lut, s_buf, s_z, lane);
}

}  // namespace mepp

extern "C" int mepp_scan(const uint32_t* codes_T_dev, const uint32_t* nmask_T_dev, const float* zhat_dev,
                         int64_t N, int L, int T, int margin, const float* lut_dev, const float* bias_dev,
                         const int32_t* mlen_dev, const int32_t* nfilt_dev, void* x_planes_dev,
                         double* rowstats_dev, int32_t* count_dev, int32_t* density_dev, void* hits_dev,
                         int64_t hits_cap, unsigned long long* hit_count_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(codes_T_dev && nmask_T_dev && lut_dev && bias_dev && mlen_dev && nfilt_dev &&
               x_planes_dev && rowstats_dev && count_dev && density_dev, "scan: null pointer");
  MEPP_REQUIRE(N > 0 && L > 0 && T > 0, "scan: empty input");
  MEPP_REQUIRE(margin >= 0 && margin <= MEPP_MAX_MARGIN, "scan: margin must be in [0, 16]");
  MEPP_REQUIRE(hits_dev == nullptr || hit_count_dev != nullptr, "scan: hits need a hit counter");
  MEPP_REQUIRE(mepp_device_ok(), "scan: no sm_100 CUDA device (no CPU fallback)");
  cudaStream_t st = as_stream(stream);
  ScanArgs a;
  a.codes_T = codes_T_dev; a.nmask_T = nmask_T_dev; a.zhat = zhat_dev;
  a.N = N; a.n_pad = mepp_pad32(N); a.KB = a.n_pad / kBlockK;
  a.L = L; a.T = T; a.margin = margin;
  a.Wc = (int)mepp_codes_words(L); a.Wm = (int)mepp_nmask_words(L);
  a.strideC = a.Wc | 1; a.strideM = a.Wm | 1;
  a.lut = lut_dev; a.bias = bias_dev; a.mlen = mlen_dev; a.nfilt = nfilt_dev;
  a.x_planes = reinterpret_cast<uint8_t*>(x_planes_dev); a.rowstats = rowstats_dev;
  a.count = count_dev; a.density = density_dev;
  a.hits = reinterpret_cast<HitRec*>(hits_dev); a.hits_cap = hits_cap; a.hit_count = hit_count_dev;
  const int nbuf = 32 + 2 * margin;
  size_t smem = sizeof(float) * kScanWarps * nbuf * kRowStride + sizeof(float2) * kScanWarps * MEPP_MAX_MOTIF_LEN * 4 +
                sizeof(uint32_t) * 32 * (a.strideC + a.strideM) + sizeof(float) * 32;
  MEPP_REQUIRE(smem <= 200 * 1024, "scan: sequence too long for the shared-memory staging (L too large)");
  MEPP_CUDA_CHECK(cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MEPP_CUDA_CHECK(cudaMemsetAsync(rowstats_dev, 0, sizeof(double) * 3 * (size_t)T * L, st));
  MEPP_CUDA_CHECK(cudaMemsetAsync(count_dev, 0, sizeof(int32_t) * (size_t)T * L, st));
  if (hit_count_dev) MEPP_CUDA_CHECK(cudaMemsetAsync(hit_count_dev, 0, sizeof(unsigned long long), st));
  dim3 grid((unsigned)a.KB, (unsigned)((T + kScanWarps - 1) / kScanWarps));
  MEPP_REQUIRE(grid.y <= 65535, "scan: too many tasks in one call");
  scan_kernel<<<grid, kScanWarps * 32, 0 + smem, st>>>(a);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

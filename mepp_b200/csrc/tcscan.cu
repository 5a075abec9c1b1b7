// K1a: first stage of the PWM scan on the 5th-generation tensor cores.
//
// The reference scores every window with a Keras Conv1D over the one-hot sequence
// (mepp/motif_layers.py:50-76, 80-165).  A convolution over a one-hot input IS a matrix product: with the flat
// one-hot stream F (4 bytes per base: channels A, C, G, T) the window that starts at base 4r + phi scores
//     sum_k F[16 r + k] * Wq[phi][k],      k = 4 j' + c,   Wq[phi][4 (j + phi) + c] = weight of letter c at tap j,
// so the rows of the A operand are OVERLAPPING 16-byte-strided views of one byte stream -- exactly what a K-major
// no-swizzle UMMA descriptor can address (8-row core matrices of 16-byte rows: SBO = 128 bytes; the next 16-byte K
// chunk of the same rows lies 16 bytes further: LBO = 16 bytes) -- and the four window phases are four shifted
// copies of the weights in the B operand.  One tcgen05.mma.kind::i8 (M = 128 windows, N = up to 256 columns =
// 32 scan units x 4 phases x 2 strands, K = 32 channels = 8 taps) replaces 128 x 256 table lookups of the
// CUDA-core pre-filter; the int32 accumulators stay in TMEM and only their SIGN is read back.
//
// Conservative integer weights (tc_weights_kernel): per filter, penalties pen_j(c) = max_c' W[c'][j] - W[c][j] >= 0
// are scaled by s = 127 / D, D = sum_j max_j - threshold + rounding slack, and floored: q = -min(128, floor(s pen)).
// A window can only pass the float32 threshold when its penalty is below D, i.e. sum_j q_j >= -126; 127 is folded
// into tap 0, a base contributes 4 q (one-hot value 4) and an N base sum_c q(c) (one-hot 1,1,1,1 -- the reference's N
// scores the column mean), so "accumulator >= 0" is a test with NO false negatives, the same for every column.
// Survivors (~3e-4 of the windows) are appended to per-warp candidate lists; scan.cu evaluates them in the
// canonical float32 order, so every reported score still comes from the exact path.
//
// Structure: persistent CTAs (one per SM), warp 0 = producer (cp.async.bulk of the weights once, then a 4-stage
// ring of 2 KB stream tiles), warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue (tcgen05.ld of the
// double-buffered 2 x 256-column accumulator, AND-reduction of the sign bits, candidate compaction).
#include "tcscan.cuh"
#include "tcgen05.cuh"

namespace mepp {

constexpr int kTcThreads = 320;
constexpr int kTcStages = 4;
constexpr int kTcStageBytes = 2304;   // 2048 + 32 * ks_max, ks_max <= 5 (motifs up to 32 taps + 3 phases)
constexpr int kTcAccCols = 256;

__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(kTcThreads, 1) tc_scan_kernel(const __grid_constant__ TcLaunch ln,
                                                                const __grid_constant__ TcScanArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b_pad = (ln.b_bytes + 1023u) & ~1023u;
  uint8_t* sB = smem;
  uint8_t* sA = smem + b_pad;
  uint64_t* afull = reinterpret_cast<uint64_t*>(sA + kTcStages * kTcStageBytes);   // [kTcStages]
  uint64_t* aempty = afull + kTcStages;                                            // [kTcStages]
  uint64_t* tfull = aempty + kTcStages;                                            // [2]
  uint64_t* tempty = tfull + 2;                                                    // [2]
  uint64_t* bfull = tempty + 2;                                                    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(smem_u32(&afull[s]), 1); mbar_init(smem_u32(&aempty[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tfull[s]), 1); mbar_init(smem_u32(&tempty[s]), kTcEpiWarps); }
    mbar_init(smem_u32(bfull), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(2u * kTcAccCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t a_bytes = 2048u + 32u * (uint32_t)ln.ks_max;

  if (warp == 0) {
    // ------------------------------------------------------------ producer
    if (lane == 0) {
      const uint32_t bb = smem_u32(bfull);
      mbar_expect_tx(bb, ln.b_bytes);
      for (uint32_t off = 0; off < ln.b_bytes; off += 32768u) {
        const uint32_t n = ln.b_bytes - off < 32768u ? ln.b_bytes - off : 32768u;
        bulk_g2s(smem_u32(sB + off), g.B + off, n, bb);
      }
      uint32_t stage = 0, phase = 0;
      for (int64_t tile = blockIdx.x; tile < g.n_mtiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&aempty[stage]), phase ^ 1u);
        const uint32_t fb = smem_u32(&afull[stage]);
        mbar_expect_tx(fb, a_bytes);
        bulk_g2s(smem_u32(sA + stage * kTcStageBytes), g.onehot + tile * 2048, a_bytes, fb);
        if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      mbar_wait(smem_u32(bfull), 0u);
      tc_fence_after();
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      const uint32_t sb0 = smem_u32(sB);
      for (int64_t tile = blockIdx.x; tile < g.n_mtiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&afull[stage]), phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(sA + stage * kTcStageBytes);
        for (int nt = 0; nt < ln.n_tiles; ++nt) {
          const TcTile tl = ln.tiles[nt];
          const uint32_t bn = tl.bn;
          // instruction descriptor: D = s32 (2 at [4,6)), A = u8 (0 at [7,10)), B = s8 (1 at [10,13)), K-major both,
          // N >> 3 at [17,23), M >> 4 at [24,29)
          const uint32_t idesc = (2u << 4) | (1u << 10) | ((bn >> 3) << 17) | ((uint32_t)(kTcRowsPerTile >> 4) << 24);
          mbar_wait(smem_u32(&tempty[acc]), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kTcAccCols;
          const uint32_t sb = sb0 + tl.b_off;
          for (uint32_t ks = 0; ks < tl.ksteps; ++ks) {
            // A: row r = bytes [16 r + 32 ks, +32) of the stream tile (overlapping rows): LBO 16, SBO 128
            const uint64_t a_desc = make_desc(sa + ks * 32u, 16u, 128u);
            const uint64_t b_desc = make_desc(sb + ks * 2u * bn * 16u, bn * 16u, 128u);
            tc_mma_i8(d_tmem, a_desc, b_desc, idesc, ks > 0 ? 1u : 0u);
          }
          tc_commit(smem_u32(&tfull[acc]));
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        tc_commit(smem_u32(&aempty[stage]));     // the stream tile is free once this tile's MMAs retire
        if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const unsigned full = 0xffffffffu;
    const int ew = warp - 2, q = warp & 3, h = ew >> 2;
    uint32_t acc = 0, acc_phase = 0, my_count = 0;
    const int64_t list = (int64_t)blockIdx.x * kTcEpiWarps + ew;
    uint2* const my_list = g.cand + list * (int64_t)g.list_cap;
    for (int64_t tile = blockIdx.x; tile < g.n_mtiles; tile += gridDim.x) {
      const uint32_t row = (uint32_t)(tile * kTcRowsPerTile) + (uint32_t)(q * 32 + lane);
      int dbg_col = 0;
      for (int nt = 0; nt < ln.n_tiles; ++nt) {
        const TcTile tl = ln.tiles[nt];
        mbar_wait(smem_u32(&tfull[acc]), acc_phase);
        tc_fence_after();
        const int n_chunks = tl.bn >> 5;
        for (int c = h; c < n_chunks; c += 2) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kTcAccCols + (uint32_t)(c * 32), v);
          if (g.dbg_acc != nullptr && tile == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) g.dbg_acc[(int64_t)(q * 32 + lane) * g.dbg_ld + dbg_col + c * 32 + j] = (int32_t)v[j];
          }
          uint32_t t = v[0];
#pragma unroll
          for (int j = 1; j < 32; ++j) t &= v[j];
          const bool has = (int32_t)t >= 0;           // some accumulator of this row is non-negative
          if (__any_sync(full, has)) {
            uint32_t cm = 0u;                        // bit p: (unit p / 4 of the chunk, phase p % 4) passes on a strand
            if (has) {
#pragma unroll
              for (int p = 0; p < 16; ++p)
                if ((int32_t)(v[2 * p] & v[2 * p + 1]) >= 0) cm |= 1u << p;
            }
            const int n = __popc(cm);
            int pre = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int x = __shfl_up_sync(full, pre, o);
              if (lane >= o) pre += x;
            }
            const int total = __shfl_sync(full, pre, 31);
            uint32_t slot = my_count + (uint32_t)(pre - n);
            my_count += (uint32_t)total;
            const uint32_t ubase = ((uint32_t)tl.unit0 + (uint32_t)c * 4u) << 2;
            while (cm != 0u) {
              const int p = __ffs((int)cm) - 1;
              cm &= cm - 1u;
              if (slot < g.list_cap) my_list[slot] = make_uint2(row, ubase + (uint32_t)p);
              ++slot;
            }
          }
        }
        dbg_col += tl.bn;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    if (lane == 0) g.cand_count[list] = my_count;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * kTcAccCols) : "memory");
  }
}

// One CTA per (N-tile, unit slot): the 8 columns (4 phases x 2 filters) of one scan unit, or never-passing columns for
// the padding slots of the tile.
__global__ void __launch_bounds__(128) tc_weights_kernel(const __grid_constant__ TcLaunch ln, const float* __restrict__ lut,
                                                         const float* __restrict__ bias, const int32_t* __restrict__ mlen,
                                                         const int32_t* __restrict__ nfilt, const int32_t* __restrict__ units,
                                                         int unit_base, int n_units_total, uint8_t* __restrict__ B) {
  __shared__ int s_q[2][MEPP_MAX_MOTIF_LEN][4];
  __shared__ int s_on[2];
  const TcTile tl = ln.tiles[blockIdx.x];
  const int slot = blockIdx.y;
  if (slot * 8 >= tl.bn) return;
  const int u = unit_base + tl.unit0 + slot;
  const bool valid = slot < tl.n_units && u < n_units_total;
  const int t = valid ? (units ? units[2 * u] : u) : -1;
  const int m = valid ? mlen[t] : 0;
  if (threadIdx.x < 2) {
    const int f = threadIdx.x;
    int on = 0;
    if (valid && f < nfilt[t]) {
      const float* l = lut + (size_t)t * (MEPP_MAX_MOTIF_LEN * 8);
      double summax = 0.0, absmax = 0.0;
      for (int j = 0; j < m; ++j) {
        double mx = -1e300, am = 0.0;
        for (int c = 0; c < 4; ++c) {
          const double w = (double)l[(j * 4 + c) * 2 + f];
          mx = fmax(mx, w); am = fmax(am, fabs(w));
        }
        summax += mx; absmax += am;
      }
      // |float32 sequential sum - real sum| < delta (host.cu: build_filter_table uses the same slack)
      const double delta = 2.0 * (4.0 * m) * absmax * 5.97e-8 + 1e-6;
      const double D = summax + (double)bias[t] + delta;           // bias = float32(-threshold)
      if (D > 0.0) {
        on = 1;
        const double s = 127.0 / D * (1.0 - 1e-9);
        for (int j = 0; j < m; ++j) {
          double mx = -1e300;
          for (int c = 0; c < 4; ++c) mx = fmax(mx, (double)l[(j * 4 + c) * 2 + f]);
          for (int c = 0; c < 4; ++c) {
            const double p = floor(s * (mx - (double)l[(j * 4 + c) * 2 + f]));
            const int qq = p >= 128.0 ? 128 : (int)p;
            s_q[f][j][c] = (j == 0 ? 127 : 0) - qq;
          }
        }
      }
    }
    s_on[f] = on;
  }
  __syncthreads();
  const int K = 32 * tl.ksteps, bn = tl.bn;
  for (int idx = threadIdx.x; idx < 8 * K; idx += blockDim.x) {
    const int col8 = idx / K, k = idx - col8 * K;
    const int phi = col8 >> 1, f = col8 & 1;
    const int jp = k >> 2, c = k & 3, j = jp - phi;
    int val = 0;
    if (s_on[f]) { if (j >= 0 && j < m) val = s_q[f][j][c]; }
    else if (jp == 0) val = -1;                                   // never reaches zero on a real base
    const int n = slot * 8 + col8, ks = k >> 5, kc = (k >> 4) & 1, e = k & 15;
    B[(size_t)tl.b_off + (size_t)((ks * 2 + kc) * bn + n) * 16 + e] = (uint8_t)(int8_t)val;
  }
}

int tc_scan_grid() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

int tc_build_weights(const TcLaunch& ln, const float* lut_dev, const float* bias_dev, const int32_t* mlen_dev,
                     const int32_t* nfilt_dev, const int32_t* units_dev, int unit_base, int n_units_total, uint8_t* B_dev,
                     cudaStream_t st) {
  dim3 grid((unsigned)ln.n_tiles, 32u);
  tc_weights_kernel<<<grid, 128, 0, st>>>(ln, lut_dev, bias_dev, mlen_dev, nfilt_dev, units_dev, unit_base, n_units_total,
                                          B_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tc_scan_launch(const TcLaunch& ln, const TcScanArgs& a, cudaStream_t st) {
  const size_t smem = ((size_t)ln.b_bytes + 1023) / 1024 * 1024 + (size_t)kTcStages * kTcStageBytes + 256;
  MEPP_REQUIRE(ln.b_bytes <= (uint32_t)kTcMaxBBytes && ln.ks_max >= 1 && 2048 + 32 * ln.ks_max <= kTcStageBytes,
               "scan_tc: weight operand or motif length outside the kernel's limits");
  MEPP_CUDA_CHECK(cudaFuncSetAttribute(tc_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = tc_scan_grid();
  if ((int64_t)grid > a.n_mtiles) grid = (int)a.n_mtiles;
  tc_scan_kernel<<<grid, kTcThreads, smem, st>>>(ln, a);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ASCII -> flat one-hot stream (4 bytes per base), zero words behind the last sequence.  Same letter rule as the
// 3-bit symbols (pack.cu: upper-case A, C, G, T; everything else is N = 1,1,1,1 -- onehot_dna.py:3-30).
__global__ void onehot_stream_kernel(const uint8_t* __restrict__ ascii, int64_t n_bases, int64_t n_words,
                                     uint32_t* __restrict__ out) {
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i0 >= n_words) return;
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = i0 + k;
    uint32_t x = 0u;
    if (i < n_bases) {
      const uint8_t ch = ascii[i];
      x = ch == 'A' ? 0x00000004u : ch == 'C' ? 0x00000400u : ch == 'G' ? 0x00040000u : ch == 'T' ? 0x04000000u : 0x01010101u;
    }
    w[k] = x;
  }
  *reinterpret_cast<uint4*>(out + i0) = make_uint4(w[0], w[1], w[2], w[3]);
}

}  // namespace mepp

extern "C" int64_t mepp_onehot_words(int64_t N, int L) {
  const int64_t tiles = (N * (int64_t)L + mepp::kTcBasesPerTile - 1) / mepp::kTcBasesPerTile;
  return tiles * mepp::kTcBasesPerTile + 128;      // halo behind the last tile: every bulk copy stays inside
}

extern "C" int mepp_pack_onehot(const uint8_t* ascii_dev, int64_t N, int L, uint32_t* onehot_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(ascii_dev && onehot_dev && N > 0 && L > 0, "pack_onehot: bad arguments");
  MEPP_REQUIRE(mepp_device_ok(), "pack_onehot: no sm_100 CUDA device (no CPU fallback)");
  const int64_t n_words = mepp_onehot_words(N, L);
  const int64_t n_thr = n_words / 4;
  onehot_stream_kernel<<<(unsigned)((n_thr + 255) / 256), 256, 0, as_stream(stream)>>>(ascii_dev, N * (int64_t)L, n_words,
                                                                                    onehot_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// K2: the positional profile contraction S[r][c] = sum_n X[r][n] * Y[n][c]
// (rows r = task*L + position, columns c = true score + P permuted scores,
// reduction over the N sequences) on the 5th-generation tensor cores.
//
// Replaces the reference's reduce_sum(x*y, 0) over a materialised (B, L, 1+P)
// product, mepp/core.py:250 (npdeloss/mepp), which is this GEMM computed without
// one.
//
// Numerics: both operands are fp16 hi+lo splits (hi + lo carries ~22 mantissa
// bits); three tcgen05.mma products hi*hi + hi*lo + lo*hi accumulate in fp32 in
// TMEM (the dropped lo*lo term is < 2^-22 relative).  The TMEM accumulator is drained
// every kChunkKB k-blocks (1024 sequences) into fp32 registers of the epilogue warps and
// restarted, so the fp32 accumulation error depends on the matches inside one chunk, not
// on the number of sequences.
//
// Structure: persistent CTAs, one 128 x BN output tile at a time.
//   warp 0   : producer -- one elected lane issues cp.async.bulk (TMA engine, 1-D
//              bulk copies of pre-tiled operand blocks) into a 4-stage smem ring,
//              completion on mbarriers
//   warp 1   : TMEM allocator + MMA issuer -- one lane issues tcgen05.mma
//              (cta_group::1, kind::f16, M=128, N=BN, K=16), tcgen05.commit frees
//              smem stages and publishes finished accumulators
//   warps 2-9: epilogue -- tcgen05.ld each finished chunk of the fp32 accumulator
//              (double buffered in TMEM: 2 x 256 columns, so draining chunk i overlaps
//              the MMAs of chunk i+1), add it into registers (128 columns per thread),
//              store S after the last chunk
// Operand tiles use the no-swizzle K-major canonical layout (8x16-byte core
// matrices, LBO = stride between the two 16-byte K chunks of one MMA, SBO = stride
// between 8-row groups), produced directly in that layout by scan.cu / pack.cu.
#include "common.cuh"
#include "tcgen05.cuh"
#include <cstdlib>
#include <cfloat>

namespace mepp {

constexpr int kStages = 4;
constexpr int kGemmThreads = 320;   // producer, MMA issuer, 8 epilogue warps
constexpr int kChunkKB = 32;        // k-blocks (of 32 sequences) per TMEM accumulation chunk
constexpr int kAccCols = 256;  // TMEM columns per accumulator buffer

struct GemmArgs {
  const uint8_t* A; const uint8_t* B; float* S;
  const uint32_t* flags;    // non-zero bit array [n_tiles_m][FW] behind the A tiles (common.cuh)
  int64_t m_rows, ldS, KB;
  int BN, n_tiles_m, n_tiles_n, variant;
  int FW, use_flags;        // use_flags = 0: treat every block as non-zero
  // fused correlation epilogue (stats.cu: col_constants_kernel): when colconst != nullptr the epilogue writes
  // r = nan_to_num(partial correlation) into S (as float [m_rows][ldS]) instead of the raw sums
  const double* rowstats; const double* colconst; int C;
  const float* xscale; int L;   // per-task scale of the A operand (row / L = task); null: MEPP_X_SCALE
};

// Does k-block kb of the current item have to be loaded and multiplied?  Yes if one of its two K=16
// halves is flagged non-zero -- or if it is the last k-block of an accumulation chunk in which
// nothing is flagged (every chunk writes its accumulator at least once, so the epilogue needs no flags).
__device__ __forceinline__ uint32_t kb_bits(const uint32_t* f, int64_t kb) {
  return (f[(2 * kb) >> 5] >> ((2 * kb) & 31)) & 3u;
}

// {K=16 steps issued, K=16 steps total} summed over all launches since the last reset (bench.py reports
// the executed tensor rate and the fraction of all-zero steps that were skipped from these).
__device__ unsigned long long g_mma_steps[2];

__global__ void __launch_bounds__(kGemmThreads, 1) profile_gemm_kernel(const GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = g.BN;
  const uint32_t a_bytes = kATileBytes;
  const uint32_t b_bytes = (uint32_t)b_tile_bytes(BN);
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* bar_area = smem + (size_t)kStages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_area);       // [kStages]
  uint64_t* empty_bar = full_bar + kStages;                         // [kStages]
  uint64_t* tfull_bar = empty_bar + kStages;                        // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                             // [2]
  uint64_t* ffull_bar = tempty_bar + 2;                             // [2] item's flag words staged in smem
  uint64_t* fempty_bar = ffull_bar + 2;                             // [2] MMA issuer done with them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fempty_bar + 2);
  uint32_t* s_flags = tmem_slot + 4;                                // [2][FW]
  const int FW = g.FW;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tfull_bar[s]), 1); mbar_init(smem_u32(&tempty_bar[s]), 8); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&ffull_bar[s]), 1); mbar_init(smem_u32(&fempty_bar[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(2u * kAccCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int64_t n_items = (int64_t)g.n_tiles_m * g.n_tiles_n;
  const int64_t KB = g.KB;

  if (warp == 0) {
    // ------------------------------------------------------------ producer
    // all lanes stage the item's flag words in shared memory, lane 0 then issues the bulk copies of the
    // k-blocks that are not all-zero
    uint32_t stage = 0, phase = 0, fbuf = 0, fphase = 0;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int64_t mt = item / g.n_tiles_n, nt = item % g.n_tiles_n;
      const uint32_t* f = s_flags + fbuf * FW;
      if (g.use_flags) {
        if (lane == 0) mbar_wait(smem_u32(&fempty_bar[fbuf]), fphase ^ 1u);
        __syncwarp();
        for (int w = lane; w < FW; w += 32) s_flags[fbuf * FW + w] = g.flags[mt * FW + w];
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ffull_bar[fbuf]));
      }
      if (lane == 0) {
        const uint8_t* a_src = g.A + a_tile_offset(mt, 0, KB);
        const uint8_t* b_src = g.B + b_tile_offset(nt, 0, KB, BN);
        for (int64_t kb0 = 0; kb0 < KB; kb0 += kChunkKB) {
          const int64_t kb1 = kb0 + kChunkKB < KB ? kb0 + kChunkKB : KB;
          const bool chunk_nz = !g.use_flags || (f[(2 * kb0) >> 5] | f[((2 * kb0) >> 5) + 1]) != 0u;
          for (int64_t kb = kb0; kb < kb1; ++kb) {
            if (g.use_flags && kb_bits(f, kb) == 0u && !(kb == kb1 - 1 && !chunk_nz)) continue;
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            mbar_expect_tx(fb, stage_bytes);
            const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            bulk_g2s(sa, a_src + kb * (int64_t)a_bytes, a_bytes, fb);
            bulk_g2s(sa + a_bytes, b_src + kb * (int64_t)b_bytes, b_bytes, fb);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
      __syncwarp();
      if (++fbuf == 2) { fbuf = 0; fphase ^= 1u; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      // instruction descriptor: D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
      const uint32_t a_lbo = g.variant == 1 ? 128u : (uint32_t)kBlockM * 16u;
      const uint32_t a_sbo = g.variant == 1 ? (uint32_t)kBlockM * 16u : 128u;
      const uint32_t b_lbo = g.variant == 1 ? 128u : (uint32_t)BN * 16u;
      const uint32_t b_sbo = g.variant == 1 ? (uint32_t)BN * 16u : 128u;
      const uint32_t a_plane = kChunks * kBlockM * 16, b_plane = (uint32_t)kChunks * BN * 16;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, fbuf = 0, fphase = 0;
      unsigned long long issued = 0, total = 0;
      for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        total += (unsigned long long)KB * (kBlockK / 16);
        const uint32_t* f = s_flags + fbuf * FW;
        if (g.use_flags) mbar_wait(smem_u32(&ffull_bar[fbuf]), fphase);
        for (int64_t kb0 = 0; kb0 < KB; kb0 += kChunkKB) {
          const int64_t kb1 = kb0 + kChunkKB < KB ? kb0 + kChunkKB : KB;
          const bool chunk_nz = !g.use_flags || (f[(2 * kb0) >> 5] | f[((2 * kb0) >> 5) + 1]) != 0u;
          mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kAccCols;
          bool acc_valid = false;                      // has this chunk's accumulator been written yet?
          for (int64_t kb = kb0; kb < kb1; ++kb) {
            // bit s set <=> K=16 step s of this k-block holds a non-zero x (scan.cu); all-zero blocks were not loaded
            const uint32_t nzmask = g.use_flags ? kb_bits(f, kb) : 3u;
            if (nzmask == 0u && !(kb == kb1 - 1 && !chunk_nz)) continue;
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint32_t sb = sa + a_bytes;
#pragma unroll
            for (int s = 0; s < kBlockK / 16; ++s) {
              const bool last = (kb == kb1 - 1) && (s == kBlockK / 16 - 1);
              if (!((nzmask >> s) & 1u) && !(last && !acc_valid)) continue;   // (an untouched chunk still needs one write)
              const uint32_t a_off = (uint32_t)s * 2u * kBlockM * 16u, b_off = (uint32_t)s * 2u * BN * 16u;
              const uint64_t a_hi = make_desc(sa + a_off, a_lbo, a_sbo);
              const uint64_t a_lo = make_desc(sa + a_plane + a_off, a_lbo, a_sbo);
              const uint64_t b_hi = make_desc(sb + b_off, b_lbo, b_sbo);
              const uint64_t b_lo = make_desc(sb + b_plane + b_off, b_lbo, b_sbo);
              tc_mma_f16(d_tmem, a_hi, b_hi, idesc, acc_valid ? 1u : 0u);
              tc_mma_f16(d_tmem, a_hi, b_lo, idesc, 1u);
              tc_mma_f16(d_tmem, a_lo, b_hi, idesc, 1u);
              acc_valid = true;
              ++issued;
            }
            tc_commit(smem_u32(&empty_bar[stage]));   // smem stage reusable once these MMAs retire
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          tc_commit(smem_u32(&tfull_bar[acc]));       // chunk accumulator complete
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (g.use_flags) mbar_arrive(smem_u32(&fempty_bar[fbuf]));
        if (++fbuf == 2) { fbuf = 0; fphase ^= 1u; }
      }
      atomicAdd(&g_mma_steps[0], issued);
      atomicAdd(&g_mma_steps[1], total);
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int h = (warp - 2) >> 2;     // which 128-column half of the accumulator this warp drains
    uint32_t acc = 0, acc_phase = 0;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int64_t mt = item / g.n_tiles_n, nt = item % g.n_tiles_n;
      float sum[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) sum[j] = 0.0f;
      for (int64_t kb0 = 0; kb0 < KB; kb0 += kChunkKB) {
        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          if (h * 128 + c0 < BN) {     // warp uniform
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccCols + (uint32_t)(h * 128 + c0), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[c0 + j] = __fadd_rn(sum[c0 + j], __uint_as_float(v[j]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      const int64_t row = mt * kBlockM + q * 32 + lane;
      if (row < g.m_rows) {
        float* dst = g.S + row * g.ldS + nt * (int64_t)BN + h * 128;
        if (g.colconst == nullptr) {
#pragma unroll
          for (int c0 = 0; c0 < 128; c0 += 4) {
            if (h * 128 + c0 < BN)
              *reinterpret_cast<float4*>(dst + c0) = make_float4(sum[c0], sum[c0 + 1], sum[c0 + 2], sum[c0 + 3]);
          }
        } else {
          // fused correlation epilogue: core.py:182-193,255-265 in float64 + np.nan_to_num (core.py:602)
          const double* cc = g.colconst;
          const double n = cc[0], sy = cc[1], ay = cc[2], sz = cc[3], az = cc[4];
          const bool use_z = cc[5] != 0.0;
          const double sx = g.rowstats[row * 3 + 0], sxx = g.rowstats[row * 3 + 1], sxz = g.rowstats[row * 3 + 2];
          const double ax = n * sxx - sx * sx;
          const double inv_xy = 1.0 / sqrt(ax * ay);
          const double sxsy = sx * sy;
          double r_xz = 0.0, row_den = 1.0;
          if (use_z) {
            r_xz = (n * sxz - sx * sz) / sqrt(ax * az);
            row_den = 1.0 / sqrt(1.0 - r_xz * r_xz);
          }
          const int col0 = (int)(nt * (int64_t)BN) + h * 128;
          const double inv_xs = g.xscale ? 1.0 / (double)__ldg(g.xscale + row / g.L) : 1.0 / (double)MEPP_X_SCALE;
#pragma unroll
          for (int c0 = 0; c0 < 128; c0 += 4) {
            if (h * 128 + c0 < BN) {
              float o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int c = col0 + c0 + j;
                float out = 0.0f;
                if (c < g.C) {
                  const double sxy = (double)sum[c0 + j] * inv_xs;
                  double r = (n * sxy - sxsy) * inv_xy;
                  if (use_z) r = (r - r_xz * __ldg(cc + 8 + 2 * c)) * row_den * __ldg(cc + 9 + 2 * c);
                  out = (float)r;
                  if (out != out) out = 0.0f;
                  else if (isinf(out)) out = out > 0 ? FLT_MAX : -FLT_MAX;
                }
                o[j] = out;
              }
              *reinterpret_cast<float4*>(dst + c0) = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * kAccCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// EXPERIMENTAL, OFF BY DEFAULT (MEPP_GEMM_2CTA=1): measured SLOWER than the 1-CTA kernel above on
// B200 (cfg 2: 31.7 ms vs 21.2 ms; tensor pipe 57 % active vs 99 %, profiles/r01_gemm_2cta_*), and
// insensitive to the ring depth (3/4/6 stages identical) -- the pair's shared B operand path, not
// staging, is the limit.  Kept because it is parity-green and documents the experiment.
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x BN tile.
// Each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 columns), so the tensor
// pipe's shared-memory operand traffic per SM drops from 12 KB to 8 KB per MMA and the ring
// holds 6 stages.  The leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 (M = 256);
// its commits are multicast to both CTAs' barriers.  The peer's bulk copies complete on the
// peer's own full barrier; one peer thread forwards each completion to the leader with a remote
// mbarrier arrive.  Each CTA drains its own 128 x BN accumulator half from its own TMEM.
constexpr int kStages2 = 6;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void tc_commit2_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_mma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
profile_gemm2_kernel(const GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();            // 0 = leader (issues the MMAs)
  const int BN = g.BN, BH = BN / 2;                   // each CTA stages BH columns of B
  const uint32_t a_bytes = kATileBytes;
  const uint32_t bh_chunk = (uint32_t)BH * 16u;       // one (plane, kc) slab of this CTA's B half
  const uint32_t b_bytes = 2u * kChunks * bh_chunk;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t n_stages = (uint32_t)g.variant;      // ring depth (runtime: 2..kStages2)
  uint8_t* bar_area = smem + (size_t)n_stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_area);       // [kStages2] own bulk copies landed
  uint64_t* pfull_bar = full_bar + n_stages;                        // [kStages2] (leader) peer's stage landed
  uint64_t* empty_bar = pfull_bar + n_stages;                       // [kStages2] MMAs of the stage retired
  uint64_t* tfull_bar = empty_bar + n_stages;                       // [2] accumulator chunk complete
  uint64_t* tempty_bar = tfull_bar + 2;                             // [2] (leader) both CTAs drained the chunk
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < n_stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&pfull_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tfull_bar[s]), 1); mbar_init(smem_u32(&tempty_bar[s]), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(2u * kAccCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // both CTAs' barriers exist before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int n_pairs_m = (g.n_tiles_m + 1) / 2;
  const int64_t n_items = (int64_t)n_pairs_m * g.n_tiles_n;
  const int64_t KB = g.KB;
  const int64_t first = cluster_id_x(), step = num_clusters_x();

  if (warp == 0) {
    // ------------------------------------------------------------ producer (both CTAs)
    // lane 0 arms the barrier and copies the A tile, lanes 1..8 copy one (plane, kc) slab of the B half each
    // (nine bulk copies per stage from one lane would make the producer issue-bound)
    if (lane < 1 + 2 * kChunks) {
      uint32_t stage = 0, phase = 0;
      for (int64_t item = first; item < n_items; item += step) {
        const int64_t mp = item / g.n_tiles_n, nt = item % g.n_tiles_n;
        int64_t mt = 2 * mp + rank;
        if (mt >= g.n_tiles_m) mt = g.n_tiles_m - 1;   // odd tile count: the peer re-reads a valid tile, stores nothing
        const uint8_t* a_src = g.A + a_tile_offset(mt, 0, KB);
        const uint8_t* b_src = g.B + b_tile_offset(nt, 0, KB, BN) + (size_t)rank * bh_chunk;
        const uint32_t full_tile_chunk = (uint32_t)BN * 16u;
        for (int64_t kb = 0; kb < KB; ++kb) {
          mbar_wait_cluster(smem_u32(&empty_bar[stage]), phase ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          if (lane == 0) {
            mbar_expect_tx(fb, stage_bytes);
            bulk_g2s(sa, a_src + kb * (int64_t)a_bytes, a_bytes, fb);
          } else {
            const int c = lane - 1;
            bulk_g2s(sa + a_bytes + (uint32_t)c * bh_chunk,
                     b_src + kb * (int64_t)b_tile_bytes(BN) + (size_t)c * full_tile_chunk, bh_chunk, fb);
          }
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 1) {
      // ---------------------------------------------------------- peer: forward "stage landed" to the leader
      uint32_t stage = 0, phase = 0;
      for (int64_t item = first; item < n_items; item += step) {
        for (int64_t kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          mbar_arrive_remote(map_to_cta(smem_u32(&pfull_bar[stage]), 0));
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (lane == 0) {
      // ---------------------------------------------------------- leader: MMA issuer
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * kBlockM) >> 4) << 24);
      const uint32_t a_lbo = (uint32_t)kBlockM * 16u, b_lbo = bh_chunk, sbo = 128u;
      const uint32_t a_plane = kChunks * kBlockM * 16, b_plane = (uint32_t)kChunks * bh_chunk;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int64_t item = first; item < n_items; item += step) {
        for (int64_t kb0 = 0; kb0 < KB; kb0 += kChunkKB) {
          const int64_t kb1 = kb0 + kChunkKB < KB ? kb0 + kChunkKB : KB;
          mbar_wait_cluster(smem_u32(&tempty_bar[acc]), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kAccCols;
          for (int64_t kb = kb0; kb < kb1; ++kb) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            mbar_wait_cluster(smem_u32(&pfull_bar[stage]), phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint32_t sb = sa + a_bytes;
#pragma unroll
            for (int s = 0; s < kBlockK / 16; ++s) {
              const uint32_t a_off = (uint32_t)s * 2u * a_lbo, b_off = (uint32_t)s * 2u * b_lbo;
              const uint64_t a_hi = make_desc(sa + a_off, a_lbo, sbo);
              const uint64_t a_lo = make_desc(sa + a_plane + a_off, a_lbo, sbo);
              const uint64_t b_hi = make_desc(sb + b_off, b_lbo, sbo);
              const uint64_t b_lo = make_desc(sb + b_plane + b_off, b_lbo, sbo);
              tc_mma2_f16(d_tmem, a_hi, b_hi, idesc, (kb > kb0 || s > 0) ? 1u : 0u);
              tc_mma2_f16(d_tmem, a_hi, b_lo, idesc, 1u);
              tc_mma2_f16(d_tmem, a_lo, b_hi, idesc, 1u);
            }
            tc_commit2_mc(smem_u32(&empty_bar[stage]), (uint16_t)3);   // frees the stage in both CTAs
            if (++stage == n_stages) { stage = 0; phase ^= 1u; }
          }
          tc_commit2_mc(smem_u32(&tfull_bar[acc]), (uint16_t)3);       // chunk complete, both CTAs
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    uint32_t acc = 0, acc_phase = 0;
    for (int64_t item = first; item < n_items; item += step) {
      const int64_t mp = item / g.n_tiles_n, nt = item % g.n_tiles_n;
      const int64_t mt = 2 * mp + rank;
      float sum[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) sum[j] = 0.0f;
      for (int64_t kb0 = 0; kb0 < KB; kb0 += kChunkKB) {
        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          if (h * 128 + c0 < BN) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccCols + (uint32_t)(h * 128 + c0), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[c0 + j] = __fadd_rn(sum[c0 + j], __uint_as_float(v[j]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(map_to_cta(smem_u32(&tempty_bar[acc]), 0));   // leader's barrier
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      const int64_t row = mt * kBlockM + q * 32 + lane;
      if (mt < g.n_tiles_m && row < g.m_rows) {
        float* dst = g.S + row * g.ldS + nt * (int64_t)BN + h * 128;
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 4) {
          if (h * 128 + c0 < BN)
            *reinterpret_cast<float4*>(dst + c0) = make_float4(sum[c0], sum[c0 + 1], sum[c0 + 2], sum[c0 + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // nobody leaves while the pair still signals / reads smem
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * kAccCols) : "memory");
  }
}

// Cross-check: the same contraction from the same fp16 planes on CUDA cores with
// float64 accumulation (all four hi/lo products), for selected rows.
__global__ void gemm_check_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B,
                                  double* __restrict__ S, int64_t m_rows, int C, int64_t KB, int BN,
                                  const int32_t* __restrict__ rows, int n_rows_sel) {
  const int ri = blockIdx.y;
  const int64_t r = rows ? rows[ri] : ri;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m_rows || c >= C || ri >= n_rows_sel) return;
  const int64_t mt = r >> 7; const int row = (int)(r & 127);
  const int nt = c / BN, col = c % BN;
  double acc = 0.0;
  for (int64_t kb = 0; kb < KB; ++kb) {
    const uint8_t* at = A + a_tile_offset(mt, kb, KB);
    const uint8_t* bt = B + b_tile_offset(nt, kb, KB, BN);
    for (int kc = 0; kc < kChunks; ++kc) {
      const __half* ah = reinterpret_cast<const __half*>(at + a_chunk_offset(0, kc, row));
      const __half* al = reinterpret_cast<const __half*>(at + a_chunk_offset(1, kc, row));
      const __half* bh = reinterpret_cast<const __half*>(bt + b_chunk_offset(0, kc, col, BN));
      const __half* bl = reinterpret_cast<const __half*>(bt + b_chunk_offset(1, kc, col, BN));
      for (int e = 0; e < 8; ++e) {
        const double x = (double)__half2float(ah[e]) + (double)__half2float(al[e]);
        if (x != 0.0) acc += x * ((double)__half2float(bh[e]) + (double)__half2float(bl[e]));
      }
    }
  }
  S[(int64_t)ri * C + c] = acc;
}

}  // namespace mepp



static int launch_profile_gemm(const void* x_planes_dev, const void* y_planes_dev, float* out_dev, int64_t m_rows,
                               int C, int64_t n_pad, const double* rowstats_dev, const double* colconst_dev,
                               const float* xscale_dev, int L, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(x_planes_dev && y_planes_dev && out_dev, "profile_gemm: null pointer");
  MEPP_REQUIRE(m_rows > 0 && C > 0 && n_pad > 0 && n_pad % kBlockK == 0, "profile_gemm: bad shape");
  MEPP_REQUIRE(mepp_device_ok(), "profile_gemm: no sm_100 CUDA device (no CPU fallback)");
  GemmArgs g;
  g.A = reinterpret_cast<const uint8_t*>(x_planes_dev);
  g.B = reinterpret_cast<const uint8_t*>(y_planes_dev);
  g.S = out_dev;
  g.m_rows = m_rows;
  g.KB = n_pad / kBlockK;
  g.BN = mepp_choose_bn(C, &g.n_tiles_n);
  g.ldS = (int64_t)g.n_tiles_n * g.BN;
  g.n_tiles_m = (int)((m_rows + kBlockM - 1) / kBlockM);
  g.rowstats = rowstats_dev; g.colconst = colconst_dev; g.C = C;
  g.xscale = xscale_dev; g.L = L > 0 ? L : 1;
  const char* var = getenv("MEPP_GEMM_DESC_VARIANT");
  g.variant = var ? atoi(var) : 0;
  g.FW = (int)flag_words(g.KB);
  g.flags = reinterpret_cast<const uint32_t*>(g.A + a_flags_offset(g.n_tiles_m, g.KB));
  g.use_flags = (g.FW <= 2048 && g.variant != 2) ? 1 : 0;   // MEPP_GEMM_DESC_VARIANT=2: never skip (A/B testing)
  size_t smem = (size_t)kStages * (kATileBytes + b_tile_bytes(g.BN)) + 256 + (g.use_flags ? 2 * 4 * (size_t)g.FW : 0);
  MEPP_CUDA_CHECK(cudaFuncSetAttribute(profile_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  MEPP_CUDA_CHECK(cudaGetDevice(&dev));
  MEPP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int64_t n_items = (int64_t)g.n_tiles_m * g.n_tiles_n;
  unsigned grid = (unsigned)(n_items < sms ? n_items : sms);
  const char* two = getenv("MEPP_GEMM_2CTA");
  if (two && atoi(two) != 0 && g.BN % 32 == 0 && colconst_dev == nullptr) {
    const char* stg = getenv("MEPP_GEMM_STAGES");
    int n_stages = stg ? atoi(stg) : kStages2;
    if (n_stages < 2) n_stages = 2;
    if (n_stages > kStages2) n_stages = kStages2;
    g.variant = n_stages;
    size_t smem2 = (size_t)n_stages * (kATileBytes + b_tile_bytes(g.BN) / 2) + 512;
    MEPP_CUDA_CHECK(cudaFuncSetAttribute(profile_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    int64_t n_items2 = (int64_t)((g.n_tiles_m + 1) / 2) * g.n_tiles_n;
    int64_t pairs = sms / 2;
    if (n_items2 < pairs) pairs = n_items2;
    profile_gemm2_kernel<<<(unsigned)(2 * pairs), kGemmThreads, smem2, as_stream(stream)>>>(g);
    MEPP_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  profile_gemm_kernel<<<grid, kGemmThreads, smem, as_stream(stream)>>>(g);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" {

int mepp_profile_gemm(const void* x_planes_dev, const void* y_planes_dev, float* S_dev, int64_t m_rows,
                      int C, int64_t n_pad, void* stream) {
  return launch_profile_gemm(x_planes_dev, y_planes_dev, S_dev, m_rows, C, n_pad, nullptr, nullptr, nullptr, 0, stream);
}

int mepp_profile_corr_gemm(const void* x_planes_dev, const void* y_planes_dev, const double* rowstats_dev,
                           const double* colconst_dev, float* r_dev, int64_t m_rows, int C, int64_t n_pad,
                           const float* xscale_dev, int L, void* stream) {
  if (!rowstats_dev || !colconst_dev) return mepp::fail("mepp_b200: profile_corr_gemm: null pointer");
  if (xscale_dev && (L <= 0 || m_rows % L != 0))
    return mepp::fail("mepp_b200: profile_corr_gemm: per-task scales need L with m_rows = T * L");
  return launch_profile_gemm(x_planes_dev, y_planes_dev, r_dev, m_rows, C, n_pad, rowstats_dev, colconst_dev, xscale_dev, L,
                             stream);
}

int mepp_gemm_counters(unsigned long long* out2, int reset) {
  using namespace mepp;
  MEPP_REQUIRE(mepp_device_ok(), "gemm_counters: no sm_100 CUDA device (no CPU fallback)");
  if (out2) MEPP_CUDA_CHECK(cudaMemcpyFromSymbol(out2, g_mma_steps, sizeof(unsigned long long) * 2));
  if (reset) {
    const unsigned long long z[2] = {0ull, 0ull};
    MEPP_CUDA_CHECK(cudaMemcpyToSymbol(g_mma_steps, z, sizeof(z)));
  }
  return 0;
}

int mepp_profile_gemm_check(const void* x_planes_dev, const void* y_planes_dev, double* S_dev, int64_t m_rows,
                            int C, int64_t n_pad, const int32_t* rows_dev, int n_rows_sel, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(x_planes_dev && y_planes_dev && S_dev && n_rows_sel > 0, "profile_gemm_check: bad arguments");
  MEPP_REQUIRE(mepp_device_ok(), "profile_gemm_check: no sm_100 CUDA device (no CPU fallback)");
  int nt = 0;
  int BN = mepp_choose_bn(C, &nt);
  dim3 grid((unsigned)((C + 127) / 128), (unsigned)n_rows_sel);
  gemm_check_kernel<<<grid, 128, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint8_t*>(x_planes_dev), reinterpret_cast<const uint8_t*>(y_planes_dev), S_dev,
      m_rows, C, n_pad / kBlockK, BN, rows_dev, n_rows_sel);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"

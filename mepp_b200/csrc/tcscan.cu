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
// CUDA-core pre-filter; the int32 accumulators stay in TMEM and only their SIGN is read back.  (N-tiles are 128
// columns = 16 units wide: four accumulator stages fit the 512 TMEM columns.)
//
// Conservative integer weights (tc_weights_kernel): per filter, penalties pen_j(c) = max_c' W[c'][j] - W[c][j] >= 0
// are scaled by s = 127 / D, D = sum_j max_j - threshold + rounding slack, and floored: q = -min(128, floor(s pen)).
// A window can only pass the float32 threshold when its penalty is below D, i.e. sum_j q_j >= -126; 127 is folded
// into tap 0, a base contributes 4 q (one-hot value 4) and an N base sum_c q(c) (one-hot 1,1,1,1 -- the reference's N
// scores the column mean), so "accumulator >= 0" is a test with NO false negatives, the same for every column.
// Survivors (~3e-4 of the windows) are appended to per-warp candidate lists; scan.cu evaluates them in the
// canonical float32 order, so every reported score still comes from the exact path.
//
// Structure: persistent CTAs (one per SM); a producer thread streams 2 KB tiles of F through a 16-stage cp.async.bulk
// ring (the weights of the launch are loaded once), issuer threads feed the tensor pipe, epilogue warps read the signs
// back (tcgen05.ld.pack::16b, AND trees, candidate lists).  Three kernels, newest last -- all give identical candidates:
//   tc_scan_kernel    two accumulator stages of 256 columns, two issuers                      (MEPP_TC_STAGES=2)
//   tc_scan4_kernel   four stages of 128 columns, two issuers, four epilogue groups           (MEPP_TC_ISSUERS=2)
//   tc_scan4i_kernel  four stages, stage = issuer = epilogue group, uniform-datapath issue    (the default)
#include "tcscan.cuh"
#include "tcgen05.cuh"
#include <stdlib.h>

namespace mepp {

constexpr int kTcIssuers = 2;
constexpr int kTcThreads = 32 * (1 + kTcIssuers + kTcEpiWarps);   // producer, MMA issuers, epilogue warps
constexpr int kTcStages = 16;     // stream tiles in flight: a 2 KB bulk copy takes ~1.5 us, a tile ~0.25 us of work
constexpr int kTcStageBytes = 2304;   // 2048 + 32 * ks_max, ks_max <= 5 (motifs up to 32 taps + 3 phases)
constexpr int kTcAccCols = 256;    // TMEM columns of one accumulator stage = widest N-tile (N = 256 keeps the MMA
                                   // compute bound: at N = 128 the operand reads from shared memory halve its rate)
constexpr int kTcAccStages = 2;    // one stage per MMA issuer: 2 x 256 columns = the whole TMEM

__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// 32 accumulator columns as 16 registers: .pack::16b keeps the low halves of two adjacent columns in one register (the
// accumulators of this kernel fit 16 bits: |4 (127 + sum q)| <= 4 * 128 * 32), so the sign test needs half the
// registers and half the logic ops.
__device__ __forceinline__ void tmem_ld16p_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}

template <bool PACK>
__global__ void __launch_bounds__(kTcThreads, 1) tc_scan_kernel(const __grid_constant__ TcLaunch ln,
                                                                const __grid_constant__ TcScanArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b_pad = (ln.b_bytes + 1023u) & ~1023u;
  uint8_t* sB = smem;
  uint8_t* sA = smem + b_pad;
  uint64_t* afull = reinterpret_cast<uint64_t*>(sA + kTcStages * kTcStageBytes);   // [kTcStages]
  uint64_t* aempty = afull + kTcStages;                                            // [kTcStages]
  uint64_t* tfull = aempty + kTcStages;                                            // [4 groups][kTcAccStages]
  uint64_t* tempty = tfull + 4 * kTcAccStages;                                     // [kTcAccStages]
  uint64_t* bfull = tempty + kTcAccStages;                                         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull + 1);
  // per N-tile: {B descriptor low word of K-step 0, instruction descriptor, K-steps, B K-step stride >> 4}
  __shared__ uint4 s_nt[kTcMaxTiles];

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(smem_u32(&afull[s]), 1); mbar_init(smem_u32(&aempty[s]), kTcIssuers); }
    // an accumulator stage (256 columns) is drained by two groups of four epilogue warps, 128 columns each
    // "accumulator ready" is signalled per (epilogue group, stage): a parity wait only tells adjacent phases apart, so
    // every waiter has to see every phase of its barrier -- a group skips the N-tiles of a stage that other groups
    // drain, and two of a group's items on different stages are not ordered (tools/sim_tc_pipeline.py checks the protocol)
    for (int s = 0; s < 4 * kTcAccStages; ++s) mbar_init(smem_u32(&tfull[s]), 1);
    for (int s = 0; s < kTcAccStages; ++s) mbar_init(smem_u32(&tempty[s]), 8);
    mbar_init(smem_u32(bfull), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(kTcAccStages * kTcAccCols)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 2 && lane < ln.n_tiles) {
    // K-major, no-swizzle descriptors: low word = start >> 4 | LBO >> 4 << 16, high word = SBO >> 4 | version 1 << 14.
    //   B: [kstep][16-byte K chunk][column][16 bytes]: LBO = 16 bn, SBO 128
    const TcTile tl = ln.tiles[lane];
    const uint32_t bn = tl.bn;
    // instruction descriptor: D = s32 (2 at [4,6)), A = u8 (0 at [7,10)), B = s8 (1 at [10,13)), K-major both,
    // N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (2u << 4) | (1u << 10) | ((bn >> 3) << 17) | ((uint32_t)(kTcRowsPerTile >> 4) << 24);
    s_nt[lane] = make_uint4((((smem_u32(sB) + tl.b_off) & 0x3FFFFu) >> 4) | (bn << 16), idesc, tl.ksteps, 2u * bn);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t a_bytes = 2048u + 32u * (uint32_t)ln.ks_max;
  const int n_tiles = ln.n_tiles;
  const uint32_t park_ns = (uint32_t)g.mode >> 8;      // suspend-time hint of the barrier waits (0: plain try_wait loop)

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
        mbar_wait_parked(smem_u32(&aempty[stage]), phase ^ 1u, park_ns);
        const uint32_t fb = smem_u32(&afull[stage]);
        if (g.mode & 8) { mbar_arrive(fb); }        // (timing experiment: no stream copies)
        else {
          mbar_expect_tx(fb, a_bytes);
          bulk_g2s(smem_u32(sA + stage * kTcStageBytes), g.onehot + tile * 2048, a_bytes, fb);
        }
        if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp <= kTcIssuers) {
    // ------------------------------------------------------------ MMA issuers (warps 1, 2)
    // Issuer i owns accumulator stage i and the N-tiles with (running N-tile number) % 2 == i.  tcgen05.mma holds
    // its issuing thread until the tensor pipe takes the instruction, and a lone warp needs ~300 cycles for the
    // barrier waits and bookkeeping around an N-tile (measured with the cycle stamps of tools/gpu_tc_clk.py): with two
    // issuers the pipe works on one N-tile while the other issuer does its bookkeeping.  The whole warp walks the loop
    // (uniform datapath), one lane issues; everything an N-tile needs is one 16-byte table entry.
    //   A: row r = bytes [16 r + 32 ks, +32) of the stream tile (overlapping rows): LBO 16, SBO 128
    const uint32_t me = (uint32_t)(warp - 1);
    const uint32_t kDescHi = 8u | (1u << 14);
    mbar_wait_parked(smem_u32(bfull), 0u, park_ns);
    tc_fence_after();
    const bool issuer = lane == 0;
    const uint32_t a_lo_base = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t tfull_me = smem_u32(&tfull[me]), tempty_me = smem_u32(&tempty[me]);   // tfull[group * 2 + stage]
    const uint32_t d_tmem = tmem_base + me * kTcAccCols;
    uint32_t stage = 0, phase = 0, nseq = 0, iseq = 0, acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < g.n_mtiles; tile += gridDim.x) {
      mbar_wait_parked(smem_u32(&afull[stage]), phase, park_ns);
      tc_fence_after();
      const uint32_t a_lo = a_lo_base + stage * (uint32_t)(kTcStageBytes / 16);
      for (int nt = 0; nt < n_tiles; ++nt, ++nseq) {
        const uint4 c = s_nt[nt];
        const uint32_t n_halves = c.w > 256u ? 2u : 1u;        // (c.w = 2 bn: more than four 32-column chunks)
        const uint32_t item0 = iseq;
        iseq += n_halves;
        if ((nseq & 1u) != me) continue;
        mbar_wait(tempty_me, acc_phase ^ 1u);        // (two spinning warps cost little; a parked wait wakes ~60 cycles later)
        acc_phase ^= 1u;
        tc_fence_after();
        if (g.dbg_clk && blockIdx.x == 0 && issuer && nseq < 256) g.dbg_clk[nseq] = clock64();
        if (issuer) {
          if (!(g.mode & 2)) {
#pragma unroll
            for (uint32_t ks = 0; ks < 5; ++ks)
              if (ks < c.z)
                tc_mma_i8(d_tmem, ((uint64_t)kDescHi << 32) | (a_lo + 2u * ks), ((uint64_t)kDescHi << 32) | (c.x + ks * c.w),
                          c.y, ks > 0 ? 1u : 0u);
          }
          tc_commit(tfull_me + (item0 & 3u) * (8u * kTcAccStages));              // the group(s) that drain this N-tile
          if (n_halves == 2u) tc_commit(tfull_me + ((item0 + 1u) & 3u) * (8u * kTcAccStages));
          if (g.dbg_clk && blockIdx.x == 0 && nseq < 256) g.dbg_clk[256 + nseq] = clock64();
        }
        __syncwarp();
      }
      if (issuer) tc_commit(smem_u32(&aempty[stage]));     // the stream tile is free once both issuers' MMAs retired
      __syncwarp();
      if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 3..18)
    // Four groups of four warps (one warp per TMEM lane quarter).  Work items are the non-empty 128-column halves of
    // the N-tiles in running order; item i belongs to group i % 4, so the groups never wait for each other and share
    // the work evenly whatever the N-tile widths.  A warp's instruction stream is one dependent chain (wait, load,
    // AND-reduce, vote, branch): it takes four warps per scheduler to keep the issue slots busy.
    const unsigned full = 0xffffffffu;
    const int ew = warp - (1 + kTcIssuers), q = warp & 3, gidx = ew >> 2;
    uint32_t my_count = 0, nseq = 0, iseq = 0, seen0 = 0, seen1 = 0;   // seen: this group's items per stage so far
    const int64_t list = (int64_t)blockIdx.x * kTcEpiWarps + ew;
    uint2* const my_list = g.cand + list * (int64_t)g.list_cap;
    for (int64_t tile = blockIdx.x; tile < g.n_mtiles; tile += gridDim.x) {
      const uint32_t row = (uint32_t)(tile * kTcRowsPerTile) + (uint32_t)(q * 32 + lane);
      int dbg_col = 0;
      for (int nt = 0; nt < n_tiles; ++nt, ++nseq) {
        const uint4 cst = s_nt[nt];
        const int bn = (int)(cst.w >> 1), n_chunks = bn >> 5;
        const int n_halves = n_chunks > 4 ? 2 : 1;
        for (int hf = 0; hf < n_halves; ++hf, ++iseq) {
          if ((int)(iseq & 3u) != gidx) continue;
          const uint32_t acc = nseq & 1u;
          const uint32_t acc_phase = (acc ? seen1 : seen0) & 1u;
          if (acc) ++seen1; else ++seen0;
          mbar_wait_parked(smem_u32(&tfull[gidx * kTcAccStages + acc]), acc_phase, park_ns);
          tc_fence_after();
          const bool stamp = g.dbg_clk && blockIdx.x == 0 && lane == 0 && (ew & 3) == 0 && hf == 0 && nseq < 256;
          if (stamp) g.dbg_clk[512 + nseq] = clock64();
          const int c0 = 4 * hf, nc = n_chunks - c0 < 4 ? n_chunks - c0 : 4;     // this item's chunks: c0 .. c0 + nc - 1
          // the item's columns are read with one wait, and the accumulator stage is handed back to its MMA issuer
          // before the values are looked at
          uint32_t v[4][PACK ? 16 : 32];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < nc && !(g.mode & 1)) {
              const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kTcAccCols + (uint32_t)((c0 + k) * 32);
              if constexpr (PACK) tmem_ld16p_nowait(ta, v[k]);
              else tmem_ld32_nowait(ta, v[k]);
            }
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(smem_u32(&tempty[acc]));
            if (n_halves == 1) mbar_arrive(smem_u32(&tempty[acc]));     // (nobody drains the empty other half)
          }
          if (stamp) g.dbg_clk[768 + nseq] = clock64();
          if (!(g.mode & 5)) {
            if (g.dbg_acc != nullptr && tile == 0) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k >= nc) break;
                int32_t* d = g.dbg_acc + (int64_t)(q * 32 + lane) * g.dbg_ld + dbg_col + (c0 + k) * 32;
                if constexpr (PACK) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) { d[2 * j] = (int32_t)(int16_t)(v[k][j] & 0xffffu); d[2 * j + 1] = (int32_t)(int16_t)(v[k][j] >> 16); }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j) d[j] = (int32_t)v[k][j];
                }
              }
            }
            // sign tests of all chunks first (independent AND trees), one vote for the whole item
            constexpr int NV = PACK ? 16 : 32;
            uint32_t hasm = 0u;                      // bit k: some accumulator of this row in chunk k is non-negative
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k >= nc) break;
              uint32_t t = v[k][0];
#pragma unroll
              for (int j = 1; j < NV; ++j) t &= v[k][j];
              const bool has = PACK ? (t & 0x80008000u) != 0x80008000u : (int32_t)t >= 0;
              hasm |= has ? 1u << k : 0u;
            }
            if (__any_sync(full, hasm != 0u)) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k >= nc) break;
                unsigned bal = __ballot_sync(full, (hasm >> k) & 1u);      // lanes (rows) with a candidate in chunk k
                if (bal == 0u) continue;
                uint32_t cm = 0u;                    // bit p: (unit p / 4 of the chunk, phase p % 4) passes on a strand
                if ((hasm >> k) & 1u) {
#pragma unroll
                  for (int p = 0; p < 16; ++p) {
                    bool pass;
                    if constexpr (PACK) pass = (v[k][p] & 0x80008000u) != 0x80008000u;
                    else pass = (int32_t)(v[k][2 * p] & v[k][2 * p + 1]) >= 0;
                    if (pass) cm |= 1u << p;
                  }
                }
                const uint32_t ubase = ((uint32_t)ln.tiles[nt].unit0 + (uint32_t)(c0 + k) * 4u) << 2;
                // the few lanes with candidates take their turn: slots my_count .. in lane order
                while (bal != 0u) {
                  const int src = __ffs((int)bal) - 1;
                  bal &= bal - 1u;
                  const int n = __popc(__shfl_sync(full, cm, src));
                  if (lane == src) {
                    uint32_t slot = my_count;
                    while (cm != 0u) {
                      const int p = __ffs((int)cm) - 1;
                      cm &= cm - 1u;
                      if (slot < g.list_cap) my_list[slot] = make_uint2(row, ubase + (uint32_t)p);
                      ++slot;
                    }
                  }
                  my_count += (uint32_t)n;
                }
              }
            }
          }
          if (stamp) g.dbg_clk[1024 + nseq] = clock64();
        }
        dbg_col += bn;
      }
    }
    if (lane == 0) g.cand_count[list] = my_count;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(kTcAccStages * kTcAccCols)) : "memory");
  }
}

// ---- four-stage form ------------------------------------------------------------------------------------------------
// The two-stage kernel above is bound by its hand-overs, not by the tensor pipe (17 % active, ncu): an accumulator stage
// goes issuer -> commit -> epilogue wake-up -> TMEM reads -> arrive -> issuer wake-up (~700-900 cycles, cycle stamps of
// tools/gpu_tc_clk.py) for ~350 cycles of products, and with two stages only one other N-tile can cover that.  Here
// an N-tile is at most 128 columns (16 scan units) and TMEM holds FOUR accumulator stages of 128 columns; N-tile number
// s (running over the whole launch) uses stage s % 4, is issued by issuer s % 2 and drained by epilogue group s % 4 --
// every group owns one stage, so no barrier is shared between groups and three other N-tiles are in flight behind
// every hand-over.  (At N = 128 an MMA reads as many operand bytes per cycle as the shared memory delivers, 8 KB in
// 64 cycles: the narrower tile costs no tensor time.)
// What is left (cycle stamps, profiles/r02_tc_scan_stamps.log): one thread gets a tcgen05.mma of this size accepted only
// every ~100 cycles (64 cycles of products), and an issuer spends ~400 more cycles per N-tile between its batches
// (try_wait ~90 even when the phase is complete, fences, commit), so two issuers keep the pipe ~35 % busy.  Tried and
// dropped, all measured on cfg 3 (6.1 ms per step for this kernel): one issuer per stage (21 warps leave 80 registers
// per thread, the epilogue spills: 7.1 ms); the first warp of every epilogue group issuing the group's next N-tile before
// its sign tests (no issuer warps; the four groups fall into lock step and the issuing warp's chain load -> issue ->
// test is the critical path: 6.9 ms); three issuers over four stages (a stage would change issuers, and a parity wait
// cannot tell phase k + 1 of its barrier from phase k - 1: races).
constexpr int kTc4Stages = 4;
constexpr int kTc4Cols = 128;

template <bool PACK>
__global__ void __launch_bounds__(kTcThreads, 1) tc_scan4_kernel(const __grid_constant__ TcLaunch ln,
                                                                 const __grid_constant__ TcScanArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b_pad = (ln.b_bytes + 1023u) & ~1023u;
  uint8_t* sB = smem;
  uint8_t* sA = smem + b_pad;
  uint64_t* afull = reinterpret_cast<uint64_t*>(sA + kTcStages * kTcStageBytes);   // [kTcStages]
  uint64_t* aempty = afull + kTcStages;                                            // [kTcStages]
  uint64_t* tfull = aempty + kTcStages;                                            // [kTc4Stages]
  uint64_t* tempty = tfull + kTc4Stages;                                           // [kTc4Stages]
  uint64_t* bfull = tempty + kTc4Stages;                                           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull + 1);
  __shared__ uint4 s_nt[kTcMaxTiles];

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(smem_u32(&afull[s]), 1); mbar_init(smem_u32(&aempty[s]), kTcIssuers); }
    for (int s = 0; s < kTc4Stages; ++s) { mbar_init(smem_u32(&tfull[s]), 1); mbar_init(smem_u32(&tempty[s]), 4); }
    mbar_init(smem_u32(bfull), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(kTc4Stages * kTc4Cols)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 1 + kTcIssuers && lane < ln.n_tiles) {
    const TcTile tl = ln.tiles[lane];
    const uint32_t bn = tl.bn;
    const uint32_t idesc = (2u << 4) | (1u << 10) | ((bn >> 3) << 17) | ((uint32_t)(kTcRowsPerTile >> 4) << 24);
    s_nt[lane] = make_uint4((((smem_u32(sB) + tl.b_off) & 0x3FFFFu) >> 4) | (bn << 16), idesc, tl.ksteps, 2u * bn);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t a_bytes = 2048u + 32u * (uint32_t)ln.ks_max;
  const int n_tiles = ln.n_tiles;
  const uint32_t park_ns = (uint32_t)g.mode >> 8;

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
        mbar_wait_parked(smem_u32(&aempty[stage]), phase ^ 1u, park_ns);
        const uint32_t fb = smem_u32(&afull[stage]);
        mbar_expect_tx(fb, a_bytes);
        bulk_g2s(smem_u32(sA + stage * kTcStageBytes), g.onehot + tile * 2048, a_bytes, fb);
        if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp <= kTcIssuers) {
    // ------------------------------------------------------------ MMA issuers (warps 1, 2): N-tiles s % 2 == me, i.e.
    // issuer me owns the stages me and me + 2 (a stage must stay with ONE issuer: a parity wait only tells adjacent
    // phases of a barrier apart, so the uses of a stage have to be issued in order by one thread)
    const uint32_t me = (uint32_t)(warp - 1);
    const uint32_t kDescHi = 8u | (1u << 14);
    mbar_wait_parked(smem_u32(bfull), 0u, park_ns);
    tc_fence_after();
    const bool issuer = lane == 0;
    const uint32_t a_lo_base = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
    uint32_t stage = 0, phase = 0, seq = 0;
    for (int64_t tile = blockIdx.x; tile < g.n_mtiles; tile += gridDim.x) {
      if (g.dbg_clk && blockIdx.x == 0 && issuer && seq < 256) g.dbg_clk[9 * 256 + seq] = clock64();   // loop top
      mbar_wait_parked(smem_u32(&afull[stage]), phase, park_ns);
      tc_fence_after();
      if (g.dbg_clk && blockIdx.x == 0 && issuer && seq < 256) g.dbg_clk[5 * 256 + seq] = clock64();   // stream tile there
      const uint32_t a_lo = a_lo_base + stage * (uint32_t)(kTcStageBytes / 16);
      for (int nt = 0; nt < n_tiles; ++nt, ++seq) {
        if ((seq & 1u) != me) continue;
        const uint32_t st = seq & 3u;                        // N-tile seq is use number seq >> 2 of stage seq & 3
        const uint4 c = s_nt[nt];
        mbar_wait(smem_u32(&tempty[st]), ((seq >> 2) & 1u) ^ 1u);
        tc_fence_after();
        if (g.dbg_clk && blockIdx.x == 0 && issuer && seq < 256) g.dbg_clk[seq] = clock64();
        if (issuer) {
          const uint32_t d_tmem = tmem_base + st * kTc4Cols;
#pragma unroll
          for (uint32_t ks = 0; ks < 5; ++ks)
            if (ks < c.z) {
              tc_mma_i8(d_tmem, ((uint64_t)kDescHi << 32) | (a_lo + 2u * ks), ((uint64_t)kDescHi << 32) | (c.x + ks * c.w),
                        c.y, ks > 0 ? 1u : 0u);
              if (ks == 0 && g.dbg_clk && blockIdx.x == 0 && seq < 256) g.dbg_clk[6 * 256 + seq] = clock64();   // first MMA issued
            }
          if (g.dbg_clk && blockIdx.x == 0 && seq < 256) g.dbg_clk[7 * 256 + seq] = clock64();     // last MMA issued
          tc_commit(smem_u32(&tfull[st]));
          if (g.dbg_clk && blockIdx.x == 0 && seq < 256) g.dbg_clk[256 + seq] = clock64();
        }
        __syncwarp();
      }
      if (issuer) tc_commit(smem_u32(&aempty[stage]));     // the stream tile is free once both issuers' MMAs retired
      __syncwarp();
      if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 3..18): group = stage
    const unsigned full = 0xffffffffu;
    const int ew = warp - (1 + kTcIssuers), q = warp & 3, gidx = ew >> 2;
    uint32_t my_count = 0, seq = 0, seen = 0;
    const int64_t list = (int64_t)blockIdx.x * kTcEpiWarps + ew;
    uint2* const my_list = g.cand + list * (int64_t)g.list_cap;
    const uint32_t tfull_me = smem_u32(&tfull[gidx]), tempty_me = smem_u32(&tempty[gidx]);
    const uint32_t t_me = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)gidx * kTc4Cols;
    for (int64_t tile = blockIdx.x; tile < g.n_mtiles; tile += gridDim.x) {
      const uint32_t row = (uint32_t)(tile * kTcRowsPerTile) + (uint32_t)(q * 32 + lane);
      int dbg_col = 0;
      for (int nt = 0; nt < n_tiles; ++nt, ++seq) {
        const int bn = (int)(s_nt[nt].w >> 1);
        if ((int)(seq & 3u) != gidx) { dbg_col += bn; continue; }
        const int nc = bn >> 5;                                   // 32-column chunks of this N-tile (<= 4)
        mbar_wait_parked(tfull_me, seen & 1u, park_ns);
        ++seen;
        tc_fence_after();
        const bool stamp = g.dbg_clk && blockIdx.x == 0 && lane == 0 && q == 0 && seq < 256;
        if (stamp) g.dbg_clk[512 + seq] = clock64();
        uint32_t v[4][PACK ? 16 : 32];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < nc && !(g.mode & 1)) {
            if constexpr (PACK) tmem_ld16p_nowait(t_me + (uint32_t)(k * 32), v[k]);
            else tmem_ld32_nowait(t_me + (uint32_t)(k * 32), v[k]);
          }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_me);
        if (stamp) g.dbg_clk[768 + seq] = clock64();
        if (!(g.mode & 5)) {
          if (g.dbg_acc != nullptr && tile == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k >= nc) break;
              int32_t* d = g.dbg_acc + (int64_t)(q * 32 + lane) * g.dbg_ld + dbg_col + k * 32;
              if constexpr (PACK) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { d[2 * j] = (int32_t)(int16_t)(v[k][j] & 0xffffu); d[2 * j + 1] = (int32_t)(int16_t)(v[k][j] >> 16); }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) d[j] = (int32_t)v[k][j];
              }
            }
          }
          constexpr int NV = PACK ? 16 : 32;
          uint32_t hasm = 0u;                      // bit k: some accumulator of this row in chunk k is non-negative
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k >= nc) break;
            uint32_t t = v[k][0];
#pragma unroll
            for (int j = 1; j < NV; ++j) t &= v[k][j];
            const bool has = PACK ? (t & 0x80008000u) != 0x80008000u : (int32_t)t >= 0;
            hasm |= has ? 1u << k : 0u;
          }
          if (__any_sync(full, hasm != 0u)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k >= nc) break;
              unsigned bal = __ballot_sync(full, (hasm >> k) & 1u);      // lanes (rows) with a candidate in chunk k
              if (bal == 0u) continue;
              uint32_t cm = 0u;                    // bit p: (unit p / 4 of the chunk, phase p % 4) passes on a strand
              if ((hasm >> k) & 1u) {
#pragma unroll
                for (int p = 0; p < 16; ++p) {
                  bool pass;
                  if constexpr (PACK) pass = (v[k][p] & 0x80008000u) != 0x80008000u;
                  else pass = (int32_t)(v[k][2 * p] & v[k][2 * p + 1]) >= 0;
                  if (pass) cm |= 1u << p;
                }
              }
              const uint32_t ubase = ((uint32_t)ln.tiles[nt].unit0 + (uint32_t)k * 4u) << 2;
              while (bal != 0u) {
                const int src = __ffs((int)bal) - 1;
                bal &= bal - 1u;
                const int n = __popc(__shfl_sync(full, cm, src));
                if (lane == src) {
                  uint32_t slot = my_count;
                  while (cm != 0u) {
                    const int p = __ffs((int)cm) - 1;
                    cm &= cm - 1u;
                    if (slot < g.list_cap) my_list[slot] = make_uint2(row, ubase + (uint32_t)p);
                    ++slot;
                  }
                }
                my_count += (uint32_t)n;
              }
            }
          }
        }
        if (stamp) g.dbg_clk[1024 + seq] = clock64();
        dbg_col += bn;
      }
    }
    if (lane == 0) g.cand_count[list] = my_count;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(kTc4Stages * kTc4Cols)) : "memory");
  }
}

// ---- four-stage form with one issuer per stage ------------------------------------------------------------------------
// The issuing thread of tc_scan4_kernel is a chain of long-latency instructions (try_wait ~90 cycles, every tcgen05.mma
// ~100, commit ~60, ...: ~1200 cycles per N-tile of its own, profiles/r02_tc_scan_stamps.log), so two issuers keep the
// tensor pipe 20 % busy.  Here stage = issuer = epilogue group: N-tile number s of the CTA's running sequence belongs to
// issuer s % 4, which owns accumulator stage s % 4 and hands it to epilogue group s % 4 -- four independent
// (issuer, stage, group) pipelines that share only the stream ring and the tensor pipe, each with one private pair of
// barriers in strict alternation.  An issuer only waits for / releases the stream tiles it has an N-tile of.
// The first attempt at this (21 warps under one register cap of 80) spilled in the epilogue; here the warps sit on
// warpgroup boundaries and trade registers with setmaxnreg: warps 0-3 issuers, warp 4 producer (5-7 idle), both
// warpgroups at 40 registers, warps 8-23 epilogue at 96.  No debug stamps: the production loop carries no extra loads.
constexpr int kTc4iThreads = 32 * (8 + kTcEpiWarps);
constexpr int kTcDefaultIssuers = 4;      // one issuer per accumulator stage (MEPP_TC_ISSUERS=2: tc_scan4_kernel)
constexpr int kTcDefaultSplit = 1;
int tc_issuers();
int tc_split();

// SPLIT = 2: every (issuer, epilogue group) pipeline owns TWO half-stages of 64 columns and alternates between them
// (N-tiles of <= 64 columns = 8 scan units), so the issuer fills one half while the group drains the other: the time per
// item is the larger of the two sides instead of their sum.
template <bool PACK, int MODE, int SPLIT>      // MODE (timing experiments, MEPP_TC_MODE): 1 no TMEM reads / tests, 2 no products
__global__ void __launch_bounds__(kTc4iThreads, 1) tc_scan4i_kernel(const __grid_constant__ TcLaunch ln,
                                                                   const __grid_constant__ TcScanArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b_pad = (ln.b_bytes + 1023u) & ~1023u;
  uint8_t* sB = smem;
  uint8_t* sA = smem + b_pad;
  uint64_t* afull = reinterpret_cast<uint64_t*>(sA + kTcStages * kTcStageBytes);   // [kTcStages]
  uint64_t* aempty = afull + kTcStages;                                            // [kTcStages]
  constexpr int kAcc = kTc4Stages * SPLIT;          // accumulator (half-)stages
  constexpr uint32_t kCols = (uint32_t)(kTc4Cols / SPLIT);
  uint64_t* tfull = aempty + kTcStages;                                            // [kAcc]
  uint64_t* tempty = tfull + kAcc;                                                 // [kAcc]
  uint64_t* bfull = tempty + kAcc;                                                 // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull + 1);
  __shared__ uint4 s_nt[kTcMaxTiles];
  __shared__ uint32_t s_cnt[kTcEpiWarps];      // candidates of every epilogue warp's list so far
  const int n_tiles = ln.n_tiles;
  // issuers with an N-tile of any one stream tile: consecutive N-tile numbers go to different issuers
  const uint32_t n_release = n_tiles < kTc4Stages ? (uint32_t)n_tiles : (uint32_t)kTc4Stages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(smem_u32(&afull[s]), 1); mbar_init(smem_u32(&aempty[s]), n_release); }
    for (int s = 0; s < kAcc; ++s) { mbar_init(smem_u32(&tfull[s]), 1); mbar_init(smem_u32(&tempty[s]), 4); }
    mbar_init(smem_u32(bfull), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(kTc4Stages * kTc4Cols)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 6 && lane < kTcEpiWarps) s_cnt[lane] = 0u;
  if (warp == 5 && lane < n_tiles) {
    const TcTile tl = ln.tiles[lane];
    const uint32_t bn = tl.bn;
    const uint32_t idesc = (2u << 4) | (1u << 10) | ((bn >> 3) << 17) | ((uint32_t)(kTcRowsPerTile >> 4) << 24);
    s_nt[lane] = make_uint4((((smem_u32(sB) + tl.b_off) & 0x3FFFFu) >> 4) | (bn << 16), idesc,
                            (uint32_t)tl.ksteps | ((uint32_t)tl.unit0 << 16), 2u * bn);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t a_bytes = 2048u + 32u * (uint32_t)ln.ks_max;
  const uint32_t park_ns = (uint32_t)g.mode >> 8;
  // stream tiles of this CTA: blockIdx.x, + gridDim.x, ...; its N-tiles are numbered mt * n_tiles + nt
  const uint32_t my_mtiles = (uint32_t)((g.n_mtiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    if (warp == 4) {
      // ---------------------------------------------------------- producer
      if (lane == 0) {
        const uint32_t bb = smem_u32(bfull);
        mbar_expect_tx(bb, ln.b_bytes);
        for (uint32_t off = 0; off < ln.b_bytes; off += 32768u) {
          const uint32_t n = ln.b_bytes - off < 32768u ? ln.b_bytes - off : 32768u;
          bulk_g2s(smem_u32(sB + off), g.B + off, n, bb);
        }
        uint32_t stage = 0, phase = 0;
        for (int64_t tile = blockIdx.x; tile < g.n_mtiles; tile += gridDim.x) {
          mbar_wait_parked(smem_u32(&aempty[stage]), phase ^ 1u, park_ns);
          const uint32_t fb = smem_u32(&afull[stage]);
          mbar_expect_tx(fb, a_bytes);
          bulk_g2s(smem_u32(sA + stage * kTcStageBytes), g.onehot + tile * 2048, a_bytes, fb);
          if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp < kTc4Stages) {
      // ---------------------------------------------------------- MMA issuer of stage `me` (N-tiles me, me + 4, ...)
      // The issuing thread's time is instruction latency (ncu source page: ~190 dependent instructions per N-tile at ~7
      // cycles each), so everything it touches is kept warp-uniform for the compiler -- the warp index comes from a
      // broadcast, the N-tile's constants from the kernel parameters (constant bank) instead of shared memory, the
      // K-step loop runs `ksteps` times instead of five predicated bodies -- and the descriptors advance by one add.
      const uint32_t me = __shfl_sync(0xffffffffu, (uint32_t)warp, 0);
      const uint32_t kDescHi = 8u | (1u << 14);
      mbar_wait_parked(smem_u32(bfull), 0u, park_ns);
      tc_fence_after();
      const bool issuer = elect_one();
      const uint32_t a_lo_base = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_lo_base = smem_u32(sB);
      const uint32_t tfull_me = smem_u32(&tfull[me * SPLIT]), tempty_me = smem_u32(&tempty[me * SPLIT]);
      const uint32_t d_tmem_me = tmem_base + me * kTc4Cols;
      const uint32_t nt_n = (uint32_t)n_tiles;
      uint32_t mt = 0, nt = me, use = 0, mt_ready = 0xffffffffu;
      while (nt >= nt_n) { nt -= nt_n; ++mt; }
      while (mt < my_mtiles) {
        const uint32_t stage = mt % (uint32_t)kTcStages;
        if (mt != mt_ready) {
          mbar_wait_parked(smem_u32(&afull[stage]), (mt / (uint32_t)kTcStages) & 1u, park_ns);
          mt_ready = mt;
        }
        const TcTile tl = ln.tiles[nt];
        const uint32_t bn = tl.bn, ksteps = tl.ksteps;
        const uint32_t idesc = (2u << 4) | (1u << 10) | ((bn >> 3) << 17) | ((uint32_t)(kTcRowsPerTile >> 4) << 24);
        uint32_t b_lo = (((b_lo_base + tl.b_off) & 0x3FFFFu) >> 4) | (bn << 16);
        uint32_t a_lo = a_lo_base + stage * (uint32_t)(kTcStageBytes / 16);
        // item `use` of this pipeline: half-stage use % SPLIT, its use number use / SPLIT
        const uint32_t hs = SPLIT == 2 ? (use & 1u) : 0u;
        mbar_wait(tempty_me + 8u * hs, ((use / (uint32_t)SPLIT) & 1u) ^ 1u);
        const uint32_t d_tmem = d_tmem_me + hs * kCols;
        ++use;
        tc_fence_after();
        // this issuer's next N-tile: in another stream tile -> this one is released behind the products
        uint32_t mt2 = mt, nt2 = nt + (uint32_t)kTc4Stages;
        while (nt2 >= nt_n) { nt2 -= nt_n; ++mt2; }
        if (issuer) {
          if (!(MODE & 2)) {
            tc_mma_i8(d_tmem, ((uint64_t)kDescHi << 32) | a_lo, ((uint64_t)kDescHi << 32) | b_lo, idesc, 0u);
            for (uint32_t ks = 1; ks < ksteps; ++ks) {
              a_lo += 2u; b_lo += 2u * bn;
              tc_mma_i8(d_tmem, ((uint64_t)kDescHi << 32) | a_lo, ((uint64_t)kDescHi << 32) | b_lo, idesc, 1u);
            }
          }
          tc_commit(tfull_me + 8u * hs);
          if (mt2 != mt) tc_commit(smem_u32(&aempty[stage]));
        }
        __syncwarp();
        mt = mt2; nt = nt2;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;" ::: "memory");
    // ------------------------------------------------------------ epilogue (warps 8..23): group = stage
    // The epilogue warps are bound by the LENGTH of their instruction stream (ncu source page: ~210 instructions per item
    // at ~7 cycles each, 82 % of the time busy), and the compiler likes to rematerialise everything that derives from
    // the thread index (S2R + address arithmetic in front of every tcgen05.ld / arrive): the loop invariants are pinned
    // in registers (an empty asm the compiler cannot see through), the row number advances by addition.
    const int ew = warp - 8, q = warp & 3, gidx = ew >> 2;
    uint32_t use = 0;
    const int64_t list = (int64_t)blockIdx.x * kTcEpiWarps + ew;
    uint2* my_list = g.cand + list * (int64_t)g.list_cap;
    uint32_t tfull_me = smem_u32(&tfull[gidx * SPLIT]), tempty_me = smem_u32(&tempty[gidx * SPLIT]);
    uint32_t t_grp = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)gidx * kTc4Cols;
    uint32_t cnt_me = smem_u32(&s_cnt[ew]);
    uint32_t lane0 = lane == 0 ? 1u : 0u;
    const uint32_t nt_n = (uint32_t)n_tiles, row_step = gridDim.x * (uint32_t)kTcRowsPerTile, list_cap = g.list_cap;
    uint32_t row = blockIdx.x * (uint32_t)kTcRowsPerTile + (uint32_t)(q * 32 + lane);
    asm volatile("" : "+r"(tfull_me), "+r"(tempty_me), "+r"(t_grp), "+r"(cnt_me), "+r"(lane0), "+r"(row), "+l"(my_list));
    uint32_t mt = 0, nt = (uint32_t)gidx;
    while (nt >= nt_n) { nt -= nt_n; ++mt; row += row_step; }
    while (mt < my_mtiles) {
      const uint4 cst = s_nt[nt];
      const int nc = (int)(cst.w >> 6);                           // 32-column chunks of this N-tile (<= 4)
      const uint32_t hs = SPLIT == 2 ? (use & 1u) : 0u;
      mbar_wait_parked(tfull_me + 8u * hs, (use / (uint32_t)SPLIT) & 1u, park_ns);
      const uint32_t t_me = t_grp + hs * kCols;
      ++use;
      tc_fence_after();
      constexpr int kChunksMax = 4 / SPLIT;
      uint32_t v[kChunksMax][PACK ? 16 : 32];
      if (nc == kChunksMax && !(MODE & 9)) {          // full N-tile: the loads back to back
#pragma unroll
        for (int k = 0; k < kChunksMax; ++k) {
          if constexpr (PACK) tmem_ld16p_nowait(t_me + (uint32_t)(k * 32), v[k]);
          else tmem_ld32_nowait(t_me + (uint32_t)(k * 32), v[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < kChunksMax; ++k)
          if (k < nc && !(MODE & 1) && !((MODE & 8) && k >= 2)) {
            if constexpr (PACK) tmem_ld16p_nowait(t_me + (uint32_t)(k * 32), v[k]);
            else tmem_ld32_nowait(t_me + (uint32_t)(k * 32), v[k]);
          }
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane0) mbar_arrive(tempty_me + 8u * hs);
      constexpr int NV = PACK ? 16 : 32;
      uint32_t hasm = 0u;                      // bit k: some accumulator of this row in chunk k is non-negative
      if (!(MODE & 5)) {
        if (nc == kChunksMax) {                // full N-tile: straight-line code, the chunks' AND trees interleave
#pragma unroll
          for (int k = 0; k < kChunksMax; ++k) {
            uint32_t t = v[k][0];
#pragma unroll
            for (int j = 1; j < NV; ++j) t &= v[k][j];
            const bool has = PACK ? (t & 0x80008000u) != 0x80008000u : (int32_t)t >= 0;
            hasm |= has ? 1u << k : 0u;
          }
        } else {
#pragma unroll
          for (int k = 0; k < kChunksMax; ++k) {
            if (k >= nc || ((MODE & 8) && k >= 2)) break;
            uint32_t t = v[k][0];
#pragma unroll
            for (int j = 1; j < NV; ++j) t &= v[k][j];
            const bool has = PACK ? (t & 0x80008000u) != 0x80008000u : (int32_t)t >= 0;
            hasm |= has ? 1u << k : 0u;
          }
        }
      }
      // ~70 % of the (warp, N-tile) items hold a candidate somewhere, so this is the COMMON path: only the lanes (rows)
      // with a candidate enter it, there is no warp-wide vote / shuffle in it, and a lane reserves its slots of the
      // warp's list with one shared-memory atomic per chunk
      if (hasm != 0u) {
#pragma unroll
        for (int k = 0; k < kChunksMax; ++k) {
          if (k >= nc) break;
          if (!((hasm >> k) & 1u)) continue;
          uint32_t cm = 0u;                    // bit p: (unit p / 4 of the chunk, phase p % 4) passes on a strand
#pragma unroll
          for (int p = 0; p < 16; ++p) {
            bool pass;
            if constexpr (PACK) pass = (v[k][p] & 0x80008000u) != 0x80008000u;
            else pass = (int32_t)(v[k][2 * p] & v[k][2 * p + 1]) >= 0;
            if (pass) cm |= 1u << p;
          }
          const uint32_t ubase = ((cst.z >> 16) + (uint32_t)k * 4u) << 2;
          uint32_t slot = atoms_add(cnt_me, (uint32_t)__popc(cm));
          while (cm != 0u) {
            const int p = __ffs((int)cm) - 1;
            cm &= cm - 1u;
            if (slot < list_cap) my_list[slot] = make_uint2(row, ubase + (uint32_t)p);
            ++slot;
          }
        }
      }
      nt += (uint32_t)kTc4Stages;
      while (nt >= nt_n) { nt -= nt_n; ++mt; row += row_step; }
    }
    __syncwarp();
    if (lane == 0) g.cand_count[list] = s_cnt[ew];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(kTc4Stages * kTc4Cols)) : "memory");
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

// Scan units per N-tile: 16 (128 columns, four accumulator stages: tc_scan4_kernel, the default) or 32 (256 columns, two
// stages: tc_scan_kernel; MEPP_TC_STAGES=2).
int tc_issuers() {
  static const int v = [] { const char* e = getenv("MEPP_TC_ISSUERS"); return e ? atoi(e) : kTcDefaultIssuers; }();
  return v;
}
// 2: half-stages of 64 columns, two per (issuer, group) pipeline (tc_scan4i_kernel<.., 2>; MEPP_TC_SPLIT)
int tc_split() {
  static const int v = [] { const char* e = getenv("MEPP_TC_SPLIT"); return (e ? atoi(e) : kTcDefaultSplit) == 2 && tc_issuers() == 4 ? 2 : 1; }();
  return v;
}
int tc_units_per_tile() {
  static const int v = [] { const char* e = getenv("MEPP_TC_STAGES"); return (e && atoi(e) == 2) ? 32 : tc_split() == 2 ? 8 : 16; }();
  return v;
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
  dim3 grid((unsigned)ln.n_tiles, (unsigned)kTcUnitsPerTile);
  tc_weights_kernel<<<grid, 128, 0, st>>>(ln, lut_dev, bias_dev, mlen_dev, nfilt_dev, units_dev, unit_base, n_units_total,
                                          B_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tc_scan_launch(const TcLaunch& ln, const TcScanArgs& a, cudaStream_t st) {
  const size_t smem = ((size_t)ln.b_bytes + 1023) / 1024 * 1024 + (size_t)kTcStages * kTcStageBytes + 1024 /* barriers */;
  MEPP_REQUIRE(ln.b_bytes <= (uint32_t)kTcMaxBBytes && ln.ks_max >= 1 && 2048 + 32 * ln.ks_max <= kTcStageBytes,
               "scan_tc: weight operand or motif length outside the kernel's limits");
  static const bool pack = [] { const char* e = getenv("MEPP_TC_PACK"); return e ? atoi(e) != 0 : true; }();
  int grid = tc_scan_grid();
  if ((int64_t)grid > a.n_mtiles) grid = (int)a.n_mtiles;
  if (tc_units_per_tile() <= 16) {
    for (int k = 0; k < ln.n_tiles; ++k)
      MEPP_REQUIRE(ln.tiles[k].bn <= kTc4Cols / tc_split(), "scan_tc: N-tile wider than an accumulator stage of the four-stage form");
    // MEPP_TC_ISSUERS=4: one issuer per accumulator stage (tc_scan4i_kernel); 2: two issuers (tc_scan4_kernel)
    const int issuers = tc_issuers();
    if (issuers == 4 && pack && a.dbg_acc == nullptr && a.dbg_clk == nullptr && (a.mode & 255) <= 15) {
      // (modes: 1 no TMEM reads / tests, 2 no products, 4 reads but no tests, 8 half of the reads)
      const int md = a.mode & 15;
      void (*fn)(const TcLaunch, const TcScanArgs);
      if (tc_split() == 2) fn = md == 0 ? tc_scan4i_kernel<true, 0, 2> : md == 1 ? tc_scan4i_kernel<true, 1, 2> : md == 4 ? tc_scan4i_kernel<true, 4, 2> : tc_scan4i_kernel<true, 3, 2>;
      else fn = md == 0 ? tc_scan4i_kernel<true, 0, 1> : md == 1 ? tc_scan4i_kernel<true, 1, 1> : md == 2 ? tc_scan4i_kernel<true, 2, 1>
              : md == 3 ? tc_scan4i_kernel<true, 3, 1> : md == 4 ? tc_scan4i_kernel<true, 4, 1> : md == 8 ? tc_scan4i_kernel<true, 8, 1>
              : tc_scan4i_kernel<true, 12, 1>;
      MEPP_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      fn<<<grid, kTc4iThreads, smem, st>>>(ln, a);
    } else if (pack) {
      MEPP_CUDA_CHECK(cudaFuncSetAttribute(tc_scan4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tc_scan4_kernel<true><<<grid, kTcThreads, smem, st>>>(ln, a);
    } else {
      MEPP_CUDA_CHECK(cudaFuncSetAttribute(tc_scan4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tc_scan4_kernel<false><<<grid, kTcThreads, smem, st>>>(ln, a);
    }
    MEPP_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  if (pack) {
    MEPP_CUDA_CHECK(cudaFuncSetAttribute(tc_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_scan_kernel<true><<<grid, kTcThreads, smem, st>>>(ln, a);
  } else {
    MEPP_CUDA_CHECK(cudaFuncSetAttribute(tc_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_scan_kernel<false><<<grid, kTcThreads, smem, st>>>(ln, a);
  }
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

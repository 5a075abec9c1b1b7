// K1: PWM scan of packed DNA (3-bit symbols: 2-bit code + N flag), both Conv1D filters in one pass, threshold +
// ReLU, +-margin blur, fp16 hi/lo split of the blurred score written straight into
// the GEMM's tiled A operand, plus the per-position / per-sequence side outputs.
//
// Replaces (reference npdeloss/mepp): the Keras conv model of
// mepp/motif_layers.py:80-165 (normalise one-hot, Conv1D, reduce_max over the two
// filters, ZeroPadding1D((m-1)//2, ...), ReLU), the AveragePooling1D blur of
// mepp/core.py:59-145, the raw-moment sums over x of mepp/core.py:238-253 and the
// count / density reductions of mepp/core.py:415-470.
//
// Float32 summation order of a score (declared canonical order, identical to oracle/profile.py):
//   acc = 0; for tap j ascending: acc = fl32(acc + W[base_j][j]);  an N base adds the four
//   quarter weights .25*W[A..T][j] one after the other;  s = fl32(acc + fl32(-thr)).
//
// Two-level evaluation.  Matches are rare (p = 1e-4 per strand), so every window first goes
// through a conservative 15-bit integer pre-filter (host.cu: build_filter_table): pairs of
// taps index a 64-entry table (6 bits = two 3-bit symbols, N included) whose entries hold BOTH
// filters' quantised partial sums in one 32-bit word, so one LDS + one IADD advance two taps
// of two strands (the 16 ACGT-only entries of a table sit in 16 distinct banks).
// Only windows whose integer sum reaches the per-filter bound are queued and evaluated
// exactly in the canonical float32 order, warp-compacted so that all 32 lanes work on
// candidates.  The filter has no false negatives by construction, so
// every reported score comes from the exact path and is bit-identical to the oracle.
//
// Work decomposition: one warp = one scan unit (a task, or the '+/-' task of a motif together with its '+'
// twin) x one k-block (32 sequences); in the filter pass one lane = one sequence walking the positions left
// to right, eight windows per step, its 3-bit symbol stream read straight from global memory (word-transposed,
// so a warp's load is one 128-byte line) and prefetched one refill ahead.  Warps share nothing; shared memory
// only holds the unit's pair tables, its x0 staging rows and the candidate ring, ~14 KB per warp.
//
// scan_kernel<true> is the optional bit-parallel variant (mepp_scan_bitpar): a first stage that counts "bad" letters
// with logic ops on bit planes of the sequences, the pair-table sums as a second stage, then the same exact
// evaluation and the same phase B.  Identical outputs; measured slower on cfg 2, off by default.
#include <stdlib.h>
#include "common.cuh"
#include "tcscan.cuh"
#include <utility>
#include <vector>

namespace mepp {

struct HitRec { int32_t task, seq, pos; float x0; };
struct XEntry { int32_t row, seq; float x; uint32_t pad; };   // one non-zero of the blurred score matrix

constexpr int kScanWarps = 2;   // warps (= scan units) per CTA
constexpr int kRowStride = 36;  // floats per position row of the x0 staging buffer (32 + 4 pad)
constexpr int kQueue = 64;      // candidate ring (flushed whenever 32 are pending)
constexpr int kGroup = 8;       // windows per pre-filter step
constexpr int kAddr = 2 * (MEPP_MAX_MOTIF_LEN / 2) + 6;  // pair-index addresses a step can need

struct ScanArgs {
  const uint32_t* sym_T; const float* zhat;
  int64_t N, n_pad, KB;
  int L, T, margin, W3, qslots;
  const float* lut; const float* bias; const int32_t* mlen; const int32_t* nfilt;
  const uint32_t* qtab; const uint32_t* qthr;
  const int32_t* units; int U;   // scan units: (task scanned, twin task fed by the same pass or -1); null = one per task
  uint8_t* x_planes; double* rowstats; int32_t* count; int32_t* density;
  HitRec* hits; int64_t hits_cap; unsigned long long* hit_count;
  // sparse form of X (optional): MEPP_ENTRY_LISTS sub-lists of sub_cap records, list = k-block % lists (one
  // counter per 128-byte line, so that the appends do not serialise on one L2 address)
  XEntry* entries; int64_t sub_cap; unsigned long long* entry_count; int32_t* row_nnz;
  // bit-parallel first stage (optional): one-hot bit planes of the sequences, per-task letter sets, and the
  // number of units of this launch's list (device side: the list is partitioned by a kernel, see bp_partition_kernel)
  const uint32_t* planes_T; const uint32_t* bp; const int32_t* unit_count;
  // lean mode (optional, sparse path): the scan only finds and publishes matches; they also go to fixed-capacity
  // buckets per (task, k-block), from which bucket_entries_kernel builds the blurred entries and the row sums
  int lean; uint2* buckets; int32_t* bucket_count;
  // overflow of the buckets (optional, tensor-core scan): matches beyond a bucket's capacity go to ONE list
  // {bucket, key, x0}; bucket_entries_kernel lists the buckets that overflowed and bucket_overflow_kernel redoes those
  // from both places.  ovf_count[0] = records, [1] = overflowed buckets.
  uint4* ovf; uint32_t* ovf_count; uint32_t ovf_cap; int32_t* ovf_bkts; uint32_t ovf_bkt_cap;
  // tiled operand (optional): per-task power-of-two scale of the fp16 hi + lo pieces (mepp_x_scales); null = MEPP_X_SCALE
  const float* xscale;
};
constexpr int kBucketCap = 64;   // matches per (task, k-block) bucket; more go to the overflow list (tensor-core scan) or
                                 // send the group through the full scan (CUDA-core lean scan)
constexpr int kOvfMax = 4096;    // matches of one overflowed bucket that bucket_overflow_kernel sorts in shared memory

// Per-warp (= per unit) state shared by the helper routines below.  The helpers are deliberately NOT
// inlined and NOT templated on the task: one small instruction stream for every task keeps the kernel
// inside the instruction cache (an earlier variant specialised on motif length and strand count, 26k SASS
// instructions, and spent most of its time on instruction fetch).
struct TaskCtx {
  const float2* s_lut;     // exact float32 weights of the scanned task (shared copy)
  const uint32_t* g_sym;   // sym_T + first sequence of the k-block: word w of sequence s at [w * n_pad + s]
  float* s_buf; float* s_buf2; uint16_t* s_queue;
  int64_t n_pad;
  int t, t2, kb, m, padl, mg, L, dual;
  float bias;
  int64_t tile_stride;   // bytes between consecutive m-tiles of the A operand
  uint8_t* x_base;       // A operand + this k-block's offset inside an m-tile row
  unsigned int* flags;   // non-zero bit array [m_tile][FW]
  int FW;
  unsigned long long rowmask[2];   // staging rows that hold a match (bit b = row b), per staging buffer
};

// Exact canonical-order scores of the window starting at base i of sequence s of the k-block:
// relu(filter 0) and relu(max over the task's filters).
__device__ __forceinline__ void exact_x0(const TaskCtx& c, int s, int i, float& x_first, float& x_task) {
  const uint32_t* g = c.g_sym + s;
  const int bit0 = 3 * i, w0 = bit0 >> 5, sh = bit0 & 31;
  uint32_t x[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) x[k] = __ldg(g + (int64_t)(w0 + k) * c.n_pad);
  uint32_t r0 = __funnelshift_r(x[0], x[1], sh), r1 = __funnelshift_r(x[1], x[2], sh),
           r2 = __funnelshift_r(x[2], x[3], sh), r3 = __funnelshift_r(x[3], x[4], sh);
  float a0 = 0.0f, a1 = 0.0f;
  const int m = c.m;
  const float2* lut = c.s_lut;
  auto tap = [&](int j) {
    const uint32_t sym = r0 & 7u;
    r0 = __funnelshift_r(r0, r1, 3); r1 = __funnelshift_r(r1, r2, 3); r2 = __funnelshift_r(r2, r3, 3); r3 >>= 3;
    const float2* l4 = lut + j * 4;
    if (sym >= 4u) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const float2 w = l4[cc];
        a0 = __fadd_rn(a0, __fmul_rn(0.25f, w.x));
        a1 = __fadd_rn(a1, __fmul_rn(0.25f, w.y));
      }
    } else {
      const float2 w = l4[sym];
      a0 = __fadd_rn(a0, w.x);
      a1 = __fadd_rn(a1, w.y);
    }
  };
#pragma unroll 1
  for (int j = 0; j < m; ++j) tap(j);
  float sc = __fadd_rn(a0, c.bias);
  x_first = fmaxf(sc, 0.0f);
  if (c.dual) sc = fmaxf(sc, __fadd_rn(a1, c.bias));
  x_task = fmaxf(sc, 0.0f);
}

// One match of task t: staging buffer, per-position / per-sequence counts, optional match list.
__device__ __forceinline__ void publish_match(const ScanArgs& a, TaskCtx& c, int which, int t, float* buf, int b, int s,
                                              int p, float x0) {
  if (a.lean) {
    const int64_t bkt = (int64_t)t * a.KB + c.kb;
    const int slot = atomicAdd(&a.bucket_count[bkt], 1);
    if (slot < kBucketCap) a.buckets[bkt * kBucketCap + slot] = make_uint2(((uint32_t)s << 16) | (uint32_t)p, __float_as_uint(x0));
  } else {
    buf[b * kRowStride + s] = x0;
    atomicOr(&c.rowmask[which], 1ull << b);
  }
  const int64_t ns = (int64_t)c.kb * 32 + s;
  atomicAdd(&a.count[(int64_t)t * c.L + p], 1);
  atomicAdd(&a.density[(int64_t)t * a.n_pad + ns], 1);
  if (a.hit_count) {
    const unsigned long long slot = atomicAdd(a.hit_count, 1ull);
    if (a.hits && (int64_t)slot < a.hits_cap) {
      HitRec rec; rec.task = t; rec.seq = (int32_t)ns; rec.pos = p; rec.x0 = x0;
      a.hits[slot] = rec;
    }
  }
}

// Evaluate up to 32 queued candidates (one per lane) exactly; publish matches.  Returns (warp
// uniform) bit 0: the unit's task got a match, bit 1: its twin (the '+' task sharing filter 0) got one.
__device__ __noinline__ unsigned flush_candidates(const ScanArgs& a, TaskCtx& c, int q_head, int count, int base,
                                                  int lane) {
  __syncwarp();
  bool hit = false, hit2 = false;
  if (lane < count) {
    const uint32_t e = c.s_queue[(q_head + lane) & (kQueue - 1)];
    const int s = (int)(e >> 8), b = (int)(e & 0xffu);
    const int p = base - c.mg + b;
    float x_first, x_task;
    exact_x0(c, s, p - c.padl, x_first, x_task);
    if (x_task > 0.0f) {
      hit = true;
      publish_match(a, c, 0, c.t, c.s_buf, b, s, p, x_task);
    }
    if (c.t2 >= 0 && x_first > 0.0f) {
      hit2 = true;
      publish_match(a, c, 1, c.t2, c.s_buf2, b, s, p, x_first);
    }
  }
  return (__any_sync(0xffffffffu, hit) ? 1u : 0u) | (__any_sync(0xffffffffu, hit2) ? 2u : 0u);
}

// Phase B of one 32-position chunk that holds (or borders) matches: blur, row sums, fp16 split, tiled
// store of the non-zero 16-byte pieces; lane = (kc, rsub).  Chunks without any match store nothing.
__device__ __noinline__ void store_chunk(const ScanArgs& a, const TaskCtx& c, int t, const float* s_buf,
                                         unsigned long long rowmask, int base, int lane) {
  const unsigned full = 0xffffffffu;
  const int kc = lane & 3, rsub = lane >> 2;
  const int mg = c.mg, L = c.L;
  const int k = 1 + 2 * mg;
  const float kf = (float)k;
  const int64_t tile_stride = c.tile_stride;
  uint8_t* const x_base = c.x_base;
  const unsigned long long wmask = (1ull << (8 + 2 * mg)) - 1ull;   // staging rows feeding eight output rows
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    if (((rowmask >> (8 * it)) & wmask) == 0ull) continue;          // no match in this row group's blur window
    const int rl = rsub + 8 * it;
    const int p = base + rl;
    const bool valid = p >= 0 && p < L;
    uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = vh;
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
#pragma unroll 1
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
        for (int q = 0; q < 8; ++q)
          if (xb[q] != 0.0f) xb[q] = __fdiv_rn(xb[q], kf);   // (0 / k takes the slow division path)
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
        const int64_t n = (int64_t)c.kb * 32 + kc * 8 + q;     // (a padding sequence never holds a match)
        const double z = (a.zhat && xb[q] != 0.0f && n < a.N) ? (double)__ldg(a.zhat + n) : 0.0;
        sx += v; sxx += v * v; sxz += v * z;
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
      if (a.entries) {
        // sparse form of X: one record per non-zero, appended to the k-block's sub-list in warp-sized batches
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) cnt += xb[q] != 0.0f;
        int pre = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(full, pre, o);
          if (lane >= o) pre += v;
        }
        const int total = __shfl_sync(full, pre, 31);
        const int list = c.kb & (MEPP_ENTRY_LISTS - 1);
        unsigned long long slot0 = 0ull;
        if (lane == 0) slot0 = atomicAdd(a.entry_count + list * 16, (unsigned long long)total);
        slot0 = __shfl_sync(full, slot0, 0) + (unsigned long long)(pre - cnt);
        if (cnt != 0) {
          const int r = t * L + p;
          atomicAdd(a.row_nnz + r, cnt);
          uint4* dst = reinterpret_cast<uint4*>(a.entries + (int64_t)list * a.sub_cap);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (xb[q] != 0.0f) {
              if ((int64_t)slot0 < a.sub_cap)
                dst[slot0] = make_uint4((uint32_t)r, (uint32_t)(c.kb * 32 + kc * 8 + q), __float_as_uint(xb[q]), 0u);
              ++slot0;
            }
          }
        }
      }
    }
    if (nz && a.x_planes) {
      uint32_t h[4], l[4];
      const float xs = a.xscale ? __ldg(a.xscale + t) : (float)MEPP_X_SCALE;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        __half h0, l0, h1, l1;
        split_half(xb[2 * q] * xs, h0, l0);
        split_half(xb[2 * q + 1] * xs, h1, l1);
        h[q] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        l[q] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      }
      vh = make_uint4(h[0], h[1], h[2], h[3]);
      vl = make_uint4(l[0], l[1], l[2], l[3]);
    }
    if (nz && valid && a.x_planes) {   // only pieces that hold a non-zero value are stored: the operand is all-zero on entry
      const int r = t * L + p;                         // global row (< 2^31, checked on the host)
      uint8_t* dst = x_base + (int64_t)(r >> 7) * tile_stride + ((kc * kBlockM + (r & 127)) << 4);
      *reinterpret_cast<uint4*>(dst) = vh;
      *reinterpret_cast<uint4*>(dst + kChunks * kBlockM * 16) = vl;
      const int bit = 2 * c.kb + (kc >> 1);            // mark the K=16 block as non-zero
      atomicOr(c.flags + (int64_t)(r >> 7) * c.FW + (bit >> 5), 1u << (bit & 31));
    }
  }
}

// Shared-memory address of the pair-table entry selected by the two symbols at bits [3K, 3K+6) of the
// shift register: (index << 2) | table base.  `qbase` is the 256-byte aligned shared address of the
// unit's tables (slot T = 256 bytes at qbase + 256 T, the slot offset goes into the LDS immediate), so the
// LOP3 that masks the 6-bit index also ORs the base in.
template <int K>
__device__ __forceinline__ uint32_t pair_addr(uint32_t qbase, const uint32_t (&r)[6]) {
  constexpr int bit = 3 * K, word = bit / 32, off = bit % 32;
  uint32_t v;
  if constexpr (off + 6 > 32) v = __funnelshift_r(r[word], r[word + 1], off - 2);
  else if constexpr (off >= 2) v = r[word] >> (off - 2);
  else v = r[word] << (2 - off);
  return (v & 0xFCu) | qbase;
}
template <int K0, int K1>
__device__ __forceinline__ void pair_addrs(uint32_t qbase, const uint32_t (&r)[6], uint32_t (&A)[kAddr]) {
  if constexpr (K0 < K1) {
    A[K0] = pair_addr<K0>(qbase, r);
    pair_addrs<K0 + 1, K1>(qbase, r, A);
  }
}
template <int T>
__device__ __forceinline__ uint32_t lds_slot(uint32_t addr) {
  uint32_t x;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(x) : "r"(addr), "n"(T * 256));
  return x;
}
// Slots [T0, T1) of the eight windows of a step.  Window U slot T reads taps U+2T, U+2T+1, i.e. the
// pair index at bits [3(U+2T), +6): the same index serves window U slot T and window U+2 slot T-1, so a
// step extracts 2 np + 6 indices for its 8 np lookups.  One basic block of independent loads.
template <int T0, int T1>
__device__ __forceinline__ void add_slots(const uint32_t (&A)[kAddr], uint32_t (&s)[kGroup]) {
  if constexpr (T0 < T1) {
    s[0] += lds_slot<T0>(A[2 * T0 + 0]); s[1] += lds_slot<T0>(A[2 * T0 + 1]);
    s[2] += lds_slot<T0>(A[2 * T0 + 2]); s[3] += lds_slot<T0>(A[2 * T0 + 3]);
    s[4] += lds_slot<T0>(A[2 * T0 + 4]); s[5] += lds_slot<T0>(A[2 * T0 + 5]);
    s[6] += lds_slot<T0>(A[2 * T0 + 6]); s[7] += lds_slot<T0>(A[2 * T0 + 7]);
    add_slots<T0 + 1, T1>(A, s);
  }
}
template <int T0, int T1>
__device__ __forceinline__ void ladder_step(uint32_t qbase, const uint32_t (&r)[6], uint32_t (&A)[kAddr],
                                            uint32_t (&s)[kGroup]) {
  pair_addrs<(T0 == 0 ? 0 : 2 * T0 + 6), 2 * T1 + 6>(qbase, r, A);
  add_slots<T0, T1>(A, s);
}
// The eight packed window sums of a step (np = pair slots in use); the slot-count ladder is walked once for
// all eight windows.
__device__ __forceinline__ void window_sums(uint32_t qbase, uint32_t qinit, int np, const uint32_t (&r)[6],
                                            uint32_t (&s)[kGroup]) {
  uint32_t A[kAddr];
#pragma unroll
  for (int u = 0; u < kGroup; ++u) s[u] = qinit;
  ladder_step<0, 3>(qbase, r, A, s);
  if (np > 3) {
    ladder_step<3, 5>(qbase, r, A, s);
    if (np > 5) {
      ladder_step<5, 6>(qbase, r, A, s);
      if (np > 6) {
        ladder_step<6, 8>(qbase, r, A, s);
        if (np > 8) {
          ladder_step<8, 11>(qbase, r, A, s);
          if (np > 11) ladder_step<11, 16>(qbase, r, A, s);
        }
      }
    }
  }
}
// One window at the register file's origin (tail of the last chunk only).
__device__ __noinline__ uint32_t window_sum1(uint32_t qbase, uint32_t qinit, int np, uint32_t r0, uint32_t r1,
                                             uint32_t r2, uint32_t r3) {
  uint32_t sum = qinit;
#pragma unroll 1
  for (int T = 0; T < np; ++T) {
    const uint32_t idx = r0 & 63u;
    uint32_t x;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(qbase + (uint32_t)T * 256u + idx * 4u));
    sum += x;
    r0 = __funnelshift_r(r0, r1, 6); r1 = __funnelshift_r(r1, r2, 6); r2 = __funnelshift_r(r2, r3, 6); r3 >>= 6;
  }
  return sum;
}

template <int BITS>
__device__ __forceinline__ void advance(uint32_t (&r)[6]) {
  r[0] = __funnelshift_r(r[0], r[1], BITS); r[1] = __funnelshift_r(r[1], r[2], BITS);
  r[2] = __funnelshift_r(r[2], r[3], BITS); r[3] = __funnelshift_r(r[3], r[4], BITS);
  r[4] = __funnelshift_r(r[4], r[5], BITS); r[5] >>= BITS;
}

// ---- bit-parallel first stage ------------------------------------------------------------------------
// A match needs  sum_j pen_j(letter_j) < D,  pen_j(c) = max_c' W[c'][j] - W[c][j] >= 0,  D = sum_j max_c W[c][j] -
// threshold + rounding slack.  With a cut d0, letter c is "bad" at tap j when pen_j(c) > d0; taps with one or two
// good letters are counted.  A window with K or more bad letters at counted taps has a penalty of at least the K
// smallest per-tap minima of the bad penalties, and (d0, K) are chosen (bp_tables_kernel) so that this bound
// exceeds D: such a window cannot match -- no false negatives.  N bases are never bad (their penalty is only
// known to be >= 0).  Lane = sequence; a register word holds 32 window starts.  The sequences are staged as four
// "a letter other than c" bit planes, so the bad bits of a counted tap with good letter g are plane g shifted to
// the tap's base (the AND of two planes for two good letters): per tap and filter one shift and two logic ops
// on the 2-bit saturating counter, the plane words coming from shared memory (the load/store pipe is idle
// here, the integer pipe is the limiter of the pair tables).  Survivors (typically < 1 % of the windows) go to the
// exact evaluation, so every reported score still comes from the canonical float32 order.
struct BpUnit {
  int K, dual;                 // K = 0: no window of this unit can match
  int n1[2], n2[2];            // counted taps per filter: one good letter / two good letters
};
// descriptors in shared memory: [filter][32] uint4 = (byte offset of plane gA, tap, byte offset of plane gB, -)
__device__ __forceinline__ void bp_count(const uint4* __restrict__ desc, int n1, int n2, const uint8_t* __restrict__ my_words,
                                         uint32_t& c0, uint32_t& c1) {
#pragma unroll 4
  for (int i = 0; i < n1; ++i) {
    const uint2 d = *reinterpret_cast<const uint2*>(desc + i);
    const uint2 q = *reinterpret_cast<const uint2*>(my_words + d.x);
    const uint32_t bad = __funnelshift_r(q.x, q.y, d.y);
    const uint32_t t = c1 | (c0 & bad);
    c0 = (c0 ^ bad) | (c1 & c0); c1 = t;
  }
#pragma unroll 2
  for (int i = n1; i < n1 + n2; ++i) {
    const uint4 d = desc[i];
    const uint2 qa = *reinterpret_cast<const uint2*>(my_words + d.x);
    const uint2 qb = *reinterpret_cast<const uint2*>(my_words + d.z);
    const uint32_t bad = __funnelshift_r(qa.x, qa.y, d.y) & __funnelshift_r(qb.x, qb.y, d.y);
    const uint32_t t = c1 | (c0 & bad);
    c0 = (c0 ^ bad) | (c1 & c0); c1 = t;
  }
}
// Survivor mask of the 32 windows starting at base 32 * step: lo / hi = plane words of bases [32 step, 32 step + 64).
// s_words: this warp's staging area [4 planes][32 lanes] uint2 (lo, hi): each lane only touches its own slots.
__device__ __forceinline__ uint32_t bp_survivors(const uint32_t (&lo)[4], const uint32_t (&hi)[4], const BpUnit& u,
                                                 const uint4* __restrict__ s_desc, uint2* __restrict__ s_words, int lane) {
#pragma unroll
  for (int c = 0; c < 4; ++c) s_words[c * 32 + lane] = make_uint2(lo[c], hi[c]);
  const uint8_t* my_words = reinterpret_cast<const uint8_t*>(s_words + lane);
  uint32_t a0 = 0u, a1 = 0u, b0 = 0u, b1 = 0u;   // counters of filter 0 (a) and filter 1 (b), bit-sliced
  bp_count(s_desc, u.n1[0], u.n2[0], my_words, a0, a1);
  const uint32_t rej0 = u.K == 1 ? (a0 | a1) : u.K == 2 ? a1 : (a0 & a1);
  if (!u.dual) return ~rej0;
  bp_count(s_desc + 32, u.n1[1], u.n2[1], my_words, b0, b1);
  const uint32_t rej1 = u.K == 1 ? (b0 | b1) : u.K == 2 ? b1 : (b0 & b1);
  return ~(rej0 & rej1);
}

// Second stage for the survivors of the bit-parallel counters: the conservative integer sums of the pair tables
// (host.cu: build_filter_table -- the same tables, bounds and "no false negatives" argument as the pair-table kernel),
// one queued window per lane, the tables read from global memory (L1).  Windows that cannot reach their bound are
// cleared from the survivor bitmap.  Queue entry = lane << 11 | step << 5 | bit (window start 32 step + bit).
__device__ __noinline__ void bp_second_stage(const uint32_t* __restrict__ g_sym, int64_t n_pad, const uint32_t* __restrict__ tab,
                                             uint32_t qinit, int np, const uint16_t* __restrict__ s_queue, int q_head,
                                             int count, uint32_t* __restrict__ s_surv, int lane) {
  __syncwarp();
  if (lane < count) {
    const uint32_t e = s_queue[(q_head + lane) & (kQueue - 1)];
    const int s = (int)(e >> 11), step = (int)((e >> 5) & 63u), bit = (int)(e & 31u);
    const int bit0 = 3 * (32 * step + bit), w0 = bit0 >> 5, sh = bit0 & 31;
    const uint32_t* g = g_sym + s;
    uint32_t x[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) x[k] = __ldg(g + (int64_t)(w0 + k) * n_pad);
    uint32_t r0 = __funnelshift_r(x[0], x[1], sh), r1 = __funnelshift_r(x[1], x[2], sh),
             r2 = __funnelshift_r(x[2], x[3], sh), r3 = __funnelshift_r(x[3], x[4], sh);
    uint32_t sum = qinit;
#pragma unroll 2
    for (int T = 0; T < np; ++T) {
      sum += __ldg(tab + T * 64 + (r0 & 63u));
      r0 = __funnelshift_r(r0, r1, 6); r1 = __funnelshift_r(r1, r2, 6); r2 = __funnelshift_r(r2, r3, 6); r3 >>= 6;
    }
    if ((sum & 0x80008000u) == 0u) atomicAnd(&s_surv[step * 32 + s], ~(1u << bit));
  }
  __syncwarp();
}

template <bool BP>
__device__ __forceinline__ void scan_task(const ScanArgs& a, const int t, const int t2, const int kb,
                                          uint32_t* __restrict__ s_qtab, float2* s_lut,
                                          float* __restrict__ s_buf,
                                          float* __restrict__ s_buf2, uint16_t* __restrict__ s_queue,
                                          uint32_t* __restrict__ s_ring, TaskCtx* __restrict__ s_ctx,
                                          const int lane) {
  const unsigned full = 0xffffffffu;
  const int L = a.L, mg = a.margin;
  const int m = a.mlen[t];
  const int np = (m + 1) >> 1;           // pair slots in use
  // the slot-count ladder of window_sums reads whole buckets: 3, 5, 6, 8, 11 or 16 slots
  const int nslot = np <= 3 ? 3 : np <= 5 ? 5 : np <= 6 ? 6 : np <= 8 ? 8 : np <= 11 ? 11 : 16;
  if (nslot > a.qslots) __trap();        // contract of mepp_scan: max_motif_len bounds every task
  const int padl = (m - 1) / 2;          // motif_layers.py:99-101
  const int n_valid = L - m + 1;         // 'valid' convolution outputs
  const bool real = (int64_t)kb * 32 + lane < a.N;
  const int nbuf = 32 + 2 * mg;

  BpUnit bpu;
  bpu.K = 0; bpu.dual = 0; bpu.n1[0] = bpu.n1[1] = bpu.n2[0] = bpu.n2[1] = 0;
  if constexpr (BP) {
    // counted taps of the unit (bp_tables_kernel): word f * 32 + i of the task's record = tap | gA << 8 | gB << 16
    const uint32_t* rec = a.bp + (size_t)t * MEPP_BP_WORDS;
    uint4* dst = reinterpret_cast<uint4*>(s_qtab);
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const uint32_t w = __ldg(rec + f * 32 + lane);
      dst[f * 32 + lane] = make_uint4(((w >> 8) & 3u) * 256u, w & 0xffu, ((w >> 16) & 3u) * 256u, 0u);
    }
    const uint32_t h = __ldg(rec + 64), cn = __ldg(rec + 65);
    bpu.K = (int)(h & 0xffu); bpu.dual = ((h >> 16) & 0xffu) == 2u;
    bpu.n1[0] = (int)(cn & 0xffu); bpu.n2[0] = (int)((cn >> 8) & 0xffu);
    bpu.n1[1] = (int)((cn >> 16) & 0xffu); bpu.n2[1] = (int)(cn >> 24);
  } else {
    const uint4* src = reinterpret_cast<const uint4*>(a.qtab + (size_t)t * 1024);
    uint4* dst = reinterpret_cast<uint4*>(s_qtab);
    for (int i = lane; i < nslot * 16; i += 32) dst[i] = __ldg(src + i);   // (unused slots are zero)
  }
  {
    const float2* g_lut = reinterpret_cast<const float2*>(a.lut) + (size_t)t * (MEPP_MAX_MOTIF_LEN * 4);
    // the exact weights: a shared copy for the bit-parallel kernel (its exact path is busy); the pair-table kernel
    // evaluates ~1e-4 of the windows exactly and reads them from global memory (L1), which frees up to 2 KB per CTA
    if constexpr (BP)
      for (int i = lane; i < m * 4; i += 32) s_lut[i] = __ldg(g_lut + i);
    else
      s_lut = const_cast<float2*>(g_lut);
  }
  if (!a.lean) {
    for (int b = 0; b < nbuf; ++b) s_buf[b * kRowStride + lane] = 0.0f;
    if (t2 >= 0)
      for (int b = 0; b < nbuf; ++b) s_buf2[b * kRowStride + lane] = 0.0f;
  }
  // qinit: per 16-bit half (0x8000 - pass bound), so "sum reaches the bound" <=> bit 15 of the half is set.
  // Padding sequences never produce candidates.
  const uint32_t qinit = real ? a.qthr[t * 2 + 0] : 0u;
  const int64_t tile_stride = a.KB * (int64_t)kATileBytes;
  uint8_t* const x_base = a.x_planes + (int64_t)kb * kATileBytes;
  if (lane == 0) {
    TaskCtx w;
    w.s_lut = s_lut;
    w.g_sym = a.sym_T + (int64_t)kb * 32;
    w.s_buf = s_buf; w.s_buf2 = s_buf2; w.s_queue = s_queue;
    w.rowmask[0] = 0ull; w.rowmask[1] = 0ull;
    w.n_pad = a.n_pad;
    w.t = t; w.t2 = t2; w.kb = kb; w.m = m; w.padl = padl; w.mg = mg; w.L = L;
    w.dual = a.nfilt[t] == 2;
    w.bias = a.bias[t];
    w.tile_stride = tile_stride;
    w.x_base = x_base;
    w.FW = (int)flag_words(a.KB);
    w.flags = a.x_planes ? reinterpret_cast<unsigned int*>(
        a.x_planes + a_flags_offset(((int64_t)a.T * L + kBlockM - 1) / kBlockM, a.KB)) : nullptr;
    *s_ctx = w;
  }
  __syncwarp();
  TaskCtx& c = *s_ctx;

  const uint32_t* my = a.sym_T + (int64_t)kb * 32 + lane;   // this lane's sequence: word w at my[w * n_pad]
  const int64_t wstride = a.n_pad;
  const uint32_t qbase = (uint32_t)__cvta_generic_to_shared(s_qtab);

  // 192-bit shift register over the symbol stream (window start at bit 0 of r[0]), advanced by eight
  // bases (24 bits) per step and refilled with three words every 32 windows.  The refill words travel
  // global -> shared with cp.async one refill ahead (two stages of 3 x 32 words per warp; every lane
  // copies and reads only its own words), so no register ever waits on a global load.
  // The register starts `lead` (0..7) windows before the first real one, so that the first chunk holds a
  // multiple of eight windows (32 + mg - padl + lead) like every later chunk (32): steps never straddle a
  // refill and chunk rows stay aligned with the 8-row groups of the tiled stores.  The lead windows read
  // zero bits in front of the stream; their candidates are dropped (rows below the first real one).
  const int lead = (kGroup - ((32 + mg - padl) % kGroup)) % kGroup;
  const int lsh = 3 * lead;              // left shift of the stream inside the register (0..21 bits)
  uint32_t r[6] = {0u, 0u, 0u, 0u, 0u, 0u}, carry = 0u;
  if constexpr (!BP) {
    uint32_t w[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) w[k] = __ldg(my + (int64_t)k * wstride);
    r[0] = w[0] << lsh;
#pragma unroll
    for (int k = 1; k < 6; ++k) r[k] = __funnelshift_l(w[k - 1], w[k], lsh);
    carry = w[5];
  }
  // ---- bit-parallel pre-pass: survivors of the whole sequence, then the integer second stage in full batches ----
  uint32_t* const s_surv = s_ring + 256;   // [steps + 2][32 lanes]: bit k of word step = window start 32 step + k survives
  if constexpr (BP) {
    const int n_steps = (n_valid + 31) >> 5;
    const uint32_t* pw = a.planes_T + (int64_t)kb * 32 + lane + 4 * wstride;   // plane word 1 = bases 0..31
    uint32_t lo[4], hi[4], nx[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) { lo[cc] = __ldg(pw + cc * wstride); hi[cc] = __ldg(pw + (4 + cc) * wstride); }
    pw += 8 * wstride;
#pragma unroll 1
    for (int step = 0; step < n_steps; ++step) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) nx[cc] = __ldg(pw + cc * wstride);      // next step's words (zero padding behind L)
      pw += 4 * wstride;
      uint32_t surv = 0u;
      if (bpu.K != 0) {
        surv = bp_survivors(lo, hi, bpu, reinterpret_cast<const uint4*>(s_qtab), reinterpret_cast<uint2*>(s_ring), lane);
        const int left_w = n_valid - 32 * step;                               // real windows of this step
        if (left_w < 32) surv &= 0xffffffffu >> (32 - left_w);
        if (!real) surv = 0u;                                                 // padding sequences produce no candidates
      }
      s_surv[step * 32 + lane] = surv;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) { lo[cc] = hi[cc]; hi[cc] = nx[cc]; }
    }
    s_surv[n_steps * 32 + lane] = 0u;
    s_surv[(n_steps + 1) * 32 + lane] = 0u;
    __syncwarp();
    if (bpu.K != 0) {
      const uint32_t* tab = a.qtab + (size_t)t * 1024;
      const uint32_t qinit2 = a.qthr[t * 2 + 0];
      int qh = 0, qt = 0;
#pragma unroll 1
      for (int step = 0; step < n_steps; ++step) {
        uint32_t w = s_surv[step * 32 + lane];
        while (__any_sync(full, w != 0u)) {
          const bool cand = w != 0u;
          const int bb = __ffs((int)w) - 1;
          w &= w - 1u;
          const unsigned bal = __ballot_sync(full, cand);
          if (cand) s_queue[(qt + __popc(bal & ((1u << lane) - 1u))) & (kQueue - 1)] = (uint16_t)((lane << 11) | (step << 5) | bb);
          qt += __popc(bal);
          if (qt - qh >= 32) {
            bp_second_stage(a.sym_T + (int64_t)kb * 32, a.n_pad, tab, qinit2, np, s_queue, qh, 32, s_surv, lane);
            qh += 32;
          }
        }
      }
      if (qt != qh) bp_second_stage(a.sym_T + (int64_t)kb * 32, a.n_pad, tab, qinit2, np, s_queue, qh, qt - qh, s_surv, lane);
    }
    __syncwarp();
  }
  const uint32_t* pnext = my + 6 * wstride;
  const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(s_ring + lane);
  uint32_t stage = 0u;                   // byte offset of the stage holding the next refill (0 / 384)
  auto prefetch = [&](uint32_t st) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ring0 + st + (uint32_t)k * 128u),
                   "l"(pnext + (int64_t)k * wstride) : "memory");
    pnext += 3 * wstride;
  };
  if constexpr (!BP) prefetch(0u);
  int left = 32;
  int q_head = 0, q_tail = 0;            // warp-uniform ring indices
  unsigned dcar = 0u, dbody = 0u;        // staging rows [0, 2 mg) / [2 mg, nbuf) may hold non-zero x0 (bit 1: twin)

  for (int base = 0; base < L; base += 32) {
    if constexpr (BP) {
      // nothing carried over and no surviving window in reach of this chunk: nothing to evaluate, blur or store
      if ((dcar | dbody) == 0u) {
        const int ia = max(base - mg - padl + (base ? 2 * mg : 0), 0), ib = min(base - mg - padl + 32 + 2 * mg, n_valid);
        uint32_t found = 0u;
        if (ib > ia) {
          const int wa = ia >> 5, wb = (ib - 1) >> 5;        // at most three words
          for (int w = wa; w <= wb; ++w) {
            uint32_t v = s_surv[w * 32 + lane];
            if (w == wa) v &= 0xffffffffu << (ia & 31);
            if (w == wb) v &= 0xffffffffu >> (31 - ((ib - 1) & 31));
            found |= v;
          }
        }
        if (!__any_sync(full, found != 0u)) continue;
      }
    }
    int b_start = 0;
    if (base != 0) {
      b_start = 2 * mg;
      // staging rows: carry the last 2*margin rows to the front, clear the rest (only what can be non-zero)
#pragma unroll 1
      for (unsigned bit = 1u; bit <= (t2 >= 0 ? 2u : 1u); bit <<= 1) {
        float* buf = bit == 1u ? s_buf : s_buf2;
        if (dbody & bit) {
          for (int b = 0; b < 2 * mg; ++b) buf[b * kRowStride + lane] = buf[(32 + b) * kRowStride + lane];
          for (int b = 2 * mg; b < nbuf; ++b) buf[b * kRowStride + lane] = 0.0f;
          dcar |= bit; dbody &= ~bit;
          if (lane == 0) c.rowmask[bit >> 1] >>= 32;
        } else if (dcar & bit) {
          if (lane == 0) c.rowmask[bit >> 1] = 0ull;
          for (int b = 0; b < 2 * mg; ++b) buf[b * kRowStride + lane] = 0.0f;
          dcar &= ~bit;
        }
      }
    }
    // ---- phase A: integer pre-filter over positions base-mg+b, candidates queued ----
    // window starts i = base - mg + b - padl; only 0 <= i < n_valid are real windows
    const int i_first = base - mg + b_start - padl;
    const int b_lo = b_start + (i_first < 0 ? -i_first : 0);   // first real window of the chunk
    int b = base == 0 ? b_lo - lead : b_lo;
    int b_hi = nbuf;
    {
      const int i_end = base - mg + nbuf - padl;      // exclusive
      if (i_end > n_valid) b_hi -= i_end - n_valid;
    }
    auto refill = [&]() {
      asm volatile("cp.async.wait_all;" ::: "memory");
      uint32_t n[3];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(n[k]) : "r"(ring0 + stage + (uint32_t)k * 128u) : "memory");
      r[3] = __funnelshift_l(carry, n[0], lsh); r[4] = __funnelshift_l(n[0], n[1], lsh);
      r[5] = __funnelshift_l(n[1], n[2], lsh);
      carry = n[2];
      stage ^= 384u;
      prefetch(stage);
      left = 32;
    };
    // queue the candidates of buffer row bb (ballot `bal`), flushing the ring whenever 32 are pending
    auto push = [&](unsigned bal, bool cand, int bb) {
      if (cand) {
        const int slot_i = q_tail + __popc(bal & ((1u << lane) - 1u));
        s_queue[slot_i & (kQueue - 1)] = (uint16_t)((lane << 8) | bb);
      }
      q_tail += __popc(bal);
      if (q_tail - q_head >= 32) {
        const unsigned got = flush_candidates(a, c, q_head, 32, base, lane);
        if (!a.lean) dbody |= got;
        q_head += 32;
      }
    };
    if constexpr (BP) {
      // ---- phase A, bit-parallel: the survivors of both stages were found up front (bitmap over window starts) ----
#pragma unroll 1
      for (int b0 = b_start; b0 < nbuf; b0 += 32) {
        const int r_lo = max(b_lo - b0, 0), r_hi = min(b_hi - b0, 32);   // rows of this step that are real windows
        if (r_hi <= r_lo) continue;
        const uint32_t valid = (0xffffffffu >> (32 - (r_hi - r_lo))) << r_lo;
        const int i0 = base - mg - padl + b0;                            // window start of row b0 (>= -31)
        uint32_t surv;
        if (i0 >= 0) surv = __funnelshift_r(s_surv[(i0 >> 5) * 32 + lane], s_surv[((i0 >> 5) + 1) * 32 + lane], i0 & 31);
        else surv = s_surv[lane] << (-i0);
        surv &= valid;
        while (__any_sync(full, surv != 0u)) {
          const bool cand = surv != 0u;
          const int bb = __ffs((int)surv) - 1;
          surv &= surv - 1u;
          const unsigned bal = __ballot_sync(full, cand);
          push(bal, cand, b0 + bb);
        }
      }
    } else {
#pragma unroll 1
    for (; b + kGroup <= b_hi; b += kGroup) {
      uint32_t s[kGroup];
      window_sums(qbase, qinit, np, r, s);
      advance<3 * kGroup>(r);
      left -= kGroup;
      if (left == 0) refill();
      // pass bit = bit 15 / 31 of a packed sum
      const uint32_t any = (s[0] | s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7]) & 0x80008000u;
      if (__any_sync(full, any != 0u)) {
        // rare: merge the pass bits (window u -> bits 15-u and 31-u) and walk the windows that have candidates
        uint32_t cb = 0u;
#pragma unroll
        for (int u = 0; u < kGroup; ++u) cb |= (s[u] & 0x80008000u) >> u;
        uint32_t wm = __reduce_or_sync(full, cb);
        wm = (wm | (wm >> 16)) & 0xff00u;            // bit 15-u set: some lane has a candidate in window u
#pragma unroll 1
        while (wm != 0u) {
          const int u = 15 - (31 - __clz(wm));       // lowest window first
          wm &= ~(0x8000u >> u);
          const bool cand = (cb & (0x80008000u >> u)) != 0u;
          const unsigned bal = __ballot_sync(full, cand);
          if (b + u >= b_lo) push(bal, cand, b + u);
        }
      }
    }
#pragma unroll 1
    for (; b < b_hi; ++b) {                       // fewer than eight windows left: end of the sequence
      const bool cand = (window_sum1(qbase, qinit, np, r[0], r[1], r[2], r[3]) & 0x80008000u) != 0u;
      advance<3>(r);
      if (--left == 0) refill();
      const unsigned bal = __ballot_sync(full, cand);
      if (bal != 0u && b >= b_lo) push(bal, cand, b);
    }
    }
    if (q_tail != q_head) {
      const unsigned got = flush_candidates(a, c, q_head, q_tail - q_head, base, lane);
      if (!a.lean) dbody |= got;
      q_head = q_tail;
    }
    __syncwarp();
    // ---- phase B: blur + split + store, only for chunks that hold or border a match ----
    const unsigned nzb = dcar | dbody;
    if (nzb & 1u) store_chunk(a, c, t, s_buf, c.rowmask[0], base, lane);
    if (nzb & 2u) store_chunk(a, c, t2, s_buf2, c.rowmask[1], base, lane);
    __syncwarp();
  }
}

// words of a warp's table region: the pair tables, or (bit-parallel) 32 taps x 2 filters x uint4 letter masks
__host__ __device__ inline int scan_table_words(int qslots, bool bp) { return bp && qslots < 4 ? 256 : qslots * 64; }
__host__ __device__ inline int scan_ring_words(int L, bool bp) { return bp ? 256 + (L / 32 + 3) * 32 : 192; }

template <bool BP>
__global__ void __launch_bounds__(kScanWarps * 32) scan_kernel(const __grid_constant__ ScanArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = blockIdx.x;
  const int nbuf = 32 + 2 * a.margin;
  const int nb = a.units ? 2 : 1;                                                       // twin staging buffers
  const int tabw = scan_table_words(a.qslots, BP);
  // carve shared memory (the pair tables come first: the OR-merged index needs 256-byte alignment)
  uint8_t* sm = smem_raw + ((256u - (uint32_t)(__cvta_generic_to_shared(smem_raw) & 255u)) & 255u);
  uint32_t* s_qtabs = reinterpret_cast<uint32_t*>(sm);                                  // [warps][tabw]
  float* s_bufs = reinterpret_cast<float*>(s_qtabs + kScanWarps * tabw);                // [warps][nb][nbuf][36]
  const int bufw = a.lean ? 0 : nb * nbuf * kRowStride;                                 // (no staging rows in lean mode)
  // refill ring [2][3][32], or (bit-parallel) plane words [4][32] uint2 + survivor bitmap [L / 32 + 3][32]
  const int ringw = scan_ring_words(a.L, BP);
  uint32_t* s_rings = reinterpret_cast<uint32_t*>(s_bufs + kScanWarps * bufw);                     // [warps][ringw]
  float2* s_luts = reinterpret_cast<float2*>(s_rings + kScanWarps * ringw);                        // [warps][qslots * 8]
  const int lutw = BP ? a.qslots * 8 : 0;                                                         // (float2 per warp)
  uint16_t* s_queues = reinterpret_cast<uint16_t*>(s_luts + kScanWarps * lutw);                   // [warps][kQueue]
  __shared__ TaskCtx s_ctxs[kScanWarps];

  const int u = blockIdx.y * kScanWarps + warp;
  if (u >= (a.unit_count ? *a.unit_count : a.U)) return;
  const int t = a.units ? a.units[2 * u] : u;
  const int t2 = a.units ? a.units[2 * u + 1] : -1;
  float* s_buf = s_bufs + warp * bufw;
  float* s_buf2 = s_buf + (a.lean ? 0 : (nb - 1) * nbuf * kRowStride);
  scan_task<BP>(a, t, t2, kb, s_qtabs + warp * tabw, s_luts + warp * lutw, s_buf, s_buf2, s_queues + warp * kQueue,
                s_rings + warp * ringw, &s_ctxs[warp], lane);
}

// ---- tables of the bit-parallel first stage ----------------------------------------------------------
// One warp per task: tries every distinct penalty of filter 0 as the cut d0 (one candidate per lane and round); a
// candidate counts the taps with one or two good letters (n1, n2), takes the smallest K in {1, 2, 3} for which the
// sum of the K smallest per-tap minima of the bad penalties exceeds D (soundness, see above), and is rated by a cost
// model in integer-pipe cycles per 32 windows: counter updates + expected survivors on uniform DNA x the cost of an
// exact evaluation.  The unit takes the bit-parallel kernel when that cost is below cost_ratio x the modelled
// cost of the pair tables.  Filter 1 gets the same d0 and K; its bound is checked on its own penalties.
// Record of task t (MEPP_BP_WORDS words): words f * 32 + i = counted tap i of filter f: tap | gA << 8 | gB << 16
// (good letters; taps with one good letter first), word 64 = K | use << 8 | filters << 16 (K = 0: nothing can
// match), word 65 = n1[0] | n2[0] << 8 | n1[1] << 16 | n2[1] << 24, word 66 / 67 = survivor rate / cost (float bits).
__device__ __forceinline__ void bp_push3(double v, double& s1, double& s2, double& s3) {
  if (v < s1) { s3 = s2; s2 = s1; s1 = v; }
  else if (v < s2) { s3 = s2; s2 = v; }
  else if (v < s3) s3 = v;
}
__global__ void __launch_bounds__(128) bp_tables_kernel(const float* __restrict__ lut, const float* __restrict__ bias,
                                                        const int32_t* __restrict__ mlen, const int32_t* __restrict__ nfilt,
                                                        int T, double cost_ratio, uint32_t* __restrict__ bp) {
  __shared__ double s_pen[4][2][MEPP_MAX_MOTIF_LEN * 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + warp;
  if (t >= T) return;
  const unsigned full = 0xffffffffu;
  const int m = mlen[t], nf = nfilt[t];
  const float* l = lut + (size_t)t * (MEPP_MAX_MOTIF_LEN * 8);
  // per tap (lane = tap) and filter: penalties, column maximum and largest magnitude
  double maxT[2], absmax[2];
  for (int f = 0; f < 2; ++f) {
    double cmax = 0.0, amax = 0.0;
    if (lane < m) {
      double w[4];
      for (int c = 0; c < 4; ++c) w[c] = (double)l[(lane * 4 + c) * 2 + f];
      cmax = fmax(fmax(w[0], w[1]), fmax(w[2], w[3]));
      amax = fmax(fmax(fabs(w[0]), fabs(w[1])), fmax(fabs(w[2]), fabs(w[3])));
      for (int c = 0; c < 4; ++c) s_pen[warp][f][lane * 4 + c] = cmax - w[c];
    }
    for (int o = 16; o > 0; o >>= 1) {
      cmax += __shfl_xor_sync(full, cmax, o);
      amax += __shfl_xor_sync(full, amax, o);
    }
    maxT[f] = cmax; absmax[f] = amax;
  }
  __syncwarp();
  const double* pen = s_pen[warp][0];
  // float32 sequential sum vs real sum: <= 4 m roundings of at most ulp(absmax) / 2 each, doubled (host.cu: build_filter_table)
  double D[2];
  for (int f = 0; f < 2; ++f)
    D[f] = maxT[f] + (double)bias[t] + (2.0 * (4.0 * m) * absmax[f] * 5.97e-8 + 1e-6);   // bias = float(-threshold)
  double best_cost = 1e300, best_rate = 1.0;
  int best_q = 0x7fffffff, best_K = 0;
  for (int q = lane; q <= 4 * m; q += 32) {
    const double d0 = q == 0 ? 0.0 : pen[q - 1];
    double s1 = 1e300, s2 = 1e300, s3 = 1e300;            // three smallest per-tap minima of the bad penalties
    double p0 = 1.0, p1 = 0.0, p2 = 0.0;                  // P(0, 1, 2 bad letters so far) on uniform DNA
    int n1 = 0, n2 = 0;
    for (int j = 0; j < m; ++j) {
      int nbad = 0;
      double mb = 1e300;
      for (int c = 0; c < 4; ++c) {
        const double v = pen[j * 4 + c];
        if (v > d0) { ++nbad; mb = fmin(mb, v); }
      }
      if (nbad < 2) continue;                             // three or four good letters: the tap is not counted
      n1 += nbad == 3; n2 += nbad == 2;
      bp_push3(mb, s1, s2, s3);
      const double pj = 0.25 * nbad, qj = 1.0 - pj;
      p2 = p2 * qj + p1 * pj; p1 = p1 * qj + p0 * pj; p0 = p0 * qj;
    }
    int K = 0;
    double rate = 1.0;
    if (s1 < 1e299 && s1 > D[0]) { K = 1; rate = p0; }
    else if (s2 < 1e299 && s1 + s2 > D[0]) { K = 2; rate = p0 + p1; }
    else if (s3 < 1e299 && s1 + s2 + s3 > D[0]) { K = 3; rate = p0 + p1 + p2; }
    const double cost = nf * (8.0 * n1 + 14.0 * n2) + 6144.0 * rate * nf;   // (~6 cycles per survivor: queue + second stage)
    if (K && (cost < best_cost || (cost == best_cost && q < best_q))) {
      best_cost = cost; best_rate = rate; best_q = q; best_K = K;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double c2 = __shfl_xor_sync(full, best_cost, o), r2 = __shfl_xor_sync(full, best_rate, o);
    const int q2 = __shfl_xor_sync(full, best_q, o), K2 = __shfl_xor_sync(full, best_K, o);
    if (c2 < best_cost || (c2 == best_cost && q2 < best_q)) { best_cost = c2; best_rate = r2; best_q = q2; best_K = K2; }
  }
  uint32_t* rec = bp + (size_t)t * MEPP_BP_WORDS;
  int use = 0, n1[2] = {0, 0}, n2[2] = {0, 0};
  uint32_t desc[2] = {0u, 0u};
  if (!(D[0] > 0.0)) {
    best_K = 0; best_rate = 0.0; best_cost = 0.0; use = 1;   // the threshold is above every reachable score
  } else if (best_K > 0) {
    const double d0 = best_q == 0 ? 0.0 : pen[best_q - 1];
    bool sound = true;
    for (int f = 0; f < nf; ++f) {
      const double* pf = s_pen[warp][f];
      int nbad = 0, g[2] = {0, 0}, ng = 0;
      double mb = 1e300;
      if (lane < m)
        for (int c = 0; c < 4; ++c) {
          if (pf[lane * 4 + c] > d0) { ++nbad; mb = fmin(mb, pf[lane * 4 + c]); }
          else if (ng < 2) g[ng++] = c;
        }
      const unsigned single = __ballot_sync(full, nbad == 3), dbl = __ballot_sync(full, nbad == 2);
      n1[f] = __popc(single); n2[f] = __popc(dbl);
      const unsigned below = (1u << lane) - 1u;
      int pos = -1;
      uint32_t w = 0u;
      if (nbad == 3) { pos = __popc(single & below); w = (uint32_t)lane | ((uint32_t)g[0] << 8) | (0xffu << 16); }
      if (nbad == 2) { pos = n1[f] + __popc(dbl & below); w = (uint32_t)lane | ((uint32_t)g[0] << 8) | ((uint32_t)g[1] << 16); }
      // descriptor i of the filter goes to lane i
      for (int src = 0; src < 32; ++src) {
        const int p = __shfl_sync(full, pos, src);
        const uint32_t ww = __shfl_sync(full, w, src);
        if (p == lane) desc[f] = ww;
      }
      // the bound of this filter on its own penalties (filter 1 mirrors filter 0, so this only guards foreign tables)
      if (nbad < 2) mb = 1e300;
      double s1 = 1e300, s2 = 1e300, s3 = 1e300;
      for (int src = 0; src < 32; ++src) bp_push3(__shfl_sync(full, mb, src), s1, s2, s3);
      const double bound = best_K == 1 ? s1 : best_K == 2 ? (s2 < 1e299 ? s1 + s2 : 0.0) : (s3 < 1e299 ? s1 + s2 + s3 : 0.0);
      if (!(bound < 1e299 && bound > D[f])) sound = false;
    }
    const double cost_pairs = 32.0 * 1.55 * m;             // the pair tables: ~1.55 integer-pipe cycles per tap and window
    use = sound && best_cost <= cost_ratio * cost_pairs ? 1 : 0;
  }
  rec[lane] = desc[0];
  rec[32 + lane] = desc[1];
  if (lane == 0) {
    rec[64] = (uint32_t)best_K | ((uint32_t)use << 8) | ((uint32_t)nf << 16);
    rec[65] = (uint32_t)n1[0] | ((uint32_t)n2[0] << 8) | ((uint32_t)n1[1] << 16) | ((uint32_t)n2[1] << 24);
    rec[66] = __float_as_uint((float)best_rate);
    rec[67] = __float_as_uint((float)best_cost);
    for (int k = 68; k < MEPP_BP_WORDS; ++k) rec[k] = 0u;
  }
}

// Splits the unit list of a scan call by class (stable, so the host's ordering by motif length survives):
// out = [pair-table units (U rows)][bit-parallel units (U rows)], counts[0 / 1] = their numbers.  One warp.
__global__ void bp_partition_kernel(const int32_t* __restrict__ units, int U, const uint32_t* __restrict__ bp,
                                    int32_t* __restrict__ out, int32_t* __restrict__ counts) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x;
  int n0 = 0, n1 = 0;
  for (int u0 = 0; u0 < U; u0 += 32) {
    const int u = u0 + lane;
    int t = 0, t2 = -1;
    bool use = false;
    if (u < U) {
      t = units ? units[2 * u] : u;
      t2 = units ? units[2 * u + 1] : -1;
      use = ((bp[(size_t)t * MEPP_BP_WORDS + 64] >> 8) & 1u) != 0u;
    }
    const unsigned m1 = __ballot_sync(full, u < U && use), m0 = __ballot_sync(full, u < U && !use);
    const unsigned below = (1u << lane) - 1u;
    if (u < U) {
      int32_t* dst = use ? out + 2 * (U + n1 + __popc(m1 & below)) : out + 2 * (n0 + __popc(m0 & below));
      dst[0] = t; dst[1] = t2;
    }
    n0 += __popc(m0); n1 += __popc(m1);
  }
  if (lane == 0) { counts[0] = n0; counts[1] = n1; }
}

// Restores the all-zero state of the A operand after its consumer ran: every match (task, seq, pos) of the
// list dirtied the 16-byte pieces of rows pos-margin .. pos+margin in its sequence's 8-column group (both
// planes).  If the list overflowed (or there is none) the whole tile region is cleared instead.
__global__ void x_planes_reset_kernel(uint8_t* __restrict__ x_planes, const HitRec* __restrict__ hits,
                                      int64_t hits_cap, const unsigned long long* __restrict__ hit_count, int L,
                                      int margin, int64_t KB, int64_t tile_bytes) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  const int64_t n_hits = hits ? (int64_t)*hit_count : -1;
  if (n_hits < 0 || n_hits > hits_cap) {
    uint4* p = reinterpret_cast<uint4*>(x_planes);
    for (int64_t i = tid; i < tile_bytes / 16; i += nthr) p[i] = z4;
    return;
  }
  const int k = 2 * margin + 1;
  const int64_t tile_stride = KB * (int64_t)kATileBytes;
  for (int64_t i = tid; i < n_hits * k; i += nthr) {
    const HitRec h = hits[i / k];
    const int p = h.pos - margin + (int)(i % k);
    if (p < 0 || p >= L) continue;
    const int r = h.task * L + p;
    const int kb = h.seq >> 5, kc = (h.seq & 31) >> 3;
    uint8_t* dst = x_planes + (int64_t)(r >> 7) * tile_stride + (int64_t)kb * kATileBytes +
                   ((kc * kBlockM + (r & 127)) << 4);
    *reinterpret_cast<uint4*>(dst) = z4;
    *reinterpret_cast<uint4*>(dst + kChunks * kBlockM * 16) = z4;
  }
}

}  // namespace mepp

extern "C" int mepp_x_planes_reset(void* x_planes_dev, const void* hits_dev, int64_t hits_cap,
                                   const unsigned long long* hit_count_dev, int64_t N, int L, int T, int margin,
                                   void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(x_planes_dev != nullptr, "x_planes_reset: null pointer");
  MEPP_REQUIRE(N > 0 && L > 0 && T > 0 && margin >= 0 && margin <= MEPP_MAX_MARGIN, "x_planes_reset: bad arguments");
  MEPP_REQUIRE(hits_dev == nullptr || hit_count_dev != nullptr, "x_planes_reset: the match list needs its counter");
  MEPP_REQUIRE(mepp_device_ok(), "x_planes_reset: no sm_100 CUDA device (no CPU fallback)");
  cudaStream_t st = as_stream(stream);
  const int64_t KB = mepp_pad32(N) / kBlockK;
  const int64_t n_tiles_m = ((int64_t)T * L + kBlockM - 1) / kBlockM;
  const int64_t tile_bytes = a_flags_offset(n_tiles_m, KB);
  x_planes_reset_kernel<<<148 * 8, 256, 0, st>>>(reinterpret_cast<uint8_t*>(x_planes_dev),
                                                 reinterpret_cast<const HitRec*>(hits_dev), hits_cap, hit_count_dev, L,
                                                 margin, KB, tile_bytes);
  MEPP_CUDA_CHECK(cudaGetLastError());
  MEPP_CUDA_CHECK(cudaMemsetAsync(reinterpret_cast<uint8_t*>(x_planes_dev) + tile_bytes, 0,
                                  (size_t)(n_tiles_m * flag_words(KB) * 4), st));
  return 0;
}

namespace mepp {
// Lean mode, second half: one warp per (task, k-block) bucket of matches.  The matches are sorted by (sequence,
// position); a row of the blurred matrix belongs to the first match inside its window, which sums the window's
// matches in position order in float32 (adding the zeros of the dense window would change nothing), divides by k
// and emits the entry, the non-zero count of the row and the row sums -- the same values store_chunk produces
// from the staging rows (tools/proto_match_blur.py is the numpy form of this).  A bucket that overflowed is counted
// in entry_count[1] (the sub-list counters sit at [16 j]); the caller then redoes the group with the full scan.
__global__ void __launch_bounds__(128) bucket_entries_kernel(const uint2* __restrict__ buckets,
                                                             const int32_t* __restrict__ bucket_count, int64_t n_buckets,
                                                             int64_t KB, int L, int margin, int64_t N,
                                                             const float* __restrict__ zhat, XEntry* __restrict__ entries,
                                                             int64_t sub_cap, unsigned long long* __restrict__ entry_count,
                                                             int32_t* __restrict__ row_nnz, double* __restrict__ rowstats,
                                                             int32_t* __restrict__ ovf_bkts = nullptr,
                                                             uint32_t* __restrict__ ovf_count = nullptr,
                                                             uint32_t ovf_bkt_cap = 0) {
  __shared__ uint32_t s_key[4][2][kBucketCap];
  __shared__ float s_val[4][2][kBucketCap];
  const unsigned full = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t bkt = (int64_t)blockIdx.x * 4 + warp;
  if (bkt >= n_buckets) return;
  const int n = bucket_count[bkt];
  if (n == 0) return;
  if (n > kBucketCap) {
    if (lane == 0) {
      // listed for bucket_overflow_kernel, or (no list, list full) word 1 of the counter block: "a bucket overflowed"
      const uint32_t i = ovf_bkts ? atomicAdd(ovf_count + 1, 1u) : 0xffffffffu;
      if (i < ovf_bkt_cap) ovf_bkts[i] = (int32_t)bkt;
      else atomicAdd(entry_count + 1, 1ull);
    }
    return;
  }
  const int t = (int)(bkt / KB), kb = (int)(bkt % KB);
  const uint2* rec = buckets + bkt * kBucketCap;
  for (int i = lane; i < n; i += 32) {
    const uint2 r = rec[i];
    s_key[warp][0][i] = r.x;                                // sequence (0..31) << 16 | position
    s_val[warp][0][i] = __uint_as_float(r.y);
  }
  __syncwarp();
  // rank sort (n <= 64, the keys of a task's bucket are distinct)
  for (int i = lane; i < n; i += 32) {
    const uint32_t k = s_key[warp][0][i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += s_key[warp][0][j] < k;
    s_key[warp][1][rank] = k;
    s_val[warp][1][rank] = s_val[warp][0][i];
  }
  __syncwarp();
  const uint32_t* key = s_key[warp][1];
  const float* val = s_val[warp][1];
  const int mg = margin, k = 2 * margin + 1;
  const float kf = (float)k;
  // rows owned by match i: window contains it and no earlier match of the same sequence
  int lo = 0, hi = -1, seq = 0;
  int cnt = 0;
  const int i = lane, i2 = lane + 32;
  auto owned = [&](int idx, int& r_lo, int& r_hi, int& sq) {
    r_lo = 0; r_hi = -1; sq = 0;
    if (idx >= n) return;
    sq = (int)(key[idx] >> 16);
    const int pos = (int)(key[idx] & 0xffffu);
    r_lo = max(pos - mg, mg);
    r_hi = min(pos + mg, L - 1 - mg);
    if (idx > 0 && (int)(key[idx - 1] >> 16) == sq) r_lo = max(r_lo, (int)(key[idx - 1] & 0xffffu) + mg + 1);
  };
  int lo2, hi2, seq2;
  owned(i, lo, hi, seq);
  owned(i2, lo2, hi2, seq2);
  cnt = max(hi - lo + 1, 0) + max(hi2 - lo2 + 1, 0);
  // entries == nullptr (match-driven profile, mepp_profile_matches): only the row sums are wanted
  const bool emit_entries = entries != nullptr;
  const int list = kb & (MEPP_ENTRY_LISTS - 1);
  unsigned long long slot = 0ull;
  if (emit_entries) {
    // one slot reservation per warp in the k-block's sub-list
    int pre = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(full, pre, o);
      if (lane >= o) pre += v;
    }
    const int total = __shfl_sync(full, pre, 31);
    if (total == 0) return;
    if (lane == 0) slot = atomicAdd(entry_count + list * 16, (unsigned long long)total);
    slot = __shfl_sync(full, slot, 0) + (unsigned long long)(pre - cnt);
  }
  uint4* dst = reinterpret_cast<uint4*>(entries + (int64_t)list * sub_cap);
  auto emit = [&](int idx, int r_lo, int r_hi, int sq) {
    const int64_t ns = (int64_t)kb * 32 + sq;
    const double z = (zhat && ns < N) ? (double)__ldg(zhat + ns) : 0.0;
    for (int p = r_lo; p <= r_hi; ++p) {
      float acc = 0.0f;
      for (int j = idx; j < n && (int)(key[j] >> 16) == sq && (int)(key[j] & 0xffffu) <= p + mg; ++j)
        acc = __fadd_rn(acc, val[j]);
      const float xb = k > 1 ? __fdiv_rn(acc, kf) : acc;
      const int r = t * L + p;
      if (emit_entries) {
        if ((int64_t)slot < sub_cap) dst[slot] = make_uint4((uint32_t)r, (uint32_t)ns, __float_as_uint(xb), 0u);
        ++slot;
        atomicAdd(row_nnz + r, 1);
      }
      const double v = (double)xb;
      double* rs = rowstats + (int64_t)r * 3;
      atomicAdd(rs + 0, v); atomicAdd(rs + 1, v * v); atomicAdd(rs + 2, v * z);
    }
  };
  emit(i, lo, hi, seq);
  emit(i2, lo2, hi2, seq2);
}
}  // namespace mepp

namespace mepp {
// The buckets that overflowed (a run of matches inside 32 sequences: low-complexity repeats, a strongly enriched motif
// in neighbouring sequences): one CTA per listed bucket collects its matches from the bucket (the first kBucketCap) and
// from the overflow list, sorts them by (sequence, position) in shared memory and emits exactly what
// bucket_entries_kernel emits.  More than kOvfMax matches in one bucket, or an overflow list that itself overflowed,
// raise the group-wide flag (entry_count[1]) and the caller redoes the group with the full-form scan.
__global__ void __launch_bounds__(256) bucket_overflow_kernel(const uint2* __restrict__ buckets,
                                                              const int32_t* __restrict__ bucket_count, int64_t KB, int L,
                                                              int margin, int64_t N, const float* __restrict__ zhat,
                                                              XEntry* __restrict__ entries, int64_t sub_cap,
                                                              unsigned long long* __restrict__ entry_count,
                                                              int32_t* __restrict__ row_nnz, double* __restrict__ rowstats,
                                                              const uint4* __restrict__ ovf,
                                                              const uint32_t* __restrict__ ovf_count, uint32_t ovf_cap,
                                                              const int32_t* __restrict__ ovf_bkts, uint32_t ovf_bkt_cap) {
  __shared__ uint32_t s_key[kOvfMax];
  __shared__ float s_val[kOvfMax];
  __shared__ int s_n;
  const uint32_t n_rec = ovf_count[0];
  uint32_t n_b = ovf_count[1];
  if (n_b == 0) return;
  if (n_rec > ovf_cap) {                                     // records were dropped: nothing here can be trusted
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(entry_count + 1, 1ull);
    return;
  }
  if (n_b > ovf_bkt_cap) n_b = ovf_bkt_cap;                  // (the unlisted ones raised the flag themselves)
  const int tid = threadIdx.x;
  const int mg = margin, k = 2 * margin + 1;
  const float kf = (float)k;
  for (uint32_t b = blockIdx.x; b < n_b; b += gridDim.x) {
    const int64_t bkt = ovf_bkts[b];
    const int t = (int)(bkt / KB), kb = (int)(bkt % KB);
    __syncthreads();
    if (tid == 0) s_n = kBucketCap;
    for (int i = tid; i < kBucketCap; i += blockDim.x) {
      const uint2 r = buckets[bkt * kBucketCap + i];
      s_key[i] = r.x; s_val[i] = __uint_as_float(r.y);
    }
    __syncthreads();
    for (uint32_t i = tid; i < n_rec; i += blockDim.x) {
      const uint4 r = __ldg(ovf + i);
      if ((int64_t)r.x == bkt) {
        const int j = atomicAdd(&s_n, 1);
        if (j < kOvfMax) { s_key[j] = r.y; s_val[j] = __uint_as_float(r.z); }
      }
    }
    __syncthreads();
    const int n = s_n;
    if (n > kOvfMax || n != bucket_count[bkt]) {             // too long a run for this kernel (or an inconsistent list)
      if (tid == 0) atomicAdd(entry_count + 1, 1ull);
      continue;
    }
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int i = n + tid; i < n2; i += blockDim.x) { s_key[i] = 0xffffffffu; s_val[i] = 0.0f; }
    __syncthreads();
    for (int kk = 2; kk <= n2; kk <<= 1)                     // bitonic sort of (key, value) pairs, ascending keys
      for (int jj = kk >> 1; jj > 0; jj >>= 1) {
        for (int i = tid; i < n2; i += blockDim.x) {
          const int p = i ^ jj;
          if (p > i) {
            const bool up = (i & kk) == 0;
            const uint32_t a0 = s_key[i], a1 = s_key[p];
            if ((a0 > a1) == up) {
              s_key[i] = a1; s_key[p] = a0;
              const float v = s_val[i]; s_val[i] = s_val[p]; s_val[p] = v;
            }
          }
        }
        __syncthreads();
      }
    const int list = kb & (MEPP_ENTRY_LISTS - 1);
    uint4* dst = entries ? reinterpret_cast<uint4*>(entries + (int64_t)list * sub_cap) : nullptr;
    for (int idx = tid; idx < n; idx += blockDim.x) {
      // rows owned by match idx: its window, minus the rows an earlier match of the same sequence owns
      const int sq = (int)(s_key[idx] >> 16), pos = (int)(s_key[idx] & 0xffffu);
      int r_lo = max(pos - mg, mg);
      const int r_hi = min(pos + mg, L - 1 - mg);
      if (idx > 0 && (int)(s_key[idx - 1] >> 16) == sq) r_lo = max(r_lo, (int)(s_key[idx - 1] & 0xffffu) + mg + 1);
      const int cnt = r_hi - r_lo + 1;
      if (cnt <= 0) continue;
      unsigned long long slot = dst ? atomicAdd(entry_count + list * 16, (unsigned long long)cnt) : 0ull;
      const int64_t ns = (int64_t)kb * 32 + sq;
      const double z = (zhat && ns < N) ? (double)__ldg(zhat + ns) : 0.0;
      for (int p = r_lo; p <= r_hi; ++p) {
        float acc = 0.0f;
        for (int j = idx; j < n && (int)(s_key[j] >> 16) == sq && (int)(s_key[j] & 0xffffu) <= p + mg; ++j)
          acc = __fadd_rn(acc, s_val[j]);
        const float xb = k > 1 ? __fdiv_rn(acc, kf) : acc;
        const int r = t * L + p;
        if (dst) {
          if ((int64_t)slot < sub_cap) dst[slot] = make_uint4((uint32_t)r, (uint32_t)ns, __float_as_uint(xb), 0u);
          ++slot;
          atomicAdd(row_nnz + r, 1);
        }
        const double v = (double)xb;
        double* rs = rowstats + (int64_t)r * 3;
        atomicAdd(rs + 0, v); atomicAdd(rs + 1, v * v); atomicAdd(rs + 2, v * z);
      }
    }
  }
}
}  // namespace mepp

namespace mepp {
static int scan_impl(const uint32_t* sym_T_dev, const float* zhat_dev,
                     int64_t N, int L, int T, int margin, const float* lut_dev, const float* bias_dev,
                     const int32_t* mlen_dev, const int32_t* nfilt_dev, const uint32_t* qtab_dev,
                     const uint32_t* qthr_dev, const int32_t* units_dev, int n_units, int max_motif_len,
                     void* x_planes_dev, double* rowstats_dev, int32_t* count_dev,
                     int32_t* density_dev, void* hits_dev, int64_t hits_cap, unsigned long long* hit_count_dev,
                     void* entries_dev, int64_t entry_cap, unsigned long long* entry_count_dev,
                     int32_t* row_nnz_dev, const uint32_t* planes_T_dev, const uint32_t* bp_dev, int32_t* bp_scratch_dev,
                     void* buckets_dev, int32_t* bucket_count_dev, const float* xscale_dev, void* stream) {
  MEPP_REQUIRE(sym_T_dev && lut_dev && bias_dev && mlen_dev && nfilt_dev && qtab_dev && qthr_dev &&
               rowstats_dev && count_dev && density_dev, "scan: null pointer");
  MEPP_REQUIRE(x_planes_dev || entries_dev, "scan: neither the tiled operand nor the entry list was requested");
  MEPP_REQUIRE(entries_dev == nullptr || (entry_count_dev && row_nnz_dev && entry_cap >= MEPP_ENTRY_LISTS &&
                                          entry_cap % MEPP_ENTRY_LISTS == 0),
               "scan: the entry list needs its counters and a capacity that is a multiple of MEPP_ENTRY_LISTS");
  MEPP_REQUIRE(N > 0 && L > 0 && T > 0, "scan: empty input");
  MEPP_REQUIRE(units_dev == nullptr || (n_units > 0 && n_units <= T), "scan: bad unit list");
  MEPP_REQUIRE(max_motif_len >= 0 && max_motif_len <= MEPP_MAX_MOTIF_LEN, "scan: max_motif_len outside [0, 32]");
  MEPP_REQUIRE((int64_t)T * L < (1LL << 31), "scan: T * L must be below 2^31");
  MEPP_REQUIRE(margin >= 0 && margin <= MEPP_MAX_MARGIN, "scan: margin must be in [0, 16]");
  MEPP_REQUIRE(hits_dev == nullptr || hit_count_dev != nullptr, "scan: hits need a hit counter");
  const bool bitpar = planes_T_dev != nullptr && L <= 2016;   // (survivor queue entries hold a 6-bit word step)
  MEPP_REQUIRE(!bitpar || (bp_dev && bp_scratch_dev), "scan_bitpar: the bit planes need the task records and the scratch");
  MEPP_REQUIRE(mepp_device_ok(), "scan: no sm_100 CUDA device (no CPU fallback)");
  cudaStream_t st = as_stream(stream);
  ScanArgs a;
  a.sym_T = sym_T_dev; a.zhat = zhat_dev;
  a.N = N; a.n_pad = mepp_pad32(N); a.KB = a.n_pad / kBlockK;
  a.L = L; a.T = T; a.margin = margin;
  a.W3 = (int)mepp_sym_words(L);
  {
    const int np = max_motif_len > 0 ? (max_motif_len + 1) / 2 : MEPP_MAX_MOTIF_LEN / 2;
    a.qslots = np <= 3 ? 3 : np <= 5 ? 5 : np <= 6 ? 6 : np <= 8 ? 8 : np <= 11 ? 11 : 16;   // whole ladder buckets
  }
  a.lut = lut_dev; a.bias = bias_dev; a.mlen = mlen_dev; a.nfilt = nfilt_dev;
  a.qtab = qtab_dev; a.qthr = qthr_dev;
  a.units = units_dev; a.U = units_dev ? n_units : T;
  a.x_planes = reinterpret_cast<uint8_t*>(x_planes_dev); a.rowstats = rowstats_dev;
  a.count = count_dev; a.density = density_dev;
  a.hits = reinterpret_cast<HitRec*>(hits_dev); a.hits_cap = hits_cap; a.hit_count = hit_count_dev;
  a.entries = reinterpret_cast<XEntry*>(entries_dev); a.sub_cap = entry_cap / MEPP_ENTRY_LISTS;
  a.entry_count = entry_count_dev; a.row_nnz = row_nnz_dev;
  a.planes_T = planes_T_dev; a.bp = bp_dev; a.unit_count = nullptr;
  const bool lean = buckets_dev != nullptr;
  MEPP_REQUIRE(!lean || (bucket_count_dev && entries_dev && !x_planes_dev && !bitpar && L <= 65535),
               "scan_lean: needs the bucket counters and the entry list, no tiled operand, no bit planes, L <= 65535");
  a.lean = lean ? 1 : 0; a.buckets = reinterpret_cast<uint2*>(buckets_dev); a.bucket_count = bucket_count_dev;
  a.ovf = nullptr; a.ovf_count = nullptr; a.ovf_cap = 0; a.ovf_bkts = nullptr; a.ovf_bkt_cap = 0;
  a.xscale = xscale_dev;
  const int nbuf = 32 + 2 * margin;
  auto smem_bytes = [&](bool bp) {
    return (lean ? 0 : sizeof(float) * kScanWarps * ((units_dev || bitpar) ? 2 : 1) * nbuf * kRowStride) +
           sizeof(uint32_t) * kScanWarps * scan_table_words(a.qslots, bp) + sizeof(uint16_t) * kScanWarps * kQueue +
           sizeof(uint32_t) * kScanWarps * scan_ring_words(L, bp) + (bp ? sizeof(float2) * kScanWarps * a.qslots * 8 : 0) +
           256 /* table alignment */;
  };
  const size_t smem = smem_bytes(false);
  MEPP_CUDA_CHECK(cudaFuncSetAttribute(scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    static int carveout = [] { const char* e = getenv("MEPP_SCAN_CARVEOUT"); return e ? atoi(e) : -2; }();
    if (carveout >= -1)
      MEPP_CUDA_CHECK(cudaFuncSetAttribute(scan_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
  }
  MEPP_CUDA_CHECK(cudaMemsetAsync(rowstats_dev, 0, sizeof(double) * 3 * (size_t)T * L, st));
  MEPP_CUDA_CHECK(cudaMemsetAsync(count_dev, 0, sizeof(int32_t) * (size_t)T * L, st));
  MEPP_CUDA_CHECK(cudaMemsetAsync(density_dev, 0, sizeof(int32_t) * (size_t)T * a.n_pad, st));
  if (hit_count_dev) MEPP_CUDA_CHECK(cudaMemsetAsync(hit_count_dev, 0, sizeof(unsigned long long), st));
  if (entries_dev) {
    MEPP_CUDA_CHECK(cudaMemsetAsync(entry_count_dev, 0, sizeof(unsigned long long) * 16 * MEPP_ENTRY_LISTS, st));
    MEPP_CUDA_CHECK(cudaMemsetAsync(row_nnz_dev, 0, sizeof(int32_t) * ((size_t)T * L + 1), st));
  }
  if (x_planes_dev) {
    // reset the non-zero bit array behind the tiles
    const int64_t n_tiles_m = ((int64_t)T * L + kBlockM - 1) / kBlockM;
    MEPP_CUDA_CHECK(cudaMemsetAsync(a.x_planes + a_flags_offset(n_tiles_m, a.KB), 0,
                                    (size_t)(n_tiles_m * flag_words(a.KB) * 4), st));
  }
  dim3 grid((unsigned)a.KB, (unsigned)((a.U + kScanWarps - 1) / kScanWarps));
  MEPP_REQUIRE(grid.y <= 65535, "scan: too many tasks in one call");
  if (lean) MEPP_CUDA_CHECK(cudaMemsetAsync(bucket_count_dev, 0, sizeof(int32_t) * (size_t)T * a.KB, st));
  if (!bitpar) {
    scan_kernel<false><<<grid, kScanWarps * 32, smem, st>>>(a);
    MEPP_CUDA_CHECK(cudaGetLastError());
    if (lean) {
      const int64_t n_buckets = (int64_t)T * a.KB;
      bucket_entries_kernel<<<(unsigned)((n_buckets + 3) / 4), 128, 0, st>>>(
          a.buckets, bucket_count_dev, n_buckets, a.KB, L, margin, N, zhat_dev, a.entries, a.sub_cap, entry_count_dev,
          row_nnz_dev, rowstats_dev);
      MEPP_CUDA_CHECK(cudaGetLastError());
    }
    return 0;
  }
  // units split by class on the device (the class depends on the thresholds, which may have been computed there);
  // both kernels are launched over the whole grid and leave at once past their list's end
  const int U = a.U;
  int32_t* lists = bp_scratch_dev;            // [2][U][2]
  int32_t* counts = bp_scratch_dev + 4 * U;   // [2]
  bp_partition_kernel<<<1, 32, 0, st>>>(units_dev, U, bp_dev, lists, counts);
  MEPP_CUDA_CHECK(cudaGetLastError());
  const size_t smem_bp = smem_bytes(true);
  MEPP_CUDA_CHECK(cudaFuncSetAttribute(scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bp));
  ScanArgs b = a;
  b.units = lists + 2 * U; b.unit_count = counts + 1;
  scan_kernel<true><<<grid, kScanWarps * 32, smem_bp, st>>>(b);
  MEPP_CUDA_CHECK(cudaGetLastError());
  a.units = lists; a.unit_count = counts;
  scan_kernel<false><<<grid, kScanWarps * 32, smem, st>>>(a);
  MEPP_CUDA_CHECK(cudaGetLastError());
  if (lean) {
    const int64_t n_buckets = (int64_t)T * a.KB;
    bucket_entries_kernel<<<(unsigned)((n_buckets + 3) / 4), 128, 0, st>>>(
        a.buckets, bucket_count_dev, n_buckets, a.KB, L, margin, N, zhat_dev, a.entries, a.sub_cap, entry_count_dev,
        row_nnz_dev, rowstats_dev);
    MEPP_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}
}  // namespace mepp

extern "C" int mepp_scan(const uint32_t* sym_T_dev, const float* zhat_dev,
                         int64_t N, int L, int T, int margin, const float* lut_dev, const float* bias_dev,
                         const int32_t* mlen_dev, const int32_t* nfilt_dev, const uint32_t* qtab_dev,
                         const uint32_t* qthr_dev, const int32_t* units_dev, int n_units, int max_motif_len,
                         void* x_planes_dev, double* rowstats_dev, int32_t* count_dev,
                         int32_t* density_dev, void* hits_dev, int64_t hits_cap, unsigned long long* hit_count_dev,
                         void* entries_dev, int64_t entry_cap, unsigned long long* entry_count_dev,
                         int32_t* row_nnz_dev, const float* xscale_dev, void* stream) {
  return mepp::scan_impl(sym_T_dev, zhat_dev, N, L, T, margin, lut_dev, bias_dev, mlen_dev, nfilt_dev, qtab_dev, qthr_dev,
                         units_dev, n_units, max_motif_len, x_planes_dev, rowstats_dev, count_dev, density_dev, hits_dev,
                         hits_cap, hit_count_dev, entries_dev, entry_cap, entry_count_dev, row_nnz_dev, nullptr, nullptr,
                         nullptr, nullptr, nullptr, xscale_dev, stream);
}

extern "C" int mepp_scan_bitpar(const uint32_t* sym_T_dev, const float* zhat_dev,
                                int64_t N, int L, int T, int margin, const float* lut_dev, const float* bias_dev,
                                const int32_t* mlen_dev, const int32_t* nfilt_dev, const uint32_t* qtab_dev,
                                const uint32_t* qthr_dev, const int32_t* units_dev, int n_units, int max_motif_len,
                                void* x_planes_dev, double* rowstats_dev, int32_t* count_dev,
                                int32_t* density_dev, void* hits_dev, int64_t hits_cap, unsigned long long* hit_count_dev,
                                void* entries_dev, int64_t entry_cap, unsigned long long* entry_count_dev,
                                int32_t* row_nnz_dev, const uint32_t* planes_T_dev, const uint32_t* bp_dev,
                                int32_t* bp_scratch_dev, const float* xscale_dev, void* stream) {
  MEPP_REQUIRE(planes_T_dev && bp_dev && bp_scratch_dev, "scan_bitpar: null pointer");
  return mepp::scan_impl(sym_T_dev, zhat_dev, N, L, T, margin, lut_dev, bias_dev, mlen_dev, nfilt_dev, qtab_dev, qthr_dev,
                         units_dev, n_units, max_motif_len, x_planes_dev, rowstats_dev, count_dev, density_dev, hits_dev,
                         hits_cap, hit_count_dev, entries_dev, entry_cap, entry_count_dev, row_nnz_dev, planes_T_dev,
                         bp_dev, bp_scratch_dev, nullptr, nullptr, xscale_dev, stream);
}

extern "C" int64_t mepp_scan_lean_bucket_bytes(int T, int64_t N) {
  return (int64_t)T * (mepp_pad32(N) / mepp::kBlockK) * mepp::kBucketCap * 8;
}

extern "C" int mepp_scan_lean(const uint32_t* sym_T_dev, const float* zhat_dev,
                              int64_t N, int L, int T, int margin, const float* lut_dev, const float* bias_dev,
                              const int32_t* mlen_dev, const int32_t* nfilt_dev, const uint32_t* qtab_dev,
                              const uint32_t* qthr_dev, const int32_t* units_dev, int n_units, int max_motif_len,
                              double* rowstats_dev, int32_t* count_dev,
                              int32_t* density_dev, void* hits_dev, int64_t hits_cap, unsigned long long* hit_count_dev,
                              void* entries_dev, int64_t entry_cap, unsigned long long* entry_count_dev,
                              int32_t* row_nnz_dev, void* buckets_dev, int32_t* bucket_count_dev, void* stream) {
  MEPP_REQUIRE(buckets_dev && bucket_count_dev && entries_dev, "scan_lean: null pointer");
  return mepp::scan_impl(sym_T_dev, zhat_dev, N, L, T, margin, lut_dev, bias_dev, mlen_dev, nfilt_dev, qtab_dev, qthr_dev,
                         units_dev, n_units, max_motif_len, nullptr, rowstats_dev, count_dev, density_dev, hits_dev,
                         hits_cap, hit_count_dev, entries_dev, entry_cap, entry_count_dev, row_nnz_dev, nullptr, nullptr,
                         nullptr, buckets_dev, bucket_count_dev, nullptr, stream);
}

extern "C" int mepp_scan_bitpar_tables(const float* lut_dev, const float* bias_dev, const int32_t* mlen_dev,
                                       const int32_t* nfilt_dev, int T, double cost_ratio, uint32_t* bp_dev,
                                       void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(lut_dev && bias_dev && mlen_dev && nfilt_dev && bp_dev && T > 0, "scan_bitpar_tables: bad arguments");
  MEPP_REQUIRE(mepp_device_ok(), "scan_bitpar_tables: no sm_100 CUDA device (no CPU fallback)");
  bp_tables_kernel<<<(unsigned)((T + 3) / 4), 128, 0, as_stream(stream)>>>(lut_dev, bias_dev, mlen_dev, nfilt_dev, T,
                                                                           cost_ratio, bp_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}


// ---- tensor-core first stage (tcscan.cu) + exact evaluation of its candidates ----------------------------------
namespace mepp {
// One thread per candidate {row, unit << 2 | phase}: the window starts at flat base 4 row + phase of the one-hot
// stream, i.e. at base i of sequence seq; windows that run over the end of their sequence (or of the data) and the
// padding unit slots of an N-tile are dropped here.  Survivors are scored in the canonical float32 order
// (exact_x0's order: taps ascending, one add per base, four quarter-weight adds for N, bias last) and published
// exactly like the lean scan's matches (buckets, counts, densities, match list).
__global__ void __launch_bounds__(128) tc_exact_kernel(const ScanArgs a, const uint2* __restrict__ cand,
                                                       const uint32_t* __restrict__ cand_count, uint32_t list_cap,
                                                       int unit_base, int n_units_launch) {
  const int64_t list = blockIdx.x;
  uint32_t cnt = cand_count[list];
  if (cnt > list_cap) {
    if (threadIdx.x == 0) atomicAdd(a.entry_count + 1, 1ull);     // overflow: the caller redoes the group (like a bucket)
    cnt = list_cap;
  }
  const uint2* my = cand + list * (int64_t)list_cap;
  for (uint32_t idx = threadIdx.x; idx < cnt; idx += blockDim.x) {
    const uint2 rec = my[idx];
    const int ul = (int)(rec.y >> 2), phi = (int)(rec.y & 3u);
    if (ul >= n_units_launch) continue;
    const int u = unit_base + ul;
    const int t = a.units ? a.units[2 * u] : u;
    const int t2 = a.units ? a.units[2 * u + 1] : -1;
    const int m = a.mlen[t];
    const int64_t b = 4 * (int64_t)rec.x + phi;
    const int64_t seq = b / a.L;
    const int i = (int)(b - seq * a.L);
    if (seq >= a.N || i + m > a.L) continue;
    // exact scores of both filters
    const uint32_t* gsym = a.sym_T + seq;
    const int bit0 = 3 * i, w0 = bit0 >> 5, sh = bit0 & 31;
    uint32_t x[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) x[k] = __ldg(gsym + (int64_t)(w0 + k) * a.n_pad);
    uint32_t r0 = __funnelshift_r(x[0], x[1], sh), r1 = __funnelshift_r(x[1], x[2], sh),
             r2 = __funnelshift_r(x[2], x[3], sh), r3 = __funnelshift_r(x[3], x[4], sh);
    const float2* lut = reinterpret_cast<const float2*>(a.lut) + (size_t)t * (MEPP_MAX_MOTIF_LEN * 4);
    float a0 = 0.0f, a1 = 0.0f;
    // The weights of a task (1 KB) rarely sit in the L1 (41 % hit rate over the 162 tasks of a step), and a load per tap
    // behind the previous tap's addition made the evaluation a chain of m L2 round trips (ncu: issue slots 8 % busy, 115
    // cycles of long-scoreboard stall per instruction): the loads of eight taps are issued together, the additions stay
    // in the canonical order.
#pragma unroll 1
    for (int j0 = 0; j0 < m; j0 += 8) {
      float2 w[8];
      uint32_t sy[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        sy[u] = r0 & 7u;
        r0 = __funnelshift_r(r0, r1, 3); r1 = __funnelshift_r(r1, r2, 3); r2 = __funnelshift_r(r2, r3, 3); r3 >>= 3;
        w[u] = (j0 + u < m && sy[u] < 4u) ? __ldg(lut + (j0 + u) * 4 + sy[u]) : make_float2(0.0f, 0.0f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (j0 + u >= m) break;
        if (sy[u] >= 4u) {
          const float2* l4 = lut + (j0 + u) * 4;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const float2 wn = __ldg(l4 + cc);
            a0 = __fadd_rn(a0, __fmul_rn(0.25f, wn.x));
            a1 = __fadd_rn(a1, __fmul_rn(0.25f, wn.y));
          }
        } else {
          a0 = __fadd_rn(a0, w[u].x);
          a1 = __fadd_rn(a1, w[u].y);
        }
      }
    }
    const float bias = a.bias[t];
    float sc = __fadd_rn(a0, bias);
    const float x_first = fmaxf(sc, 0.0f);
    if (a.nfilt[t] == 2) sc = fmaxf(sc, __fadd_rn(a1, bias));
    const float x_task = fmaxf(sc, 0.0f);
    const int p = i + (m - 1) / 2;                 // motif_layers.py:99-101
    const int kb = (int)(seq >> 5), s = (int)(seq & 31);
    auto publish = [&](int tt, float x0) {
      const int64_t bkt = (int64_t)tt * a.KB + kb;
      const int slot = atomicAdd(&a.bucket_count[bkt], 1);
      const uint32_t key = ((uint32_t)s << 16) | (uint32_t)p;
      if (slot < kBucketCap) {
        a.buckets[bkt * kBucketCap + slot] = make_uint2(key, __float_as_uint(x0));
      } else if (a.ovf) {
        const uint32_t o = atomicAdd(a.ovf_count, 1u);
        if (o < a.ovf_cap) a.ovf[o] = make_uint4((uint32_t)bkt, key, __float_as_uint(x0), 0u);
      }
      atomicAdd(&a.count[(int64_t)tt * a.L + p], 1);
      atomicAdd(&a.density[(int64_t)tt * a.n_pad + seq], 1);
      if (a.hit_count) {
        // one reservation per group of lanes that publish together: every match of a launch goes through this ONE counter
        const unsigned act = __activemask();
        const int leader = __ffs((int)act) - 1, rank = __popc(act & ((1u << (threadIdx.x & 31)) - 1u));
        unsigned long long base = 0ull;
        if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(a.hit_count, (unsigned long long)__popc(act));
        const unsigned long long hs = __shfl_sync(act, base, leader) + (unsigned long long)rank;
        if (a.hits && (int64_t)hs < a.hits_cap) {
          HitRec h; h.task = tt; h.seq = (int32_t)seq; h.pos = p; h.x0 = x0;
          a.hits[hs] = h;
        }
      }
    };
    if (x_task > 0.0f) publish(t, x_task);
    if (t2 >= 0 && x_first > 0.0f) publish(t2, x_first);
  }
}
}  // namespace mepp

// Per-task power-of-two scale of the tiled operand: the largest s = 2^k with s * (largest possible score - threshold)
// <= 16384, never below MEPP_X_SCALE.  The fp16 hi + lo pair carries ~22 bits only while lo is a normal fp16 number, i.e.
// for s * xbar >~ 0.13: with the one global scale of 256, rows whose non-zeros are all below 5e-4 lost relative precision
// (3e-5 in r, found by tools/gpu_fuzz.py); a task whose scores can only be small now gets a correspondingly larger scale.
namespace mepp {
__global__ void x_scales_kernel(const float* __restrict__ lut, const float* __restrict__ bias,
                                const int32_t* __restrict__ mlen, const int32_t* __restrict__ nfilt, int T,
                                float* __restrict__ xscale) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float* l = lut + (size_t)t * (MEPP_MAX_MOTIF_LEN * 8);
  const int m = mlen[t], nf = nfilt[t];
  double best = 0.0;
  for (int f = 0; f < nf; ++f) {
    double sm = 0.0;
    for (int j = 0; j < m; ++j) {
      double mx = -1e300;
      for (int c = 0; c < 4; ++c) mx = fmax(mx, (double)l[(j * 4 + c) * 2 + f]);
      sm += mx;
    }
    best = fmax(best, sm + (double)bias[t]);             // bias = float32(-threshold)
  }
  float s = (float)MEPP_X_SCALE;
  if (best > 0.0) {
    const double want = 16384.0 / (best * 1.001 + 1e-6);
    int e = 0;
    frexp(want, &e);                                     // want = f * 2^e, 0.5 <= f < 1: 2^(e-1) <= want
    double p2 = ldexp(1.0, e - 1);
    if (p2 > 16777216.0) p2 = 16777216.0;
    if (p2 > (double)MEPP_X_SCALE) s = (float)p2;
  }
  xscale[t] = s;
}
}  // namespace mepp

extern "C" int mepp_x_scales(const float* lut_dev, const float* bias_dev, const int32_t* mlen_dev, const int32_t* nfilt_dev,
                             int T, float* xscale_dev, void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(lut_dev && bias_dev && mlen_dev && nfilt_dev && xscale_dev && T > 0, "x_scales: bad arguments");
  MEPP_REQUIRE(mepp_device_ok(), "x_scales: no sm_100 CUDA device (no CPU fallback)");
  x_scales_kernel<<<(unsigned)((T + 127) / 128), 128, 0, as_stream(stream)>>>(lut_dev, bias_dev, mlen_dev, nfilt_dev, T,
                                                                            xscale_dev);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// Optional per-launch timing of tc_scan_kernel (bench.py's roofline block): CUDA events on the launching stream around
// every launch while enabled, and the int8 multiply-adds the launches issued.
namespace mepp {
struct TcTiming {
  bool on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  double macs = 0.0;          // issued int8 multiply-adds (128 x bn x 32 per K-step and M-tile, padding included)
  double stream_bytes = 0.0;  // bytes of the one-hot stream the launches read
};
static TcTiming g_tc_timing;
}  // namespace mepp

extern "C" int mepp_scan_tc_timing(int enable, double* ms_out, double* macs_out, double* stream_bytes_out,
                                   int* launches_out) {
  using namespace mepp;
  TcTiming& t = g_tc_timing;
  double ms = 0.0;
  for (auto& e : t.ev) {
    MEPP_CUDA_CHECK(cudaEventSynchronize(e.second));
    float x = 0.0f;
    MEPP_CUDA_CHECK(cudaEventElapsedTime(&x, e.first, e.second));
    ms += x;
    cudaEventDestroy(e.first); cudaEventDestroy(e.second);
  }
  if (ms_out) *ms_out = ms;
  if (macs_out) *macs_out = t.macs;
  if (stream_bytes_out) *stream_bytes_out = t.stream_bytes;
  if (launches_out) *launches_out = (int)t.ev.size();
  t.ev.clear(); t.macs = 0.0; t.stream_bytes = 0.0;
  t.on = enable != 0;
  return 0;
}

namespace mepp {
// bucket overflow (ScanArgs::ovf): counters, the list of overflowed buckets, the records; behind the candidate lists
constexpr uint32_t kTcOvfRecs = 1u << 20, kTcOvfBkts = 1u << 16;
constexpr int64_t kTcOvfBytes = 256 + (int64_t)kTcOvfBkts * 4 + (int64_t)kTcOvfRecs * 16;
}  // namespace mepp

extern "C" int64_t mepp_scan_tc_work_bytes(int64_t N, int L, int n_units) {
  using namespace mepp;
  // weight operands of every launch (<= 8 columns x 160 channels per unit, plus tile padding), candidate lists, counters
  const int64_t lists = 148 * 2 * kTcEpiWarps;                        // (up to 2x the SM count of a B200)
  const double expect = (double)N * (double)L * (double)(n_units < 256 ? n_units : 256) * 4e-3 / (148.0 * kTcEpiWarps);
  int64_t cap = (int64_t)expect + 2048;
  if (cap > (1 << 22)) cap = 1 << 22;
  return 1024 + (int64_t)(n_units + n_units / 8 + 64) * 8 * 160 + lists * 16 + lists * cap * 8 + 1024 + kTcOvfBytes;
}

extern "C" int mepp_scan_tc(const uint32_t* sym_T_dev, const uint32_t* onehot_dev, const float* zhat_dev,
                            int64_t N, int L, int T, int margin, const float* lut_dev, const float* bias_dev,
                            const int32_t* mlen_dev, const int32_t* nfilt_dev, const int32_t* units_dev,
                            const int32_t* units_host, int n_units, const int32_t* mlen_host,
                            double* rowstats_dev, int32_t* count_dev, int32_t* density_dev, void* hits_dev,
                            int64_t hits_cap, unsigned long long* hit_count_dev, void* entries_dev, int64_t entry_cap,
                            unsigned long long* entry_count_dev, int32_t* row_nnz_dev, void* buckets_dev,
                            int32_t* bucket_count_dev, void* work_dev, int64_t work_bytes, int32_t* dbg_acc_dev, int dbg_ld,
                            void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(sym_T_dev && onehot_dev && lut_dev && bias_dev && mlen_dev && nfilt_dev && mlen_host && rowstats_dev &&
               count_dev && density_dev && entry_count_dev && row_nnz_dev && buckets_dev &&
               bucket_count_dev && work_dev, "scan_tc: null pointer");
  MEPP_REQUIRE(N > 0 && L > 0 && T > 0 && L <= 65535, "scan_tc: bad sizes (L <= 65535)");
  MEPP_REQUIRE((units_dev == nullptr) == (units_host == nullptr), "scan_tc: the unit list is needed on both sides");
  MEPP_REQUIRE(units_dev == nullptr || (n_units > 0 && n_units <= T), "scan_tc: bad unit list");
  MEPP_REQUIRE((int64_t)T * L < (1LL << 31), "scan_tc: T * L must be below 2^31");
  MEPP_REQUIRE(margin >= 0 && margin <= MEPP_MAX_MARGIN, "scan_tc: margin must be in [0, 16]");
  MEPP_REQUIRE(entries_dev == nullptr || (entry_cap >= MEPP_ENTRY_LISTS && entry_cap % MEPP_ENTRY_LISTS == 0),
               "scan_tc: bad entry capacity");
  MEPP_REQUIRE(hits_dev == nullptr || hit_count_dev != nullptr, "scan_tc: hits need a hit counter");
  MEPP_REQUIRE(N * (int64_t)L / 4 < (1LL << 32), "scan_tc: N * L must be below 2^34");
  MEPP_REQUIRE(mepp_device_ok(), "scan_tc: no sm_100 CUDA device (no CPU fallback)");
  cudaStream_t st = as_stream(stream);
  const int U = units_dev ? n_units : T;
  MEPP_REQUIRE(work_bytes >= mepp_scan_tc_work_bytes(N, L, U), "scan_tc: work buffer too small");
  ScanArgs a;
  a.sym_T = sym_T_dev; a.zhat = zhat_dev;
  a.N = N; a.n_pad = mepp_pad32(N); a.KB = a.n_pad / kBlockK;
  a.L = L; a.T = T; a.margin = margin; a.W3 = (int)mepp_sym_words(L); a.qslots = 0;
  a.lut = lut_dev; a.bias = bias_dev; a.mlen = mlen_dev; a.nfilt = nfilt_dev; a.qtab = nullptr; a.qthr = nullptr;
  a.units = units_dev; a.U = U;
  a.x_planes = nullptr; a.rowstats = rowstats_dev; a.count = count_dev; a.density = density_dev;
  a.hits = reinterpret_cast<HitRec*>(hits_dev); a.hits_cap = hits_cap; a.hit_count = hit_count_dev;
  a.entries = reinterpret_cast<XEntry*>(entries_dev); a.sub_cap = entry_cap / MEPP_ENTRY_LISTS;
  a.entry_count = entry_count_dev; a.row_nnz = row_nnz_dev;
  a.planes_T = nullptr; a.bp = nullptr; a.unit_count = nullptr;
  a.lean = 1; a.buckets = reinterpret_cast<uint2*>(buckets_dev); a.bucket_count = bucket_count_dev;
  a.xscale = nullptr;
  {
    uint8_t* ob = reinterpret_cast<uint8_t*>(work_dev) + ((work_bytes - kTcOvfBytes) & ~(int64_t)255);
    a.ovf_count = reinterpret_cast<uint32_t*>(ob);
    a.ovf_bkts = reinterpret_cast<int32_t*>(ob + 256); a.ovf_bkt_cap = kTcOvfBkts;
    a.ovf = reinterpret_cast<uint4*>(ob + 256 + (size_t)kTcOvfBkts * 4); a.ovf_cap = kTcOvfRecs;
    MEPP_CUDA_CHECK(cudaMemsetAsync(a.ovf_count, 0, 8, st));
  }
  MEPP_CUDA_CHECK(cudaMemsetAsync(rowstats_dev, 0, sizeof(double) * 3 * (size_t)T * L, st));
  MEPP_CUDA_CHECK(cudaMemsetAsync(count_dev, 0, sizeof(int32_t) * (size_t)T * L, st));
  MEPP_CUDA_CHECK(cudaMemsetAsync(density_dev, 0, sizeof(int32_t) * (size_t)T * a.n_pad, st));
  if (hit_count_dev) MEPP_CUDA_CHECK(cudaMemsetAsync(hit_count_dev, 0, sizeof(unsigned long long), st));
  MEPP_CUDA_CHECK(cudaMemsetAsync(entry_count_dev, 0, sizeof(unsigned long long) * 16 * MEPP_ENTRY_LISTS, st));
  MEPP_CUDA_CHECK(cudaMemsetAsync(row_nnz_dev, 0, sizeof(int32_t) * ((size_t)T * L + 1), st));
  MEPP_CUDA_CHECK(cudaMemsetAsync(bucket_count_dev, 0, sizeof(int32_t) * (size_t)T * a.KB, st));
  // work buffer: [weights][counters][candidate lists]
  const int grid = tc_scan_grid();
  const int64_t lists = (int64_t)grid * kTcEpiWarps;
  uint8_t* wk = reinterpret_cast<uint8_t*>(work_dev);
  wk += (1024 - (reinterpret_cast<uintptr_t>(wk) & 1023)) & 1023;
  uint8_t* B_dev = wk;
  const int64_t b_region = ((int64_t)(U + U / 8 + 64) * 8 * 160 + 1023) / 1024 * 1024;
  uint32_t* counts_dev = reinterpret_cast<uint32_t*>(wk + b_region);
  uint2* cand_dev = reinterpret_cast<uint2*>(wk + b_region + ((lists * 4 + 1023) / 1024 * 1024));
  // (the last kTcOvfBytes of the work buffer hold the bucket-overflow lists)
  uint8_t* ovf_base = reinterpret_cast<uint8_t*>(work_dev) + ((work_bytes - kTcOvfBytes) & ~(int64_t)255);
  const int64_t cand_bytes = ovf_base - reinterpret_cast<uint8_t*>(cand_dev);
  int64_t cap64 = cand_bytes / (lists * 8);
  MEPP_REQUIRE(cap64 >= 256, "scan_tc: work buffer too small for the candidate lists");
  if (cap64 > (1 << 22)) cap64 = 1 << 22;
  TcScanArgs g;
  g.onehot = reinterpret_cast<const uint8_t*>(onehot_dev);
  g.n_mtiles = (N * (int64_t)L + kTcBasesPerTile - 1) / kTcBasesPerTile;
  g.cand = cand_dev; g.cand_count = counts_dev; g.list_cap = (uint32_t)cap64;
  g.dbg_acc = dbg_ld > 0 ? dbg_acc_dev : nullptr; g.dbg_ld = dbg_ld;
  g.dbg_clk = dbg_ld < 0 ? reinterpret_cast<unsigned long long*>(dbg_acc_dev) : nullptr;   // (dbg_ld < 0: cycle stamps)
  { static const int mode = [] { const char* e = getenv("MEPP_TC_MODE"); const char* p = getenv("MEPP_TC_PARK_NS");
                                 return (e ? atoi(e) & 255 : 0) | ((p ? atoi(p) : 0) << 8); }(); g.mode = mode; }
  // The units of a launch are dealt EVENLY over its N-tiles (e.g. 41 units -> 24 + 17, not 32 + 9): an N-tile comes back
  // to the same accumulator stage on the next stream tile and has to wait for its own epilogue (commit, TMEM reads,
  // barrier hand-over: ~450 cycles, cycle stamps of tools/gpu_tc_clk.py), which only the OTHER N-tile's products can
  // cover; and N stays >= ~160, where the MMA is bound by the tensor pipe rather than by its operand reads.
  // (MEPP_TC_TILES_PER_LAUNCH = 1: one N-tile per launch, measured slower -- the stream is scanned once per launch.)
  static const int tiles_per_launch = [] { const char* e = getenv("MEPP_TC_TILES_PER_LAUNCH"); int v = e ? atoi(e) : kTcMaxTiles;
                                           return v < 1 ? 1 : v > kTcMaxTiles ? kTcMaxTiles : v; }();
  const int n_ntiles = (U + tc_units_per_tile() - 1) / tc_units_per_tile();
  const int units_per_tile = ((U + n_ntiles - 1) / n_ntiles + 3) / 4 * 4;
  int u0 = 0;
  int64_t b_used = 0;
  while (u0 < U) {
    TcLaunch ln;
    ln.n_tiles = 0; ln.ks_max = 1; ln.b_bytes = 0;
    int u = u0;
    while (u < U && ln.n_tiles < tiles_per_launch) {
      const int nu = U - u < units_per_tile ? U - u : units_per_tile;
      int mmax = 1;
      for (int k = 0; k < nu; ++k) {
        const int t = units_host ? units_host[2 * (u + k)] : u + k;
        MEPP_REQUIRE(t >= 0 && t < T, "scan_tc: unit list refers to a task outside [0, T)");
        const int m = mlen_host[t];
        MEPP_REQUIRE(m >= 1 && m <= MEPP_MAX_MOTIF_LEN, "scan_tc: motif length outside [1, 32]");
        if (m > mmax) mmax = m;
      }
      TcTile tl;
      tl.ksteps = (uint16_t)((4 * (mmax + 3) + 31) / 32);
      tl.bn = (uint16_t)(32 * ((nu + 3) / 4));
      tl.unit0 = (uint16_t)(u - u0);
      tl.n_units = (uint16_t)nu;
      tl.b_off = ln.b_bytes;
      const uint32_t bytes = (uint32_t)tl.ksteps * 2u * tl.bn * 16u;
      if (ln.b_bytes + bytes > (uint32_t)kTcMaxBBytes) break;
      ln.b_bytes += bytes;
      if (tl.ksteps > ln.ks_max) ln.ks_max = tl.ksteps;
      ln.tiles[ln.n_tiles++] = tl;
      u += nu;
    }
    MEPP_REQUIRE(ln.n_tiles > 0, "scan_tc: a tile does not fit the weight operand");
    for (int k = ln.n_tiles; k < kTcMaxTiles; ++k) { ln.tiles[k].b_off = 0; ln.tiles[k].ksteps = 0; ln.tiles[k].bn = 0; ln.tiles[k].unit0 = 0; ln.tiles[k].n_units = 0; }
    MEPP_REQUIRE(b_used + ln.b_bytes <= b_region, "scan_tc: weight region overflow");
    uint8_t* B_launch = B_dev + b_used;
    b_used += ((int64_t)ln.b_bytes + 1023) / 1024 * 1024;
    if (tc_build_weights(ln, lut_dev, bias_dev, mlen_dev, nfilt_dev, units_dev, u0, U, B_launch, st)) return 1;
    g.B = B_launch;
    if (g_tc_timing.on) {
      cudaEvent_t e0, e1;
      MEPP_CUDA_CHECK(cudaEventCreate(&e0)); MEPP_CUDA_CHECK(cudaEventCreate(&e1));
      MEPP_CUDA_CHECK(cudaEventRecord(e0, st));
      if (tc_scan_launch(ln, g, st)) return 1;
      MEPP_CUDA_CHECK(cudaEventRecord(e1, st));
      g_tc_timing.ev.emplace_back(e0, e1);
      for (int k = 0; k < ln.n_tiles; ++k)
        g_tc_timing.macs += (double)g.n_mtiles * kTcRowsPerTile * ln.tiles[k].bn * 32.0 * ln.tiles[k].ksteps;
      g_tc_timing.stream_bytes += (double)g.n_mtiles * 2048.0;
    } else if (tc_scan_launch(ln, g, st)) return 1;
    int n_lists = grid;
    if ((int64_t)n_lists > g.n_mtiles) n_lists = (int)g.n_mtiles;
    tc_exact_kernel<<<(unsigned)(n_lists * kTcEpiWarps), 128, 0, st>>>(a, g.cand, g.cand_count, g.list_cap, u0, u - u0);
    MEPP_CUDA_CHECK(cudaGetLastError());
    u0 = u;
  }
  const int64_t n_buckets = (int64_t)T * a.KB;
  bucket_entries_kernel<<<(unsigned)((n_buckets + 3) / 4), 128, 0, st>>>(
      a.buckets, bucket_count_dev, n_buckets, a.KB, L, margin, N, zhat_dev, a.entries, a.sub_cap, entry_count_dev,
      row_nnz_dev, rowstats_dev, a.ovf_bkts, a.ovf_count, a.ovf_bkt_cap);
  // the buckets that overflowed, one by one (returns at once when there is none)
  bucket_overflow_kernel<<<74, 256, 0, st>>>(a.buckets, bucket_count_dev, a.KB, L, margin, N, zhat_dev, a.entries, a.sub_cap,
                                             entry_count_dev, row_nnz_dev, rowstats_dev, a.ovf, a.ovf_count, a.ovf_cap,
                                             a.ovf_bkts, a.ovf_bkt_cap);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

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
// Work decomposition: one CTA = one k-block (32 sequences) x 4 tasks; one warp = one
// task; in the filter pass one lane = one sequence walking the positions left to right.
#include "common.cuh"

namespace mepp {

struct HitRec { int32_t task, seq, pos; float x0; };

constexpr int kScanWarps = 4;
constexpr int kRowStride = 36;  // floats per position row of the x0 staging buffer (32 + 4 pad)
constexpr int kQueue = 64;      // candidate ring (flushed whenever 32 are pending)

struct ScanArgs {
  const uint32_t* sym_T; const float* zhat;
  int64_t N, n_pad, KB;
  int L, T, margin, W3, strideS;
  const float* lut; const float* bias; const int32_t* mlen; const int32_t* nfilt;
  const uint32_t* qtab; const uint32_t* qthr;
  uint8_t* x_planes; double* rowstats; int32_t* count; int32_t* density;
  HitRec* hits; int64_t hits_cap; unsigned long long* hit_count;
};

// Per-warp (= per task) state shared by the helper routines below.  The helpers are
// deliberately NOT inlined and NOT templated: one small instruction stream for every task keeps
// the kernel inside the instruction cache (an earlier variant specialised on motif length and
// strand count, 26k SASS instructions, and spent most of its time on instruction fetch).
struct TaskCtx {
  const float2* s_lut; const uint32_t* s_sym; float* s_buf; const float* s_z;
  uint16_t* s_queue;
  int t, kb, m, padl, mg, L, strideS, dual;
  float bias;
  int64_t tile_stride;   // bytes between consecutive m-tiles of the A operand
  uint8_t* x_base;       // A operand + this k-block's offset inside an m-tile row
  unsigned int* flags;   // non-zero bit array [m_tile][FW]
  int FW;
};

// Exact canonical-order score of the window starting at base i of one sequence (`stream` = that
// sequence's 3-bit symbol stream in shared memory).
__device__ __forceinline__ float exact_x0(const TaskCtx& c, const uint32_t* __restrict__ stream, int i) {
  float a0 = 0.0f, a1 = 0.0f;
#pragma unroll 1
  for (int j = 0; j < c.m; ++j) {
    const int bit = 3 * (i + j);
    const uint32_t sym = __funnelshift_r(stream[bit >> 5], stream[(bit >> 5) + 1], bit & 31) & 7u;
    const float2* l4 = c.s_lut + j * 4;
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
  }
  float s = __fadd_rn(a0, c.bias);
  if (c.dual) s = fmaxf(s, __fadd_rn(a1, c.bias));
  return fmaxf(s, 0.0f);
}

// Evaluate up to 32 queued candidates (one per lane) exactly; publish matches.  Returns (warp
// uniform) whether any candidate was a match.
__device__ __noinline__ bool flush_candidates(const ScanArgs& a, const TaskCtx& c, int q_head, int count, int base,
                                              int lane) {
  __syncwarp();
  bool any_hit = false;
  if (lane < count) {
    const uint32_t e = c.s_queue[(q_head + lane) & (kQueue - 1)];
    const int s = (int)(e >> 8), b = (int)(e & 0xffu);
    const int p = base - c.mg + b;
    const float x0 = exact_x0(c, c.s_sym + s * c.strideS, p - c.padl);
    if (x0 > 0.0f) {
      any_hit = true;
      c.s_buf[b * kRowStride + s] = x0;
      const int64_t ns = (int64_t)c.kb * 32 + s;
      atomicAdd(&a.count[(int64_t)c.t * c.L + p], 1);
      atomicAdd(&a.density[(int64_t)c.t * a.n_pad + ns], 1);
      if (a.hit_count) {
        const unsigned long long slot = atomicAdd(a.hit_count, 1ull);
        if (a.hits && (int64_t)slot < a.hits_cap) {
          HitRec rec; rec.task = c.t; rec.seq = (int32_t)ns; rec.pos = p; rec.x0 = x0;
          a.hits[slot] = rec;
        }
      }
    }
  }
  return __any_sync(0xffffffffu, any_hit);
}

// Phase B of one 32-position chunk: blur, row sums, fp16 split, tiled store; lane = (kc, rsub).
__device__ __noinline__ void store_chunk(const ScanArgs& a, const TaskCtx& c, int base, bool chunk_nz, int lane) {
  const unsigned full = 0xffffffffu;
  const int kc = lane & 3, rsub = lane >> 2;
  const int mg = c.mg, L = c.L;
  const int k = 1 + 2 * mg;
  const float kf = (float)k;
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    const int rl = rsub + 8 * it;
    const int p = base + rl;
    const bool valid = p < L;
    uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = vh;
    if (chunk_nz) {                    // warp uniform; most chunks hold no match at all
      const bool interior = valid && p >= mg && p <= L - 1 - mg;   // core.py:79-91 zero edges
      float xb[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) xb[q] = 0.0f;
      if (interior) {
        const float* row = c.s_buf + rl * kRowStride + kc * 8;  // buffer row rl+d <-> position p-mg+d
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
          sx += v; sxx += v * v; sxz += v * (double)c.s_z[kc * 8 + q];
        }
      }
      if (__any_sync(full, nz)) {
        sx += __shfl_xor_sync(full, sx, 1);   sx += __shfl_xor_sync(full, sx, 2);
        sxx += __shfl_xor_sync(full, sxx, 1); sxx += __shfl_xor_sync(full, sxx, 2);
        sxz += __shfl_xor_sync(full, sxz, 1); sxz += __shfl_xor_sync(full, sxz, 2);
        if (kc == 0 && valid && sxx != 0.0) {
          double* rs = a.rowstats + ((int64_t)c.t * L + p) * 3;
          atomicAdd(rs + 0, sx); atomicAdd(rs + 1, sxx); atomicAdd(rs + 2, sxz);
        }
      }
      if (nz) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          __half h0, l0, h1, l1;
          split_half(xb[2 * q] * MEPP_X_SCALE, h0, l0);
          split_half(xb[2 * q + 1] * MEPP_X_SCALE, h1, l1);
          h[q] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          l[q] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        vh = make_uint4(h[0], h[1], h[2], h[3]);
        vl = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
    if (valid) {
      const int r = c.t * L + p;                       // global row (< 2^31, checked on the host)
      uint8_t* tile = c.x_base + (int64_t)(r >> 7) * c.tile_stride;
      uint8_t* dst = tile + ((kc * kBlockM + (r & 127)) << 4);
      *reinterpret_cast<uint4*>(dst) = vh;
      *reinterpret_cast<uint4*>(dst + kChunks * kBlockM * 16) = vl;
      if ((vh.x | vh.y | vh.z | vh.w | vl.x | vl.y | vl.z | vl.w) != 0u) {  // rare: mark the K=16 block as non-zero
        const int bit = 2 * c.kb + (kc >> 1);
        atomicOr(c.flags + (int64_t)(r >> 7) * c.FW + (bit >> 5), 1u << (bit & 31));
      }
    }
  }
}

// Pair-slot lookup: slot T covers taps 2T, 2T+1 = bits [6T, 6T+6) of the window's symbol stream
// (w_lo = the stream word holding bit 6T, w_hi = the next word, only used by the two slots that
// straddle a word).  `qbase` is the 256-byte aligned shared address of the task's 16 x 256-byte
// tables, so the same LOP3 that masks the 6-bit index ORs the table base in (slot offset = LDS immediate).
template <int T, int U>
__device__ __forceinline__ uint32_t slot(uint32_t qbase, const uint32_t (&r)[6]) {
  constexpr int bit = 6 * T + 3 * U;       // window U bases ahead of the register file's origin
  constexpr int word = bit / 32, off = bit % 32;
  uint32_t v;
  if constexpr (off + 6 > 32) v = __funnelshift_r(r[word], r[word + 1], off - 2);
  else if constexpr (off >= 2) v = r[word] >> (off - 2);
  else v = r[word] << (2 - off);
  const uint32_t addr = (v & 0xFCu) | qbase;
  uint32_t x;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(x) : "r"(addr), "n"(T * 256));
  return x;
}

// Pre-filter sum of the window U bases ahead of the register file's origin (np = pair slots in use).
template <int U>
__device__ __forceinline__ uint32_t window_sum(uint32_t qbase, uint32_t qinit, int np, const uint32_t (&r)[6]) {
  uint32_t sum = qinit + slot<0, U>(qbase, r) + slot<1, U>(qbase, r) + slot<2, U>(qbase, r);
  if (np > 3) {
    sum += slot<3, U>(qbase, r) + slot<4, U>(qbase, r);
    if (np > 5) {
      sum += slot<5, U>(qbase, r);
      if (np > 6) {
        sum += slot<6, U>(qbase, r) + slot<7, U>(qbase, r);
        if (np > 8) {
          sum += slot<8, U>(qbase, r) + slot<9, U>(qbase, r) + slot<10, U>(qbase, r);
          if (np > 11)
            sum += slot<11, U>(qbase, r) + slot<12, U>(qbase, r) + slot<13, U>(qbase, r) +
                   slot<14, U>(qbase, r) + slot<15, U>(qbase, r);
        }
      }
    }
  }
  return sum;
}

template <int BITS>
__device__ __forceinline__ void advance(uint32_t (&r)[6]) {
  r[0] = __funnelshift_r(r[0], r[1], BITS); r[1] = __funnelshift_r(r[1], r[2], BITS);
  r[2] = __funnelshift_r(r[2], r[3], BITS); r[3] = __funnelshift_r(r[3], r[4], BITS);
  r[4] = __funnelshift_r(r[4], r[5], BITS); r[5] >>= BITS;
}

__device__ __forceinline__ void scan_task(const ScanArgs& a, const int t, const int kb,
                                          const uint32_t* __restrict__ s_sym, float2* __restrict__ s_lut,
                                          uint32_t* __restrict__ s_qtab, float* __restrict__ s_buf,
                                          uint16_t* __restrict__ s_queue, const float* __restrict__ s_z,
                                          TaskCtx* __restrict__ s_ctx, const int lane) {
  const unsigned full = 0xffffffffu;
  const int L = a.L, mg = a.margin;
  const int m = a.mlen[t];
  const int np = (m + 1) >> 1;           // pair slots in use
  const int padl = (m - 1) / 2;          // motif_layers.py:99-101
  const int n_valid = L - m + 1;         // 'valid' convolution outputs
  const bool real = (int64_t)kb * 32 + lane < a.N;

  const float2* g_lut = reinterpret_cast<const float2*>(a.lut) + (size_t)t * (MEPP_MAX_MOTIF_LEN * 4);
  for (int i = lane; i < MEPP_MAX_MOTIF_LEN * 4; i += 32) s_lut[i] = g_lut[i];
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.qtab + (size_t)t * 1024);
    uint4* dst = reinterpret_cast<uint4*>(s_qtab);
    for (int i = lane; i < 256; i += 32) dst[i] = src[i];
  }
  // qinit: per 16-bit half (0x8000 - pass bound), so "sum reaches the bound" <=> bit 15 of the half is set.
  // Padding sequences never produce candidates.
  const uint32_t qinit = real ? a.qthr[t * 2 + 0] : 0u;
  __syncwarp();

  if (lane == 0) {
    TaskCtx w;
    w.s_lut = s_lut; w.s_sym = s_sym; w.s_buf = s_buf; w.s_z = s_z; w.s_queue = s_queue;
    w.t = t; w.kb = kb; w.m = m; w.padl = padl; w.mg = mg; w.L = L; w.strideS = a.strideS;
    w.dual = a.nfilt[t] == 2;
    w.bias = a.bias[t];
    w.tile_stride = a.KB * (int64_t)kATileBytes;
    w.x_base = a.x_planes + (int64_t)kb * kATileBytes;
    w.FW = (int)flag_words(a.KB);
    w.flags = reinterpret_cast<unsigned int*>(
        a.x_planes + a_flags_offset(((int64_t)a.T * L + kBlockM - 1) / kBlockM, a.KB));
    *s_ctx = w;
  }
  __syncwarp();
  const TaskCtx& c = *s_ctx;

  const int nbuf = 32 + 2 * mg;
  const uint32_t* my = s_sym + lane * a.strideS;
  const uint32_t qbase = (uint32_t)__cvta_generic_to_shared(s_qtab);

  // 192-bit shift register over the symbol stream (window start at bit 0 of r[0]), advanced by four
  // bases (12 bits) per group of four windows and refilled with three new words every 32 positions.
  uint32_t r[6] = {my[0], my[1], my[2], my[3], my[4], my[5]};
  int next_w = 6, left = 32;
  int q_head = 0, q_tail = 0;            // warp-uniform ring indices
  bool carry_nz = false;                 // previous chunk left non-zero x0 in the carried rows

  for (int base = 0; base < L; base += 32) {
    int b_start = 0;
    if (base != 0) {
      // carry the last 2*margin rows of the previous chunk to the front
      for (int b = 0; b < 2 * mg; ++b) s_buf[b * kRowStride + lane] = s_buf[(32 + b) * kRowStride + lane];
      b_start = 2 * mg;
    }
    for (int b = b_start; b < nbuf; ++b) s_buf[b * kRowStride + lane] = 0.0f;
    bool new_nz = false;
    // ---- phase A: integer pre-filter over positions base-mg+b, candidates queued ----
    // window starts i = base - mg + b - padl; only 0 <= i < n_valid are real windows
    const int i_first = base - mg + b_start - padl;
    const int b_lo = b_start + (i_first < 0 ? -i_first : 0);
    int b_hi = nbuf;
    {
      const int i_end = base - mg + nbuf - padl;      // exclusive
      if (i_end > n_valid) b_hi -= i_end - n_valid;
    }
    auto refill = [&]() { r[3] = my[next_w]; r[4] = my[next_w + 1]; r[5] = my[next_w + 2]; next_w += 3; left = 32; };
    // queue the candidates of buffer row b (ballot `bal`), flushing the ring whenever 32 are pending
    auto push = [&](unsigned bal, bool cand, int b) {
      if (cand) {
        const int slot_i = q_tail + __popc(bal & ((1u << lane) - 1u));
        s_queue[slot_i & (kQueue - 1)] = (uint16_t)((lane << 8) | b);
      }
      q_tail += __popc(bal);
      if (q_tail - q_head >= 32) { new_nz |= flush_candidates(a, c, q_head, 32, base, lane); q_head += 32; }
    };
    int b = b_lo;
    // scalar steps until the refill counter is a multiple of 4, groups of 4 windows, scalar tail
#pragma unroll 1
    for (; b < b_hi && (left & 3) != 0; ++b) {
      const bool cand = (window_sum<0>(qbase, qinit, np, r) & 0x80008000u) != 0u;
      advance<3>(r);
      if (--left == 0) refill();
      const unsigned bal = __ballot_sync(full, cand);
      if (bal != 0u) push(bal, cand, b);
    }
#pragma unroll 1
    for (; b + 4 <= b_hi; b += 4) {
      // four independent windows read the same registers at offsets 0, 3, 6, 9 bits; their pass bits
      // (bit 15 / 31 of each packed sum) are merged into one word: window u lands on bits 15-u and 31-u
      const uint32_t cb = (window_sum<0>(qbase, qinit, np, r) & 0x80008000u) |
                          ((window_sum<1>(qbase, qinit, np, r) & 0x80008000u) >> 1) |
                          ((window_sum<2>(qbase, qinit, np, r) & 0x80008000u) >> 2) |
                          ((window_sum<3>(qbase, qinit, np, r) & 0x80008000u) >> 3);
      advance<12>(r);
      left -= 4;
      if (left == 0) refill();
      if (__any_sync(full, cb != 0u)) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool cand = (cb & (0x80008000u >> u)) != 0u;
          const unsigned bal = __ballot_sync(full, cand);
          if (bal != 0u) push(bal, cand, b + u);
        }
      }
    }
#pragma unroll 1
    for (; b < b_hi; ++b) {
      const bool cand = (window_sum<0>(qbase, qinit, np, r) & 0x80008000u) != 0u;
      advance<3>(r);
      if (--left == 0) refill();
      const unsigned bal = __ballot_sync(full, cand);
      if (bal != 0u) push(bal, cand, b);
    }
    if (q_tail != q_head) { new_nz |= flush_candidates(a, c, q_head, q_tail - q_head, base, lane); q_head = q_tail; }
    __syncwarp();
    store_chunk(a, c, base, carry_nz | new_nz, lane);
    carry_nz = new_nz;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kScanWarps * 32) scan_kernel(const __grid_constant__ ScanArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = blockIdx.x;
  const int nbuf = 32 + 2 * a.margin;
  // carve shared memory (the filter tables come first: the OR-merged index needs 256-byte alignment)
  uint8_t* sm = smem_raw + ((256u - (uint32_t)(__cvta_generic_to_shared(smem_raw) & 255u)) & 255u);
  uint32_t* s_qtabs = reinterpret_cast<uint32_t*>(sm);                                  // [warps][1024]
  float* s_bufs = reinterpret_cast<float*>(s_qtabs + kScanWarps * 1024);                // [warps][nbuf][36]
  float2* s_luts = reinterpret_cast<float2*>(s_bufs + kScanWarps * nbuf * kRowStride);  // [warps][128]
  uint32_t* s_sym = reinterpret_cast<uint32_t*>(s_luts + kScanWarps * MEPP_MAX_MOTIF_LEN * 4);
  float* s_z = reinterpret_cast<float*>(s_sym + 32 * a.strideS);                        // [32]
  uint16_t* s_queues = reinterpret_cast<uint16_t*>(s_z + 32);                           // [warps][kQueue]

  for (int idx = threadIdx.x; idx < a.W3 * 32; idx += blockDim.x) {
    const int s = idx & 31, w = idx >> 5;
    s_sym[s * a.strideS + w] = a.sym_T[(int64_t)w * a.n_pad + (int64_t)kb * 32 + s];
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
  uint32_t* s_qtab = s_qtabs + warp * 1024;
  uint16_t* s_queue = s_queues + warp * kQueue;
  __shared__ TaskCtx s_ctxs[kScanWarps];
  scan_task(a, t, kb, s_sym, s_lut, s_qtab, s_buf, s_queue, s_z, &s_ctxs[warp], lane);
}

}  // namespace mepp

extern "C" int mepp_scan(const uint32_t* sym_T_dev, const float* zhat_dev,
                         int64_t N, int L, int T, int margin, const float* lut_dev, const float* bias_dev,
                         const int32_t* mlen_dev, const int32_t* nfilt_dev, const uint32_t* qtab_dev,
                         const uint32_t* qthr_dev, void* x_planes_dev, double* rowstats_dev, int32_t* count_dev,
                         int32_t* density_dev, void* hits_dev, int64_t hits_cap, unsigned long long* hit_count_dev,
                         void* stream) {
  using namespace mepp;
  MEPP_REQUIRE(sym_T_dev && lut_dev && bias_dev && mlen_dev && nfilt_dev && qtab_dev && qthr_dev &&
               x_planes_dev && rowstats_dev && count_dev && density_dev, "scan: null pointer");
  MEPP_REQUIRE(N > 0 && L > 0 && T > 0, "scan: empty input");
  MEPP_REQUIRE((int64_t)T * L < (1LL << 31), "scan: T * L must be below 2^31");
  MEPP_REQUIRE(margin >= 0 && margin <= MEPP_MAX_MARGIN, "scan: margin must be in [0, 16]");
  MEPP_REQUIRE(hits_dev == nullptr || hit_count_dev != nullptr, "scan: hits need a hit counter");
  MEPP_REQUIRE(mepp_device_ok(), "scan: no sm_100 CUDA device (no CPU fallback)");
  cudaStream_t st = as_stream(stream);
  ScanArgs a;
  a.sym_T = sym_T_dev; a.zhat = zhat_dev;
  a.N = N; a.n_pad = mepp_pad32(N); a.KB = a.n_pad / kBlockK;
  a.L = L; a.T = T; a.margin = margin;
  a.W3 = (int)mepp_sym_words(L);
  a.strideS = a.W3 | 1;
  a.lut = lut_dev; a.bias = bias_dev; a.mlen = mlen_dev; a.nfilt = nfilt_dev;
  a.qtab = qtab_dev; a.qthr = qthr_dev;
  a.x_planes = reinterpret_cast<uint8_t*>(x_planes_dev); a.rowstats = rowstats_dev;
  a.count = count_dev; a.density = density_dev;
  a.hits = reinterpret_cast<HitRec*>(hits_dev); a.hits_cap = hits_cap; a.hit_count = hit_count_dev;
  const int nbuf = 32 + 2 * margin;
  size_t smem = sizeof(float) * kScanWarps * nbuf * kRowStride + sizeof(float2) * kScanWarps * MEPP_MAX_MOTIF_LEN * 4 +
                sizeof(uint32_t) * kScanWarps * 1024 + sizeof(uint32_t) * 32 * a.strideS +
                sizeof(float) * 32 + sizeof(uint16_t) * kScanWarps * kQueue + 256 /* table alignment */;
  MEPP_REQUIRE(smem <= 200 * 1024, "scan: sequence too long for the shared-memory staging (L too large)");
  MEPP_CUDA_CHECK(cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MEPP_CUDA_CHECK(cudaMemsetAsync(rowstats_dev, 0, sizeof(double) * 3 * (size_t)T * L, st));
  MEPP_CUDA_CHECK(cudaMemsetAsync(count_dev, 0, sizeof(int32_t) * (size_t)T * L, st));
  MEPP_CUDA_CHECK(cudaMemsetAsync(density_dev, 0, sizeof(int32_t) * (size_t)T * a.n_pad, st));
  if (hit_count_dev) MEPP_CUDA_CHECK(cudaMemsetAsync(hit_count_dev, 0, sizeof(unsigned long long), st));
  {
    // reset the non-zero bit array behind the tiles
    const int64_t n_tiles_m = ((int64_t)T * L + kBlockM - 1) / kBlockM;
    MEPP_CUDA_CHECK(cudaMemsetAsync(a.x_planes + a_flags_offset(n_tiles_m, a.KB), 0,
                                    (size_t)(n_tiles_m * flag_words(a.KB) * 4), st));
  }
  dim3 grid((unsigned)a.KB, (unsigned)((T + kScanWarps - 1) / kScanWarps));
  MEPP_REQUIRE(grid.y <= 65535, "scan: too many tasks in one call");
  scan_kernel<<<grid, kScanWarps * 32, smem, st>>>(a);
  MEPP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

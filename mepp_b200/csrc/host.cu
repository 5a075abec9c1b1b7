// Host-side pieces of the C ABI: error state, layout sizes, and the MOODS
// log-odds / p-value threshold DP that the reference runs on the host
// (mepp/motif_layers.py:36-48 -> MOODS.tools.log_odds / threshold_from_p).
#include "common.cuh"
#include <cmath>
#include <vector>
#include <algorithm>
#include <limits>

namespace mepp {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(const std::string& msg) { g_err = msg; return 1; }
}  // namespace mepp


namespace mepp {
// Conservative integer pre-filter for the scan kernel (csrc/scan.cu): per filter f and tap j the
// float32 weights are shifted to be >= 0 and quantised DOWN to a common grid, q = floor((W - min_c W) * scale),
// scale chosen so that the largest possible sum fits 15 bits.  For the real-valued score S of a window,
//   sum_q <= scale * (S - lo) < sum_q + m,        lo = sum_j min_c W[c][j],
// and the float32 sequentially rounded score differs from S by less than delta, so a window can only
// pass the threshold if sum_q >= qthr.  Both filters are packed into one 32-bit entry (low / high half),
// and two adjacent taps share one 64-entry table indexed by the 6-bit pair of 3-bit symbols (A,C,G,T,N).
// qpar (optional, 4 doubles): {scale, delta, lo[0], lo[1]}, what the pass bounds need besides the threshold.
static void build_filter_table(const float* lut, int m, int n_filters, float bias, uint32_t* qtab, uint32_t* qthr,
                               double* qpar = nullptr) {
  double rsum[2] = {0.0, 0.0}, lo[2] = {0.0, 0.0}, absmax = 0.0;
  double mn[2][MEPP_MAX_MOTIF_LEN];
  for (int f = 0; f < n_filters; ++f)
    for (int j = 0; j < m; ++j) {
      double a = lut[(j * 4 + 0) * 2 + f], b = a, am = std::fabs(a);
      for (int c = 1; c < 4; ++c) {
        double w = lut[(j * 4 + c) * 2 + f];
        a = std::min(a, w); b = std::max(b, w); am = std::max(am, std::fabs(w));
      }
      mn[f][j] = a; lo[f] += a; rsum[f] += b - a; absmax += am;
    }
  const double rmax = std::max(rsum[0], rsum[1]);
  const double scale = rmax > 0.0 ? 32767.0 / rmax * (1.0 - 1e-9) : 1.0;   // sums stay below 2^15
  // |float32 sequential sum - real sum| <= m roundings of at most ulp(absmax)/2 each; doubled for safety
  const double delta = 2.0 * (4.0 * m) * absmax * 5.97e-8 + 1e-6;   // up to 4 adds per tap (N bases)
  const double thr32 = -(double)bias;
  if (qpar) { qpar[0] = scale; qpar[1] = delta; qpar[2] = lo[0]; qpar[3] = lo[1]; }
  // symbol 4 = N: the four quarter-weight adds sum (in real arithmetic) to the column mean
  double qn[2][MEPP_MAX_MOTIF_LEN];
  for (int f = 0; f < n_filters; ++f)
    for (int j = 0; j < m; ++j) {
      double mean = 0.0;
      for (int c = 0; c < 4; ++c) mean += 0.25 * (double)lut[(j * 4 + c) * 2 + f];
      qn[f][j] = std::floor((mean - mn[f][j]) * scale);
      if (qn[f][j] < 0.0) qn[f][j] = 0.0;
    }
  const int n_slots = (m + 1) / 2;                     // slots past the motif hold zeros
  for (int k = n_slots * 64; k < 16 * 64; ++k) qtab[k] = 0u;
  for (int t = 0; t < n_slots; ++t)
    for (int e = 0; e < 64; ++e) {
      uint32_t packed = 0;
      for (int f = 0; f < n_filters; ++f) {
        uint32_t q = 0;
        for (int h = 0; h < 2; ++h) {
          const int j = 2 * t + h, sym = (e >> (3 * h)) & 7;
          if (j >= m) continue;
          if (sym < 4) q += (uint32_t)std::floor(((double)lut[(j * 4 + sym) * 2 + f] - mn[f][j]) * scale);
          else q += (uint32_t)qn[f][j];
        }
        packed |= (q & 0xffffu) << (16 * f);
      }
      qtab[t * 64 + e] = packed;
    }
  // qthr[0] = qinit: per 16-bit half (0x8000 - pass bound), added to the packed sum so that "the half reaches
  // its bound" shows up as bit 15 of the half; 0 = can never pass, 0x8000 = always passes.  qthr[1] keeps the
  // two pass bounds (low 16 / high 16 bits) for inspection.
  uint32_t qinit = 0, bounds = 0;
  for (int f = 0; f < n_filters; ++f) {
    const double need = scale * (thr32 - delta - lo[f]) - (double)m;  // candidates have sum_q > need
    long long qq = need < 0.0 ? 0 : (long long)std::floor(need) + 1;  // pass bound: sum_q >= qq
    uint32_t half = qq > 0x7fff ? 0u : (uint32_t)(0x8000 - qq);
    qinit |= half << (16 * f);
    bounds |= (uint32_t)(qq > 0xffff ? 0xffff : qq) << (16 * f);
  }
  qthr[0] = qinit;
  qthr[1] = bounds;
}
}  // namespace mepp

extern "C" {

int mepp_abi_version(void) { return MEPP_B200_ABI_VERSION; }
const char* mepp_last_error(void) { return mepp::g_err.c_str(); }

int mepp_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return 0; }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int64_t mepp_sym_words(int L) { return (3 * (int64_t)L + 31) / 32 + 12; }
int64_t mepp_pad32(int64_t n) { return (n + 31) / 32 * 32; }

int64_t mepp_x_planes_bytes(int64_t m_rows, int64_t n_pad) {
  int64_t mt = (m_rows + mepp::kBlockM - 1) / mepp::kBlockM;
  int64_t KB = n_pad / mepp::kBlockK;
  return mepp::a_flags_offset(mt, KB) + mt * mepp::flag_words(KB) * 4;
}

int mepp_choose_bn(int C, int* n_tiles_n) {
  int nt = (C + 255) / 256;
  if (nt < 1) nt = 1;
  int bn = ((C + nt - 1) / nt + 31) / 32 * 32;
  if (n_tiles_n) *n_tiles_n = nt;
  return bn;
}

// tiles of the B operand, followed by n_pad uint32 of scratch for mepp_build_y (the packed fp16 hi | lo of every score)
int64_t mepp_y_planes_bytes(int C, int64_t n_pad) {
  int nt = 0;
  int bn = mepp_choose_bn(C, &nt);
  return (int64_t)nt * (n_pad / mepp::kBlockK) * (int64_t)mepp::b_tile_bytes(bn) + n_pad * 4;
}

// MOODS.tools.log_odds: W[j][i] = ln((mat[j][i] + ps*bg[j]) / col_i) - ln(bg[j]).
int mepp_log_odds(const double* pwm, int m, const double* bg, double ps, double* W) {
  MEPP_REQUIRE(pwm && bg && W && m > 0, "log_odds: bad arguments");
  for (int i = 0; i < m; ++i) {
    double col = 0.0;
    for (int j = 0; j < 4; ++j) col += pwm[j * m + i] + ps * bg[j];
    for (int j = 0; j < 4; ++j)
      W[j * m + i] = std::log((pwm[j * m + i] + ps * bg[j]) / col) - std::log(bg[j]);
  }
  return 0;
}

// dst[r] += b * src[r]: the inner loop of the threshold DP (a few 1e6 updates per motif).  Compiled for AVX2 as
// well and dispatched at load time; every lane does the same multiply and add, so the results do not depend on
// the vector width.
#if defined(__GNUC__) && defined(__x86_64__) && !defined(__CUDA_ARCH__)
__attribute__((target_clones("avx2", "default")))
#endif
static void dp_axpy(double* __restrict__ dst, const double* __restrict__ src, double b, long n) {
  for (long r = 0; r < n; ++r) dst[r] += b * src[r];
}

// MOODS.tools.threshold_from_p: distribution of the integer-rounded score under
// bg by dynamic programming, then the smallest grid score whose upper tail <= p.
int mepp_threshold_from_p(const double* W, int m, const double* bg, double p,
                          double precision, double* thr) {
  MEPP_REQUIRE(W && bg && thr && m > 0 && precision > 0, "threshold_from_p: bad arguments");
  std::vector<long> mat(4 * (size_t)m);
  for (int k = 0; k < 4 * m; ++k) {
    double v = W[k];
    mat[k] = v > 0.0 ? (long)(precision * v + 0.5) : (long)(precision * v - 0.5);
  }
  long maxT = 0, minV = std::numeric_limits<long>::max();
  for (int i = 0; i < m; ++i) {
    long mx = mat[i], mn = mat[i];
    for (int j = 1; j < 4; ++j) {
      mx = std::max(mx, mat[j * m + i]);
      mn = std::min(mn, mat[j * m + i]);
    }
    maxT += mx;
    minV = std::min(minV, mn);
  }
  long R = maxT - (long)m * minV;
  MEPP_REQUIRE(R >= 0 && R < (1L << 28), "threshold_from_p: score range too large");
  std::vector<double> t0((size_t)R + 1, 0.0), t1((size_t)R + 1, 0.0);
  for (int j = 0; j < 4; ++j) t0[(size_t)(mat[j * m] - minV)] += bg[j];
  long hi = 0;  // highest reachable index so far (keeps the DP sparse in early columns)
  for (int j = 0; j < 4; ++j) hi = std::max(hi, mat[j * m] - minV);
  for (int i = 1; i < m; ++i) {
    long new_hi = 0;
    for (int j = 0; j < 4; ++j) {
      long s = mat[j * m + i] - minV;
      double b = bg[j];
      dp_axpy(t1.data() + s, t0.data(), b, hi + 1);
      new_hi = std::max(new_hi, hi + s);
    }
    std::fill(t0.begin(), t0.begin() + hi + 1, 0.0);
    t0.swap(t1);
    hi = new_hi;
  }
  double sum = 0.0;
  for (long r = R; r >= 0; --r) {
    sum += t0[(size_t)r];
    if (sum > p) { *thr = (double)(r + (long)m * minV + 1) / precision; return 0; }
  }
  *thr = (double)((long)m * minV) / precision;
  return 0;
}

static int motif_table_impl(const double* pwm, int m, int orientation_char,
                            double pseudocount, double pvalue, const double* bg_in, int use_flags,
                            double dp_precision, const double* thr_given, float* lut, float* bias, int* n_filters,
                            double* thr_out, uint32_t* qtab, uint32_t* qthr) {
  MEPP_REQUIRE(pwm && lut && bias && n_filters, "motif_table: null argument");
  MEPP_REQUIRE(m >= 1 && m <= MEPP_MAX_MOTIF_LEN, "motif_table: motif length must be in [1, 32]");
  double bg[4] = {0.25, 0.25, 0.25, 0.25};
  double ps = 1e-4, pv = 1e-4;  // core.py:513-534 never forwards the user's values
  if (use_flags) {
    ps = pseudocount; pv = pvalue;
    if (bg_in) for (int j = 0; j < 4; ++j) bg[j] = bg_in[j];
  }
  std::vector<double> W(4 * (size_t)m), Wr(4 * (size_t)m);
  if (mepp_log_odds(pwm, m, bg, ps, W.data())) return 1;
  for (int c = 0; c < 4; ++c)
    for (int j = 0; j < m; ++j) Wr[c * m + j] = W[(3 - c) * m + (m - 1 - j)];  // W[::-1, ::-1]
  const double* f0 = W.data();
  const double* f1 = nullptr;
  if (orientation_char == '+') {
  } else if (orientation_char == '-') {
    f0 = Wr.data();  // motif_layers.py:36-42; threshold from the reversed matrix (:44-48)
  } else {
    f1 = Wr.data();  // dual: filter 1 = W[::-1,::-1], shared bias -thr(W) (:66-74)
  }
  double thr = 0.0;
  if (thr_given) thr = *thr_given;
  else if (mepp_threshold_from_p(f0, m, bg, pv, dp_precision, &thr)) return 1;
  for (int k = 0; k < MEPP_MAX_MOTIF_LEN * 8; ++k) lut[k] = 0.0f;
  for (int j = 0; j < m; ++j)
    for (int c = 0; c < 4; ++c) {
      lut[(j * 4 + c) * 2 + 0] = (float)f0[c * m + j];
      lut[(j * 4 + c) * 2 + 1] = f1 ? (float)f1[c * m + j] : 0.0f;
    }
  *bias = (float)(-thr);
  *n_filters = f1 ? 2 : 1;
  if (thr_out) *thr_out = thr;
  if (qtab && qthr) mepp::build_filter_table(lut, m, *n_filters, *bias, qtab, qthr);
  return 0;
}

int mepp_prepare_tables(const double* pwms, const int32_t* mlen, const int32_t* ochar, int T, double pseudocount,
                        double pvalue, const double* bg_in, int use_flags, double dp_precision, float* lut,
                        int32_t* nfilt, uint32_t* qtab, double* qpar, int32_t* smat, double* dp_par) {
  MEPP_REQUIRE(pwms && mlen && ochar && lut && nfilt && qtab && qpar && smat && dp_par && T > 0,
               "prepare_tables: bad arguments");
  double bg[4] = {0.25, 0.25, 0.25, 0.25};
  double ps = 1e-4, pv = 1e-4;  // core.py:513-534 never forwards the user's values
  if (use_flags) {
    ps = pseudocount; pv = pvalue;
    if (bg_in) for (int j = 0; j < 4; ++j) bg[j] = bg_in[j];
  }
  for (int t = 0; t < T; ++t) {
    const int m = mlen[t];
    MEPP_REQUIRE(m >= 1 && m <= MEPP_MAX_MOTIF_LEN, "prepare_tables: motif length must be in [1, 32]");
    // the padded (4, 32) matrix of task t -> contiguous (4, m)
    double pwm[4 * MEPP_MAX_MOTIF_LEN], W[4 * MEPP_MAX_MOTIF_LEN], Wr[4 * MEPP_MAX_MOTIF_LEN];
    for (int c = 0; c < 4; ++c)
      for (int j = 0; j < m; ++j) pwm[c * m + j] = pwms[((size_t)t * 4 + c) * MEPP_MAX_MOTIF_LEN + j];
    if (mepp_log_odds(pwm, m, bg, ps, W)) return 1;
    for (int c = 0; c < 4; ++c)
      for (int j = 0; j < m; ++j) Wr[c * m + j] = W[(3 - c) * m + (m - 1 - j)];
    const double* f0 = W;
    const double* f1 = nullptr;
    if (ochar[t] == '-') f0 = Wr;
    else if (ochar[t] != '+') f1 = Wr;
    float* l = lut + (size_t)t * MEPP_MAX_MOTIF_LEN * 8;
    for (int k = 0; k < MEPP_MAX_MOTIF_LEN * 8; ++k) l[k] = 0.0f;
    for (int j = 0; j < m; ++j)
      for (int c = 0; c < 4; ++c) {
        l[(j * 4 + c) * 2 + 0] = (float)f0[c * m + j];
        l[(j * 4 + c) * 2 + 1] = f1 ? (float)f1[c * m + j] : 0.0f;
      }
    nfilt[t] = f1 ? 2 : 1;
    uint32_t qthr_unused[2];
    mepp::build_filter_table(l, m, nfilt[t], 0.0f, qtab + (size_t)t * 1024, qthr_unused, qpar + (size_t)t * 4);
    // integer matrix of filter 0 for the threshold DP, shifted to >= 0 (mepp_threshold_from_p)
    long mat[4 * MEPP_MAX_MOTIF_LEN];
    for (int k = 0; k < 4 * m; ++k) {
      const double v = f0[k];
      mat[k] = v > 0.0 ? (long)(dp_precision * v + 0.5) : (long)(dp_precision * v - 0.5);
    }
    long maxT = 0, minV = std::numeric_limits<long>::max();
    for (int i = 0; i < m; ++i) {
      long mx = mat[i], mn = mat[i];
      for (int j = 1; j < 4; ++j) { mx = std::max(mx, mat[j * m + i]); mn = std::min(mn, mat[j * m + i]); }
      maxT += mx;
      minV = std::min(minV, mn);
    }
    const long R = maxT - (long)m * minV;
    MEPP_REQUIRE(R >= 0 && R < (1L << 28), "prepare_tables: score range too large");
    int32_t* sm = smat + (size_t)t * MEPP_MAX_MOTIF_LEN * 4;
    for (int k = 0; k < MEPP_MAX_MOTIF_LEN * 4; ++k) sm[k] = 0;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < 4; ++j) sm[i * 4 + j] = (int32_t)(mat[j * m + i] - minV);
    double* dp = dp_par + (size_t)t * 8;
    dp[0] = bg[0]; dp[1] = bg[1]; dp[2] = bg[2]; dp[3] = bg[3];
    dp[4] = pv; dp[5] = dp_precision; dp[6] = (double)((long)m * minV); dp[7] = (double)R;
  }
  return 0;
}

int mepp_motif_table(const double* pwm, int m, int orientation_char,
                     double pseudocount, double pvalue, const double* bg_in, int use_flags,
                     double dp_precision, float* lut, float* bias, int* n_filters, double* thr_out,
                     uint32_t* qtab, uint32_t* qthr) {
  return motif_table_impl(pwm, m, orientation_char, pseudocount, pvalue, bg_in, use_flags, dp_precision, nullptr, lut,
                          bias, n_filters, thr_out, qtab, qthr);
}

int mepp_motif_table_with_threshold(const double* pwm, int m, int orientation_char, double pseudocount,
                                    const double* bg_in, int use_flags, double thr, float* lut, float* bias,
                                    int* n_filters, uint32_t* qtab, uint32_t* qthr) {
  return motif_table_impl(pwm, m, orientation_char, pseudocount, 0.0, bg_in, use_flags, 0.0, &thr, lut, bias, n_filters,
                          nullptr, qtab, qthr);
}

}  // extern "C"

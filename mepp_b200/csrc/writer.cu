// Host side of the output layout: tab-separated tables of numeric columns, written the way
// pandas.DataFrame.to_csv(path, sep='\t', index=False) writes them (mepp/single.py:223-241 writes positions_df.tsv and
// ranks_df.tsv of every task like that; at cfg 3 the 324 ranks_df tables have 100 000 rows each and pandas needs
// 0.2 s per table -- a minute per run for a 30 ms compute step).
//
// pandas formats a float column with numpy's str(): the shortest digit string that round-trips IN THE COLUMN'S OWN
// PRECISION (float32 columns get float32 digits), positional notation for 1e-4 <= |x| < 1e16 (float32: < 1e6) with at
// least one fractional digit ("3.0"), scientific otherwise with a two-digit exponent ("1e-05", "1.5e+16"); NaN is
// written as the empty string (na_rep=''; a table of ONE column gets "" there, the csv module's rule for a row that
// would otherwise be empty), infinities as inf / -inf.  std::to_chars yields the same shortest digit strings; only the
// layout rule is restated here.  No CUDA in this file: it is plain C++ behind the same C ABI.
#include "common.cuh"
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace mepp {

// digits / exponent of the shortest round-trip representation of |x| (x finite, non-zero): "d[.ddd]e[+-]XX"
template <typename F>
static inline char* put_float(char* out, F x) {
  if (std::isnan(x)) return out;                                  // na_rep = ''
  if (std::isinf(x)) {
    if (x < 0) *out++ = '-';
    std::memcpy(out, "inf", 3);
    return out + 3;
  }
  if (std::signbit(x)) { *out++ = '-'; x = -x; }
  if (x == (F)0) { std::memcpy(out, "0.0", 3); return out + 3; }
  char tmp[48];
  const auto res = std::to_chars(tmp, tmp + sizeof(tmp), x, std::chars_format::scientific);
  // tmp = d[.ddd]e[+-]XX
  char digits[32];
  int nd = 0;
  const char* p = tmp;
  for (; p < res.ptr && *p != 'e'; ++p)
    if (*p != '.') digits[nd++] = *p;
  int e = 0;
  {
    ++p;                                                          // 'e'
    const bool neg = *p == '-';
    ++p;
    for (; p < res.ptr; ++p) e = e * 10 + (*p - '0');
    if (neg) e = -e;
  }
  // numpy's str(): positional notation for 1e-4 <= |x| < 1e16 (float64) / < 1e6 (float32), decided on the VALUE
  // (float32(1e-4) lies below 1e-4 and prints as 1e-04)
  const double ax = (double)x;
  const double hi_cut = sizeof(F) == 4 ? 1.e6 : 1.e16;
  if (ax >= 1.e-4 && ax < hi_cut) {
    if (e >= 0) {
      for (int i = 0; i <= e; ++i) *out++ = i < nd ? digits[i] : '0';
      *out++ = '.';
      if (nd > e + 1) { std::memcpy(out, digits + e + 1, (size_t)(nd - e - 1)); out += nd - e - 1; }
      else *out++ = '0';
    } else {
      *out++ = '0'; *out++ = '.';
      for (int i = 0; i < -e - 1; ++i) *out++ = '0';
      std::memcpy(out, digits, (size_t)nd);
      out += nd;
    }
  } else {
    *out++ = digits[0];
    if (nd > 1) { *out++ = '.'; std::memcpy(out, digits + 1, (size_t)(nd - 1)); out += nd - 1; }
    *out++ = 'e';
    *out++ = e < 0 ? '-' : '+';
    const int ae = e < 0 ? -e : e;
    if (ae < 10) *out++ = '0';
    out = std::to_chars(out, out + 8, ae).ptr;
  }
  return out;
}

template <typename I>
static inline char* put_int(char* out, I v) { return std::to_chars(out, out + 24, v).ptr; }

}  // namespace mepp

extern "C" {

int mepp_write_tsv(const char* path, const char* header_line, int n_cols, const int32_t* col_kind,
                   const void* const* col_ptr, const int64_t* col_stride_bytes, int64_t n_rows) {
  using namespace mepp;
  MEPP_REQUIRE(path && header_line && n_cols > 0 && col_kind && col_ptr && col_stride_bytes && n_rows >= 0,
               "write_tsv: bad arguments");
  for (int c = 0; c < n_cols; ++c)
    MEPP_REQUIRE(col_kind[c] >= 0 && col_kind[c] <= 5 && (col_ptr[c] || n_rows == 0), "write_tsv: unknown column kind");
  std::FILE* f = std::fopen(path, "wb");
  MEPP_REQUIRE(f != nullptr, "write_tsv: cannot open the output file");
  std::string buf;
  const size_t hl = std::strlen(header_line);
  const int64_t chunk_rows = 4096;
  buf.resize(hl + 1 + (size_t)chunk_rows * (size_t)n_cols * 28);
  char* out = &buf[0];
  std::memcpy(out, header_line, hl);
  out += hl;
  *out++ = '\n';
  bool ok = true;
  for (int64_t r0 = 0; r0 < n_rows && ok; r0 += chunk_rows) {
    const int64_t r1 = r0 + chunk_rows < n_rows ? r0 + chunk_rows : n_rows;
    for (int64_t r = r0; r < r1; ++r) {
      const char* const row_start = out;
      for (int c = 0; c < n_cols; ++c) {
        const char* src = static_cast<const char*>(col_ptr[c]) + r * col_stride_bytes[c];
        switch (col_kind[c]) {
          case 0: { double v; std::memcpy(&v, src, 8); out = put_float(out, v); break; }
          case 1: { float v; std::memcpy(&v, src, 4); out = put_float(out, v); break; }
          case 2: { int64_t v; std::memcpy(&v, src, 8); out = put_int(out, v); break; }
          case 3: { int32_t v; std::memcpy(&v, src, 4); out = put_int(out, v); break; }
          case 4: { uint16_t v; std::memcpy(&v, src, 2); out = put_int(out, (uint32_t)v); break; }
          default: { uint16_t v; std::memcpy(&v, src, 2); out = put_float(out, (double)v); break; }   // uint16 as float64
        }
        if (n_cols == 1 && out == row_start) { *out++ = '"'; *out++ = '"'; }
        *out++ = c + 1 < n_cols ? '\t' : '\n';
      }
    }
    const size_t n = (size_t)(out - &buf[0]);
    ok = std::fwrite(&buf[0], 1, n, f) == n;
    out = &buf[0];
  }
  if (n_rows == 0) ok = std::fwrite(&buf[0], 1, (size_t)(out - &buf[0]), f) == (size_t)(out - &buf[0]);
  ok = (std::fclose(f) == 0) && ok;
  MEPP_REQUIRE(ok, "write_tsv: write failed");
  return 0;
}

}  // extern "C"

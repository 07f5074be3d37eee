// Text formatting of one MatrixMarket coordinate entry ("row col value\n").
//
// Replaces the body writer behind scipy.io.mmwrite at the reference's call site
// (stellarscope/utils/model.py:1418).  The functions are __host__ __device__ so
// that the digit generation can be unit-tested on a machine without a GPU (tests/tools/mtx_fmt_host.cpp
// compiles this header with g++); the product path is the kernel in ss_mtx.cu.
//
// Integers: plain decimal.  Reals: the correctly rounded (round-half-even on the exact binary value)
// 17-significant-digit decimal, fixed notation, trailing zeros removed -- every double parses back to
// itself.  scipy's writer prints the shortest round-trip form ("1.4E1"); the files are equal as
// matrices, not as bytes (compare by content).
// Supported magnitudes: 0 and 2^-64 <= |x| < 2^63 (count matrices hold sums of 1/nbest fractions and
// integers); anything else is reported as an error by the caller, never printed wrong.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SS_HD __host__ __device__ inline
#else
#define SS_HD inline
#endif

namespace ssmtx {

constexpr int MAX_INT_CHARS = 20;      // sign + 19 digits
constexpr int MAX_REAL_CHARS = 64;     // sign + 19 integer digits + '.' + (19 leading zeros + 17 digits), rounded up
constexpr int MAX_LINE_CHARS = 11 + 1 + 11 + 1 + MAX_REAL_CHARS + 1;

// indices and counts fit 32 bits: comparisons and 32-bit multiply-high divisions instead of 64-bit division loops
SS_HD int ndigits_u32(uint32_t v) {
  return v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6
       : v < 10000000u ? 7 : v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}

SS_HD int ndigits_u64(uint64_t v) {
  if ((v >> 32) == 0) return ndigits_u32((uint32_t)v);
  int n = 1;
  while (v >= 10) { v /= 10; ++n; }
  return n;
}

// decimal digits of v at p (no terminator); returns the number of characters
SS_HD int put_u64(char* p, uint64_t v) {
  const int n = ndigits_u64(v);
  int i = n - 1;
  while ((v >> 32) != 0) { p[i--] = (char)('0' + (int)(v % 10)); v /= 10; }
  uint32_t w = (uint32_t)v;
  for (; i >= 0; --i) { const uint32_t q = w / 10u; p[i] = (char)('0' + (int)(w - 10u * q)); w = q; }
  return n;
}

SS_HD int put_i64(char* p, long long v) {
  if (v < 0) {
    *p = '-';
    return 1 + put_u64(p + 1, (uint64_t)(-(v + 1)) + 1u);
  }
  return put_u64(p, (uint64_t)v);
}

SS_HD bool real_supported(double x) {
  union { double d; uint64_t u; } b;
  b.d = x;
  const int bexp = (int)((b.u >> 52) & 0x7ff);
  if ((b.u << 1) == 0) return true;               // +-0
  return bexp >= 1023 - 64 && bexp < 1023 + 63;
}

// x must satisfy real_supported(x).  Returns the number of characters written to p (<= MAX_REAL_CHARS).
SS_HD int put_real(char* p, double x) {
  union { double d; uint64_t u; } b;
  b.d = x;
  int n = 0;
  if (b.u >> 63) p[n++] = '-';
  if ((b.u << 1) == 0) { p[n++] = '0'; return n; }
  const int bexp = (int)((b.u >> 52) & 0x7ff);
  const uint64_t m = (b.u & 0xfffffffffffffull) | (1ull << 52);
  const int e = bexp - 1075;                       // |x| = m * 2^e,  -116 <= e <= 10
  if (e >= 0) return n + put_u64(p + n, m << e);   // an integer below 2^63: exact
  const int s = -e;                                // 1..116
  typedef unsigned __int128 u128;
  const u128 M = (u128)m;
  uint64_t ip = s < 64 ? (uint64_t)(M >> s) : 0;   // m < 2^53, so the integer part is 0 for s >= 53
  const u128 one = (u128)1 << s;
  u128 frac = M & (one - 1);
  int sig = ip ? ndigits_u64(ip) : 0;              // significant digits so far
  char fd[40];                                     // fractional digits
  int nf = 0;
  while (frac != 0 && sig < 17) {
    frac *= 10;                                    // < 2^116 * 10 < 2^120
    const int d = (int)(frac >> s);
    frac &= one - 1;
    fd[nf++] = (char)('0' + d);
    if (sig > 0 || d != 0) ++sig;
  }
  if (frac != 0) {
    // round half to even on the exact remainder
    const u128 half = one >> 1;
    const bool last_odd = nf > 0 ? ((fd[nf - 1] - '0') & 1) : (ip & 1);
    if (frac > half || (frac == half && last_odd)) {
      int i = nf - 1;
      while (i >= 0 && fd[i] == '9') fd[i--] = '0';
      if (i >= 0) ++fd[i];
      else ++ip;                                   // carry into the integer part (ip < 2^63)
    }
  }
  while (nf > 0 && fd[nf - 1] == '0') --nf;
  n += put_u64(p + n, ip);
  if (nf > 0) {
    p[n++] = '.';
    for (int i = 0; i < nf; ++i) p[n++] = fd[i];
  }
  return n;
}

// length of the line put_line writes
SS_HD int line_len(int row, int col, double v, bool integer) {
  int n = ndigits_u64((uint64_t)((long long)row + 1)) + ndigits_u64((uint64_t)((long long)col + 1)) + 3;
  if (integer) {
    const long long iv = (long long)v;
    return n + (iv < 0 ? 1 + ndigits_u64((uint64_t)(-(iv + 1)) + 1u) : ndigits_u64((uint64_t)iv));
  }
  char tmp[MAX_REAL_CHARS];
  return n + put_real(tmp, v);
}

// one line: "<row+1> <col+1> <value>\n" (MatrixMarket indices are 1-based)
SS_HD int put_line(char* p, int row, int col, double v, bool integer) {
  int n = put_u64(p, (uint64_t)((long long)row + 1));
  p[n++] = ' ';
  n += put_u64(p + n, (uint64_t)((long long)col + 1));
  p[n++] = ' ';
  n += integer ? put_i64(p + n, (long long)v) : put_real(p + n, v);
  p[n++] = '\n';
  return n;
}

}  // namespace ssmtx

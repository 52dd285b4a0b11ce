// b2s_codec.cpp -- body codec of the serving boundary (host code, no CUDA).
//
// For HTTP / stream triggers the reference json-decodes every request body (GraphServer.run,
// serving/server.py:262-277) and json.dumps every response (_process_response, :298-308); for V2 bodies
// {"inputs": [[...], ...]} that is where a worker's time goes once the model is fast.
// This file parses the "inputs" matrix of a body straight into float32 rows (the layout b2s_submit /
// b2s_run_host take) and prints result matrices the way json.dumps does:
//   * numbers are converted with std::from_chars<double> (correctly rounded, like Python's float()) and then
//     rounded to float32 -- the same two roundings as np.asarray(json.loads(body)["inputs"], dtype=float32);
//   * floats are printed with the shortest round-trip digits (std::to_chars) laid out by Python's repr rule
//     (fixed notation for 1e-4 <= |x| < 1e16, else d.ddde+XX), so the text is byte-identical to json.dumps.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <charconv>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <limits>
#include <mutex>
#include <pthread.h>
#include <sched.h>
#include <thread>
#include <vector>

#include "../../include/b200serve.h"
#include "b2s_internal.h"

namespace {

struct Cur {
  const char* p;
  const char* end;
  void ws() {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
  }
  bool eat(char c) {
    ws();
    if (p < end && *p == c) {
      ++p;
      return true;
    }
    return false;
  }
};

bool skip_string(Cur& c) {  // at the opening quote
  if (c.p >= c.end || *c.p != '"') return false;
  ++c.p;
  while (c.p < c.end) {
    const char ch = *c.p++;
    if (ch == '\\') {
      if (c.p >= c.end) return false;
      ++c.p;
    } else if (ch == '"') {
      return true;
    }
  }
  return false;
}

bool skip_value(Cur& c, int depth = 0);

bool skip_container(Cur& c, char open, char close, int depth) {
  if (depth > 64) return false;
  ++c.p;  // open
  c.ws();
  if (c.p < c.end && *c.p == close) {
    ++c.p;
    return true;
  }
  for (;;) {
    c.ws();
    if (open == '{') {
      if (!skip_string(c)) return false;
      if (!c.eat(':')) return false;
    }
    if (!skip_value(c, depth + 1)) return false;
    c.ws();
    if (c.p >= c.end) return false;
    if (*c.p == ',') {
      ++c.p;
      continue;
    }
    if (*c.p == close) {
      ++c.p;
      return true;
    }
    return false;
  }
}

bool skip_value(Cur& c, int depth) {
  c.ws();
  if (c.p >= c.end) return false;
  const char ch = *c.p;
  if (ch == '"') return skip_string(c);
  if (ch == '{') return skip_container(c, '{', '}', depth);
  if (ch == '[') return skip_container(c, '[', ']', depth);
  const char* s = c.p;  // literal / number: up to a delimiter
  while (c.p < c.end && *c.p != ',' && *c.p != ']' && *c.p != '}' && *c.p != ' ' && *c.p != '\t' && *c.p != '\n' && *c.p != '\r') ++c.p;
  return c.p > s;
}

// ---- fast path for ordinary decimal numbers -------------------------------------------------------------------------------
// One pass over the token collects its digits (at most 19 after leading zeros) in a uint64 (w) and the decimal exponent (e10),
// checking the JSON number grammar on the way; value = w * 10^e10.  For |e10| <= 27 both w (< 2^64) and 10^|e10| (5^27 < 2^63) are exact in the
// x87 80-bit format, so ONE multiplication or division there is the correctly rounded 64-bit significand of the exact value.
// Rounding that to double is the correctly rounded double (what json.loads / float() produce) unless the 64-bit result sits
// exactly on a double's rounding boundary (low 11 bits == 0x400): rounding to 64 bits is monotonic and the boundary itself is
// a 64-bit number, so the exact value can only be on the other side of a boundary the result does not touch.  Those tokens
// (1 in 2048), longer digit strings, larger exponents, literals and everything malformed go to the general path below, which
// also decides every error.  Integers of up to 19 digits are converted by the uint64 -> double instruction (one rounding).
#if defined(__x86_64__) && LDBL_MANT_DIG == 64
#define B2S_CODEC_FAST 1
const long double kPow10[28] = {1e0L,  1e1L,  1e2L,  1e3L,  1e4L,  1e5L,  1e6L,  1e7L,  1e8L,  1e9L,  1e10L, 1e11L, 1e12L, 1e13L,
                                1e14L, 1e15L, 1e16L, 1e17L, 1e18L, 1e19L, 1e20L, 1e21L, 1e22L, 1e23L, 1e24L, 1e25L, 1e26L, 1e27L};

bool x87_is_extended() {  // the precision-control field of this thread's x87 unit (Linux default: 64-bit significands)
  volatile long double a = 1.0L, b = 0x1p-63L;
  volatile long double c = a + b;
  return c != a;
}

const uint64_t kPow10u[20] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, 1000000000ull,
                             10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull, 100000000000000ull,
                             1000000000000000ull, 10000000000000000ull, 100000000000000000ull, 1000000000000000000ull,
                             10000000000000000000ull};

// Appends the run of decimal digits at p to w (w = w * 10^n + digits), eight bytes at a time: the number of leading digit
// bytes of a chunk comes from one SWAR test + count-trailing-zeros, the digits are shifted to the end of an eight-digit
// field (zeros in front) and converted by three multiplications -- no per-digit branch to mispredict.  Returns the end of
// the run, or nullptr when w would pass 19 digits.
inline const char* append_digits(const char* p, const char* end, uint64_t& w) {
  while (end - p >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    uint64_t x = v ^ 0x3030303030303030ull;  // digit bytes become 0..9
    // bit 7 of a byte is set when the byte is > 9 (a carry out of such a byte can only disturb LATER bytes)
    const uint64_t stop = ((x + 0x7676767676767676ull) | x) & 0x8080808080808080ull;
    const int n = stop ? (__builtin_ctzll(stop) >> 3) : 8;
    if (w >= kPow10u[19 - n]) return n ? nullptr : p;
    const int sh = (8 - n) * 4;
    x = (x << sh) << sh;  // first digit in the low byte: the n digits move to the last n places
    x = (x * 10) + (x >> 8);
    const uint64_t mask = 0x000000FF000000FFull;
    const uint64_t val = (((x & mask) * 0x000F424000000064ull) + (((x >> 16) & mask) * 0x0000271000000001ull)) >> 32;
    w = w * kPow10u[n] + val;
    p += n;
    if (n < 8) return p;
  }
  unsigned d;
  while (p < end && (d = (unsigned char)*p - '0') <= 9) {  // the last bytes of the buffer (or of a row)
    if (w >= kPow10u[18]) return nullptr;
    w = w * 10 + d;
    ++p;
  }
  return p;
}

// 1: *out set, c.p moved past the token; 0: not taken (c.p unchanged)
inline int parse_number_fast(Cur& c, float* out) {
  static const bool ok = x87_is_extended();
  if (!ok) return 0;
  const char* p = c.p;
  const char* const end = c.end;
  if (p >= end) return 0;
  const uint64_t neg = *p == '-';
  p += neg;
  uint64_t w = 0;
  const char* const i0 = p;
  p = append_digits(p, end, w);
  if (!p || p == i0) return 0;                  // too long / not a digit (literals, garbage)
  if (*i0 == '0' && p - i0 > 1) return 0;       // leading zero
  long e10 = 0;
  bool integral = true;
  if (p < end && *p == '.') {
    integral = false;
    const char* const f0 = ++p;
    p = append_digits(p, end, w);
    if (!p || p == f0) return 0;                // too long / "5."
    e10 = -(long)(p - f0);
  }
  if (p < end && (*p == 'e' || *p == 'E')) {
    integral = false;
    ++p;
    bool eneg = false;
    if (p < end && (*p == '+' || *p == '-')) {
      eneg = *p == '-';
      ++p;
    }
    const char* x0 = p;
    long ex = 0;
    unsigned d;
    while (p < end && (d = (unsigned char)*p - '0') <= 9) {
      if (ex > 9999) return 0;
      ex = ex * 10 + d;
      ++p;
    }
    if (p == x0) return 0;
    e10 += eneg ? -ex : ex;
  }
  double v;
  uint64_t sign = neg << 63;
  if (integral) {
    v = (double)w;
    if (w == 0) sign = 0;  // "-0" is the int 0
  } else if (w == 0) {
    v = 0.0;
  } else {
    if (e10 < -27 || e10 > 27) return 0;
    const long double y = e10 < 0 ? (long double)w / kPow10[-e10] : (long double)w * kPow10[e10];
    uint64_t m;
    memcpy(&m, &y, 8);
    if ((m & 0x7FFull) == 0x400ull) return 0;  // on a double's rounding boundary: let the exact converter decide
    v = (double)y;
  }
  uint64_t bits;
  memcpy(&bits, &v, 8);
  bits |= sign;
  memcpy(&v, &bits, 8);
  *out = (float)v;
  c.p = p;
  return 1;
}
#else
#define B2S_CODEC_FAST 0
#endif

// one JSON number (json.loads grammar, plus the NaN / Infinity / -Infinity literals it accepts, plus null -> NaN)
bool parse_number(Cur& c, float* out) {
  c.ws();
#if B2S_CODEC_FAST
  if (parse_number_fast(c, out)) return true;
#endif
  const char* s = c.p;
  if (s >= c.end) return false;
  auto lit = [&](const char* w, float v) {
    const size_t n = strlen(w);
    if ((size_t)(c.end - s) >= n && memcmp(s, w, n) == 0) {
      c.p = s + n;
      *out = v;
      return true;
    }
    return false;
  };
  if (*s == 'N') return lit("NaN", std::numeric_limits<float>::quiet_NaN());
  if (*s == 'I') return lit("Infinity", std::numeric_limits<float>::infinity());
  if (*s == 'n') return lit("null", std::numeric_limits<float>::quiet_NaN());
  if (*s == '-' && s + 1 < c.end && s[1] == 'I') {
    ++s;
    if (lit("Infinity", -std::numeric_limits<float>::infinity())) return true;
    return false;
  }
  // std::from_chars does the conversion and finds the end of the token; what it accepts beyond the JSON grammar (leading
  // zeros, ".5", "5.", "inf" / "nan") is rejected by three checks on the span it consumed -- no second pass over the digits
  const bool neg = *s == '-';
  const char* t = neg ? s + 1 : s;
  if (t >= c.end || *t < '0' || *t > '9') return false;
  double d = 0.0;
  auto r = std::from_chars(s, c.end, d);
  if (r.ec == std::errc::invalid_argument) return false;
  const char* q = r.ptr;
  if (*t == '0' && t + 1 < q && t[1] >= '0' && t[1] <= '9') return false;  // 007
  const char* dot = t;
  while (dot < q && *dot >= '0' && *dot <= '9') ++dot;  // end of the integer digits
  if (dot < q && *dot == '.' && !(dot + 1 < q && dot[1] >= '0' && dot[1] <= '9')) return false;  // "5." / "5.e3"
  auto is_integral = [&] {  // json.loads makes an int of it: "-0" is 0, not -0.0
    for (const char* u = t; u < q; ++u)
      if (*u == '.' || *u == 'e' || *u == 'E') return false;
    return true;
  };
  if (r.ec == std::errc::result_out_of_range) {
    if (is_integral()) return false;  // an int too large for a double: np.asarray raises OverflowError; leave it to that path
    // json.loads gives +-inf for 1e999 and +-0.0 for 1e-999 (or 0.000...1 with 400 zeros): decided by the decimal
    // exponent of the first non-zero digit
    const char* nz = t;
    while (nz < q && (*nz == '0' || *nz == '.')) ++nz;
    long lead = nz < dot ? (long)(dot - nz - 1) : -(long)(nz - dot);
    const char* ep = dot;
    while (ep < q && *ep != 'e' && *ep != 'E') ++ep;
    if (ep < q) {
      long ex = 0;
      const char* u = ep + 1;
      const bool eneg = u < q && *u == '-';
      if (u < q && (*u == '+' || *u == '-')) ++u;
      for (; u < q; ++u)
        if (ex < 100000000L) ex = ex * 10 + (*u - '0');
      lead += eneg ? -ex : ex;
    }
    d = lead < 0 ? (neg ? -0.0 : 0.0) : (neg ? -HUGE_VAL : HUGE_VAL);
  } else if (r.ec != std::errc()) {
    return false;
  }
  if (d == 0.0 && is_integral()) d = 0.0;  // "-0" is the int 0
  *out = (float)d;
  c.p = q;
  return true;
}

// Python repr(float) of a double, into buf; returns the length
int repr_double(double v, char* buf) {
  if (std::isnan(v)) return (int)(stpcpy(buf, "NaN") - buf);  // json.dumps spelling
  if (std::isinf(v)) return (int)(stpcpy(buf, v < 0 ? "-Infinity" : "Infinity") - buf);
  char sci[40];
  auto r = std::to_chars(sci, sci + sizeof(sci), v, std::chars_format::scientific);  // shortest round-trip digits
  *r.ptr = 0;
  char* o = buf;
  const char* s = sci;
  if (*s == '-') *o++ = *s++;
  char digits[24];
  int nd = 0;
  digits[nd++] = *s++;
  if (*s == '.') {
    ++s;
    while (*s && *s != 'e') digits[nd++] = *s++;
  }
  int exp10 = 0;  // after 'e': sign and at least two digits
  {
    const char* x = s + 1;
    const bool xneg = *x == '-';
    if (*x == '-' || *x == '+') ++x;
    while (*x >= '0' && *x <= '9') exp10 = exp10 * 10 + (*x++ - '0');
    if (xneg) exp10 = -exp10;
  }
  if (exp10 >= -4 && exp10 < 16) {
    if (exp10 < 0) {
      *o++ = '0';
      *o++ = '.';
      for (int i = 0; i < -exp10 - 1; ++i) *o++ = '0';
      for (int i = 0; i < nd; ++i) *o++ = digits[i];
    } else {
      for (int i = 0; i <= exp10; ++i) *o++ = i < nd ? digits[i] : '0';
      *o++ = '.';
      if (nd > exp10 + 1) {
        for (int i = exp10 + 1; i < nd; ++i) *o++ = digits[i];
      } else {
        *o++ = '0';
      }
    }
  } else {
    *o++ = digits[0];
    if (nd > 1) {
      *o++ = '.';
      for (int i = 1; i < nd; ++i) *o++ = digits[i];
    }
    *o++ = 'e';
    *o++ = exp10 < 0 ? '-' : '+';
    const int a = exp10 < 0 ? -exp10 : exp10;
    if (a < 10) *o++ = '0';
    o = std::to_chars(o, o + 8, a).ptr;
  }
  return (int)(o - buf);
}

int codec_threads() {
  static const int n = [] {
    if (const char* e = getenv("B2S_CODEC_THREADS")) return atoi(e) < 1 ? 1 : atoi(e);
    cpu_set_t set;
    int cpus = 1;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cpus = CPU_COUNT(&set);
    return cpus > 8 ? 8 : (cpus < 1 ? 1 : cpus);
  }();
  return n;
}

// ---- rows of a numeric matrix, two stages (x86-64 with AVX2; anything else takes the scalar row loop) --------------------
// A scalar parser is one long dependency chain: where token k + 1 starts is known only when token k has been scanned, about
// 14 cycles per eight-byte step and four steps per 17-digit number.  Stage 1 breaks the chain: 32 bytes at a time are
// classified as separator (comma / white space) or token byte, and the token boundaries fall out of the bit masks
// (throughput bound, no conversion).  Stage 2 converts every token from its known extent: sign, position of the '.', digit
// runs of known length converted eight at a time at fixed offsets -- independent work the core overlaps across tokens.  It
// takes plain decimals ("-12.345", up to 8 integer and 19 total digits); exponents, literals (NaN, Infinity, null), longer
// numbers and anything malformed go to parse_number token by token, which also rejects what is not a number.
#if B2S_CODEC_FAST
#define B2S_AVX2 __attribute__((target("avx2,bmi,bmi2,lzcnt,popcnt")))

bool cpu_has_avx2() {
  static const bool ok = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi") && __builtin_cpu_supports("bmi2") &&
                         __builtin_cpu_supports("popcnt") && getenv("B2S_CODEC_SCALAR") == nullptr;
  return ok;
}

inline uint64_t digits_value(const char* p, int n) {  // n <= 8 digit bytes at p (eight bytes are read)
  uint64_t x;
  memcpy(&x, p, 8);
  x ^= 0x3030303030303030ull;
  const int sh = (8 - n) * 4;
  x = (x << sh) << sh;
  x = (x * 10) + (x >> 8);
  const uint64_t mask = 0x000000FF000000FFull;
  return (((x & mask) * 0x000F424000000064ull) + (((x >> 16) & mask) * 0x0000271000000001ull)) >> 32;
}

// shuffle controls that move the first n bytes of a 16-byte register to its end and clear the rest (n = 0 .. 16)
struct RightAlign16 {
  alignas(16) uint8_t m[17][16];
  constexpr RightAlign16() : m() {
    for (int n = 0; n <= 16; ++n)
      for (int i = 0; i < 16; ++i) m[n][i] = i >= 16 - n ? (uint8_t)(i - (16 - n)) : (uint8_t)0x80;
  }
  const uint8_t* operator[](int n) const { return m[n]; }
};
constexpr RightAlign16 kRightAlign16{};

// the token [s, e) as a plain decimal; false = not taken (the caller asks parse_number).  40 readable bytes from s are required.
B2S_AVX2 inline bool convert_plain_decimal(const char* s, const char* e, float* out) {
  const uint64_t neg = *s == '-';
  const char* q = s + neg;
  const int n = (int)(e - q);
  if (n < 1 || n > 21) return false;
  const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(q));
  const uint32_t in_token = (uint32_t)((1ull << n) - 1);
  const uint32_t dots = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('.'))) & in_token;
  const __m256i t = _mm256_sub_epi8(v, _mm256_set1_epi8('0'));
  const uint32_t digs = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_min_epu8(t, _mm256_set1_epi8(9)), t)) & in_token;
  if ((digs | dots) != in_token || (dots & (dots - 1))) return false;  // exponent, literal, garbage, two dots
  const int ni = dots ? (int)_tzcnt_u32(dots) : n;
  const int nf = dots ? n - ni - 1 : 0;
  if (ni < 1 || ni > 8 || ni + nf > 19 || (dots && nf < 1) || (*q == '0' && ni > 1)) return false;
  uint64_t w = digits_value(q, ni);
  double val;
  uint64_t sign = neg << 63;
  if (!dots) {
    val = (double)w;
    if (w == 0) sign = 0;  // "-0" is the int 0
  } else {
    // the fraction: its first min(nf, 16) digits in ONE 128-bit step -- moved to the end of a 16-digit field (zeros in front) by a
    // byte shuffle, then pairs -> fours -> eights by three multiply-adds; a 17th / 18th digit is appended one at a time
    const char* f = q + ni + 1;
    const int n16 = nf < 16 ? nf : 16;
    __m128i d = _mm_sub_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(f)), _mm_set1_epi8('0'));
    d = _mm_shuffle_epi8(d, _mm_load_si128(reinterpret_cast<const __m128i*>(kRightAlign16[n16])));
    const __m128i pairs = _mm_maddubs_epi16(d, _mm_set_epi8(1, 10, 1, 10, 1, 10, 1, 10, 1, 10, 1, 10, 1, 10, 1, 10));
    const __m128i fours = _mm_madd_epi16(pairs, _mm_set_epi16(1, 100, 1, 100, 1, 100, 1, 100));
    const __m128i packed = _mm_packus_epi32(fours, fours);
    const __m128i eights = _mm_madd_epi16(packed, _mm_set_epi16(1, 10000, 1, 10000, 1, 10000, 1, 10000));
    const uint64_t two = (uint64_t)_mm_cvtsi128_si64(eights);  // low word: digits 1-8 of the field, high word: digits 9-16
    uint64_t frac = (two & 0xFFFFFFFFull) * 100000000ull + (two >> 32);
    for (int i = 16; i < nf; ++i) frac = frac * 10 + (uint64_t)(f[i] - '0');
    w = w * kPow10u[nf] + frac;
    if (w == 0) {
      val = 0.0;
    } else {
      const long double y = (long double)w / kPow10[nf];
      uint64_t m;
      memcpy(&m, &y, 8);
      if ((m & 0x7FFull) == 0x400ull) return false;  // on a double's rounding boundary: the exact converter decides
      val = (double)y;
    }
  }
  uint64_t bits;
  memcpy(&bits, &val, 8);
  bits |= sign;
  memcpy(&val, &bits, 8);
  *out = (float)val;
  return true;
}

// Row text (b, e) -- between '[' and ']' -- into dst[0 .. cols); `limit` is the end of the readable buffer.  false: not `cols`
// well-formed numbers separated by single commas (the caller falls back to the sequential parser for the verdict).
B2S_AVX2 bool parse_row_simd(const char* b, const char* e, const char* limit, float* dst, int64_t cols, std::vector<uint32_t>& marks) {
  const size_t len = (size_t)(e - b);
  if (len >= (1ull << 31)) return false;
  if (marks.size() < (size_t)(2 * cols + 2 + 64)) marks.resize((size_t)(2 * cols + 2 + 64));
  uint32_t* const mk = marks.data();  // token k: [mk[2k], mk[2k + 1])
  const size_t cap = (size_t)(2 * cols + 2);
  size_t nm = 0;
  uint32_t carry = 0;  // was the previous byte a token byte?
  const __m256i comma = _mm256_set1_epi8(','), space = _mm256_set1_epi8(' '), tab = _mm256_set1_epi8('\t'), nl = _mm256_set1_epi8('\n'),
                cr = _mm256_set1_epi8('\r');
  for (size_t off = 0; off < len; off += 32) {
    __m256i v;
    if (off + 32 <= len) {
      v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(b + off));
    } else {  // the row's last bytes, padded with separators
      alignas(32) char tail[32];
      memset(tail, ' ', 32);
      memcpy(tail, b + off, len - off);
      v = _mm256_load_si256(reinterpret_cast<const __m256i*>(tail));
    }
    const __m256i sep = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, comma), _mm256_cmpeq_epi8(v, space)),
                                        _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, tab), _mm256_cmpeq_epi8(v, nl)), _mm256_cmpeq_epi8(v, cr)));
    const uint32_t tok = ~(uint32_t)_mm256_movemask_epi8(sep);
    const uint32_t prev = (tok << 1) | carry;
    uint32_t edges = tok ^ prev;  // bit i: byte i starts a token (tok) or is the first byte after one (!tok)
    carry = tok >> 31;
    if (nm + (size_t)_mm_popcnt_u32(edges) > cap) return false;  // more tokens than columns
    while (edges) {
      mk[nm++] = (uint32_t)off + _tzcnt_u32(edges);
      edges = _blsr_u32(edges);
    }
  }
  if (carry) mk[nm++] = (uint32_t)len;  // the last token ends with the row  (the padded tail cannot carry; a full last chunk can)
  if (nm != (size_t)(2 * cols)) return false;
  // separators: nothing but white space in front of the first token and behind the last, exactly one comma between neighbours
  uint32_t prev_end = 0;
  for (int64_t k = 0; k < cols; ++k) {
    const uint32_t s0 = mk[2 * k], e0 = mk[2 * k + 1];
    const uint32_t gap = s0 - prev_end;
    int commas = 0;
    if (gap == 2) commas = (b[prev_end] == ',') + (b[prev_end + 1] == ',');
    else if (gap == 1) commas = b[prev_end] == ',';
    else
      for (uint32_t i = prev_end; i < s0; ++i) commas += b[i] == ',';
    if (commas != (k ? 1 : 0)) return false;
    prev_end = e0;
    const char* ts = b + s0;
    const char* te = b + e0;
    if (ts + 40 <= limit && convert_plain_decimal(ts, te, dst + k)) continue;
    Cur c{ts, te};
    if (!parse_number(c, dst + k) || c.p != te) return false;
  }
  for (uint32_t i = prev_end; i < (uint32_t)len; ++i)
    if (b[i] == ',') return false;
  return true;
}
#endif

// ---- worker pool of the parallel parse ---------------------------------------------------------------------------------
// Threads are created once and parked on a condition variable: a body of a few MB is parsed in about a millisecond per
// thread, less than it takes freshly created threads to be spread over the cores.  One job at a time: a caller that finds the
// pool taken parses its body on its own thread.  After fork() the child starts without a pool and builds its own on first use.
struct Pool {
  std::mutex caller;  // held for the duration of a job
  std::mutex m;
  std::condition_variable cv_job, cv_done;
  uint64_t generation = 0;
  int pending = 0;
  void (*fn)(void*) = nullptr;
  void* arg = nullptr;
  int workers = 0;
};
std::atomic<Pool*> g_pool{nullptr};
std::mutex g_pool_make;

void pool_worker(Pool* pl) {
  uint64_t seen = 0;
  for (;;) {
    void (*fn)(void*);
    void* arg;
    {
      std::unique_lock<std::mutex> lk(pl->m);
      pl->cv_job.wait(lk, [&] { return pl->generation != seen; });
      seen = pl->generation;
      fn = pl->fn;
      arg = pl->arg;
    }
    fn(arg);
    {
      std::lock_guard<std::mutex> lk(pl->m);
      if (--pl->pending == 0) pl->cv_done.notify_one();
    }
  }
}

Pool* get_pool(int threads) {
  Pool* pl = g_pool.load(std::memory_order_acquire);
  if (pl) return pl;
  std::lock_guard<std::mutex> lk(g_pool_make);
  pl = g_pool.load(std::memory_order_acquire);
  if (pl) return pl;
  static const bool hooked = [] {
    // fork: the child has no workers, so it starts without a pool; the pool-creation lock is held across the fork so that the
    // child never inherits it locked by a thread that does not exist there
    pthread_atfork([] { g_pool_make.lock(); }, [] { g_pool_make.unlock(); },
                   [] {
                     g_pool.store(nullptr, std::memory_order_release);
                     g_pool_make.unlock();
                   });
    return true;
  }();
  (void)hooked;
  pl = new Pool;  // never destroyed: its threads are parked for the life of the process
  try {
    for (int t = 1; t < threads; ++t) {
      std::thread(pool_worker, pl).detach();
      ++pl->workers;
    }
  } catch (const std::exception&) {  // no more threads to be had: the job runs on those that started (and the caller)
  }
  g_pool.store(pl, std::memory_order_release);
  return pl;
}

// runs `work` on the caller's thread and on every pool thread at once; false (nothing ran) when another job holds the pool
template <class F>
bool run_on_pool(int threads, F& work) {
  Pool* pl = get_pool(threads);
  std::unique_lock<std::mutex> job(pl->caller, std::try_to_lock);
  if (!job.owns_lock()) return false;
  {
    std::lock_guard<std::mutex> lk(pl->m);
    pl->fn = [](void* a) { (*static_cast<F*>(a))(); };
    pl->arg = &work;
    pl->pending = pl->workers;
    ++pl->generation;
  }
  pl->cv_job.notify_all();
  work();
  std::unique_lock<std::mutex> lk(pl->m);
  pl->cv_done.wait(lk, [&] { return pl->pending == 0; });
  return true;
}

// Optimistic parallel parse of a large numeric matrix.  `first` points at the '[' of row 0.  Inside a numeric matrix every
// ']' closes a row (or the matrix), so one memchr pass finds the rows; worker threads then parse disjoint row ranges with
// the same parse_number as the sequential path.  Anything unexpected -- a string, a nested list, a ragged row, a row that
// does not end where the scan said -- makes this return false WITHOUT a verdict: the caller re-parses sequentially, which
// also produces the right error.  On success: rows / cols / n are set and `end` is just past the matrix's closing ']'.
bool parse_matrix_parallel(const char* first, const char* limit, float* out, int64_t out_cap, int64_t* rows_out, int64_t* cols_out,
                           const char** end) {
  // several threads for bodies over 256 KB; smaller ones (and B2S_CODEC_THREADS=1) take the same row parser on this thread
  const int threads = limit - first < (1 << 18) ? 1 : codec_threads();
#if B2S_CODEC_FAST
  const bool simd = cpu_has_avx2();
#else
  const bool simd = false;
#endif
  if (threads < 2 && !simd) return false;  // nothing to gain over the sequential parser
  auto ws = [&](const char* p) {
    while (p < limit && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
    return p;
  };
  std::vector<const char*> begin, stop;  // row r is the text (begin[r], stop[r]): after its '[' up to its ']'
  const char* p = first;
  for (;;) {
    if (p >= limit || *p != '[') return false;
    const char* close = static_cast<const char*>(memchr(p + 1, ']', (size_t)(limit - p - 1)));
    if (!close) return false;
    begin.push_back(p + 1);
    stop.push_back(close);
    p = ws(close + 1);
    if (p >= limit) return false;
    if (*p == ',') {
      p = ws(p + 1);
      continue;
    }
    if (*p != ']') return false;
    ++p;  // the matrix's own ']'
    break;
  }
  const int64_t rows = (int64_t)begin.size();
  const int use_threads = rows < 2 * (int64_t)threads ? 1 : threads;
  // row 0 fixes the width
  int64_t cols = 0;
  {
    Cur c{begin[0], stop[0]};
    c.ws();
    if (c.p < c.end) {
      float scratch;
      for (;;) {
        if (!parse_number(c, &scratch)) return false;
        ++cols;
        if (c.eat(',')) continue;
        break;
      }
      c.ws();
      if (c.p != c.end) return false;
    }
  }
  if (rows * cols > out_cap) return false;  // the sequential path reports it
  // row blocks are handed out through one atomic counter: a worker that is scheduled late simply takes fewer of them
  const int64_t block = std::max<int64_t>(16, rows / (8 * (int64_t)use_threads));
  const int64_t n_blocks = (rows + block - 1) / block;
  std::atomic<int64_t> next{0};
  std::atomic<int> bad{0};
  auto work = [&] {
    std::vector<uint32_t> marks;  // token boundaries of one row (stage 1 of the SIMD row parser)
    for (;;) {
      const int64_t b = next.fetch_add(1, std::memory_order_relaxed);
      if (b >= n_blocks || bad.load(std::memory_order_relaxed)) return;
      const int64_t r0 = b * block, r1 = std::min(rows, r0 + block);
      for (int64_t r = r0; r < r1; ++r) {
        float* dst = out + r * cols;
#if B2S_CODEC_FAST
        if (simd) {
          if (parse_row_simd(begin[(size_t)r], stop[(size_t)r], limit, dst, cols, marks)) continue;
          bad.store(1, std::memory_order_relaxed);
          return;
        }
#endif
        Cur c{begin[(size_t)r], stop[(size_t)r]};
        int64_t w = 0;
        c.ws();
        if (c.p < c.end) {
          for (;;) {
            float v;
            if (!parse_number(c, &v)) {
              bad.store(1, std::memory_order_relaxed);
              return;
            }
            if (w < cols) dst[w] = v;
            ++w;
            if (c.eat(',')) continue;
            break;
          }
          c.ws();
        }
        if (c.p != c.end || w != cols) {
          bad.store(1, std::memory_order_relaxed);
          return;
        }
      }
    }
  };
  if (use_threads < 2 || !run_on_pool(threads, work)) work();  // (the pool is busy with another caller's body: this thread alone)
  if (bad.load()) return false;
  *rows_out = rows;
  *cols_out = cols;
  *end = p;
  return true;
}

}  // namespace

static int parse_inputs_impl(const char* body, int64_t len, float* out, int64_t out_cap, int64_t* n_rows, int64_t* n_cols,
                             int64_t* value_begin, int64_t* value_end) {
  if (!body || len <= 0 || !out || !n_rows || !n_cols) return b2s_int_fail(B2S_ERR_INVALID, "bad arguments");
  Cur c{body, body + len};
  if (!c.eat('{')) return b2s_int_fail(B2S_ERR_UNSUPPORTED, "body is not a JSON object");
  c.ws();
  if (c.p < c.end && *c.p == '}') return b2s_int_fail(B2S_ERR_UNSUPPORTED, "no \"inputs\" member");
  for (;;) {
    c.ws();
    const char* k = c.p;
    if (!skip_string(c)) return b2s_int_fail(B2S_ERR_INVALID, "malformed JSON at offset %lld", (long long)(c.p - body));
    const bool is_inputs = (c.p - k) == 8 && memcmp(k, "\"inputs\"", 8) == 0;
    if (!c.eat(':')) return b2s_int_fail(B2S_ERR_INVALID, "malformed JSON at offset %lld", (long long)(c.p - body));
    if (is_inputs) {
      c.ws();
      const char* v0 = c.p;
      if (!c.eat('[')) return b2s_int_fail(B2S_ERR_UNSUPPORTED, "\"inputs\" is not a list");
      int64_t rows = 0, cols = -1, n = 0;
      c.ws();
      if (c.p < c.end && *c.p == ']') {
        ++c.p;
        cols = 0;
      } else {
        c.ws();
        const bool nested = c.p < c.end && *c.p == '[';
        if (nested) {  // large matrices: rows parsed by several threads (same values; falls through when unsure)
          const char* after = nullptr;
          if (parse_matrix_parallel(c.p, c.end, out, out_cap, &rows, &cols, &after)) {
            *n_rows = rows;
            *n_cols = cols;
            if (value_begin) *value_begin = v0 - body;
            if (value_end) *value_end = after - body;
            return B2S_OK;
          }
          rows = 0;
          cols = -1;
        }
        for (;;) {
          if (nested) {
            if (!c.eat('[')) return b2s_int_fail(B2S_ERR_UNSUPPORTED, "\"inputs\" mixes rows and scalars");
            int64_t w = 0;
            c.ws();
            if (!(c.p < c.end && *c.p == ']')) {
              for (;;) {
                if (n >= out_cap) return b2s_int_fail(B2S_ERR_INVALID, "inputs do not fit the %lld-float buffer", (long long)out_cap);
                if (!parse_number(c, out + n)) return b2s_int_fail(B2S_ERR_UNSUPPORTED, "non-numeric input at offset %lld", (long long)(c.p - body));
                ++n;
                ++w;
                if (c.eat(',')) continue;
                break;
              }
            }
            if (!c.eat(']')) return b2s_int_fail(B2S_ERR_INVALID, "malformed JSON at offset %lld", (long long)(c.p - body));
            if (cols < 0) cols = w;
            if (w != cols) return b2s_int_fail(B2S_ERR_UNSUPPORTED, "ragged \"inputs\" rows (%lld vs %lld)", (long long)w, (long long)cols);
          } else {  // a flat list: one scalar per event
            if (n >= out_cap) return b2s_int_fail(B2S_ERR_INVALID, "inputs do not fit the %lld-float buffer", (long long)out_cap);
            if (!parse_number(c, out + n)) return b2s_int_fail(B2S_ERR_UNSUPPORTED, "non-numeric input at offset %lld", (long long)(c.p - body));
            ++n;
            cols = 1;
          }
          ++rows;
          if (c.eat(',')) continue;
          break;
        }
        if (!c.eat(']')) return b2s_int_fail(B2S_ERR_INVALID, "malformed JSON at offset %lld", (long long)(c.p - body));
      }
      *n_rows = rows;
      *n_cols = cols < 0 ? 0 : cols;
      if (value_begin) *value_begin = v0 - body;
      if (value_end) *value_end = c.p - body;
      return B2S_OK;
    }
    if (!skip_value(c)) return b2s_int_fail(B2S_ERR_INVALID, "malformed JSON at offset %lld", (long long)(c.p - body));
    if (c.eat(',')) continue;
    if (c.eat('}')) return b2s_int_fail(B2S_ERR_UNSUPPORTED, "no \"inputs\" member");
    return b2s_int_fail(B2S_ERR_INVALID, "malformed JSON at offset %lld", (long long)(c.p - body));
  }
}

extern "C" int b2s_json_parse_inputs(const char* body, int64_t len, float* out, int64_t out_cap, int64_t* n_rows, int64_t* n_cols,
                                     int64_t* value_begin, int64_t* value_end) {
  try {  // no C++ exception (an allocation failure on a huge body) crosses the C boundary
    return parse_inputs_impl(body, len, out, out_cap, n_rows, n_cols, value_begin, value_end);
  } catch (const std::exception& e) {
    return b2s_int_fail(B2S_ERR_INVALID, "body parser failed: %s", e.what());
  }
}

extern "C" int b2s_json_format_outputs(const void* vals, int32_t is_int, int64_t n_rows, int64_t n_cols, int32_t flat, char* out,
                                       int64_t out_cap, int64_t* out_len) {
  if (!vals || !out || !out_len || n_rows < 0 || n_cols < 0) return b2s_int_fail(B2S_ERR_INVALID, "bad arguments");
  if (flat && n_cols != 1) return b2s_int_fail(B2S_ERR_INVALID, "a flat list needs one value per row");
  char* o = out;
  char* const end = out + out_cap;
  const float* f = static_cast<const float*>(vals);
  const int32_t* iv = static_cast<const int32_t*>(vals);
  auto room = [&](int n) { return o + n <= end; };
  if (!room(2)) return b2s_int_fail(B2S_ERR_INVALID, "output buffer too small");
  *o++ = '[';
  for (int64_t r = 0; r < n_rows; ++r) {
    if (r) {
      if (!room(2)) return b2s_int_fail(B2S_ERR_INVALID, "output buffer too small");
      *o++ = ',';
      *o++ = ' ';
    }
    if (!flat) {
      if (!room(1)) return b2s_int_fail(B2S_ERR_INVALID, "output buffer too small");
      *o++ = '[';
    }
    for (int64_t k = 0; k < n_cols; ++k) {
      if (!room(40)) return b2s_int_fail(B2S_ERR_INVALID, "output buffer too small");
      if (k) {
        *o++ = ',';
        *o++ = ' ';
      }
      if (is_int) o = std::to_chars(o, o + 16, iv[r * n_cols + k]).ptr;
      else o += repr_double((double)f[r * n_cols + k], o);
    }
    if (!flat) {
      if (!room(1)) return b2s_int_fail(B2S_ERR_INVALID, "output buffer too small");
      *o++ = ']';
    }
  }
  if (!room(1)) return b2s_int_fail(B2S_ERR_INVALID, "output buffer too small");
  *o++ = ']';
  *out_len = o - out;
  return B2S_OK;
}

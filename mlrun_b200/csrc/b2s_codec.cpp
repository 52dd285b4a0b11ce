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
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <limits>
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

// one JSON number (json.loads grammar, plus the NaN / Infinity / -Infinity literals it accepts, plus null -> NaN)
bool parse_number(Cur& c, float* out) {
  c.ws();
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
  const int exp10 = atoi(s + 1);  // after 'e'
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
    o += snprintf(o, 8, "%d", a);
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

// Optimistic parallel parse of a large numeric matrix.  `first` points at the '[' of row 0.  Inside a numeric matrix every
// ']' closes a row (or the matrix), so one memchr pass finds the rows; worker threads then parse disjoint row ranges with
// the same parse_number as the sequential path.  Anything unexpected -- a string, a nested list, a ragged row, a row that
// does not end where the scan said -- makes this return false WITHOUT a verdict: the caller re-parses sequentially, which
// also produces the right error.  On success: rows / cols / n are set and `end` is just past the matrix's closing ']'.
bool parse_matrix_parallel(const char* first, const char* limit, float* out, int64_t out_cap, int64_t* rows_out, int64_t* cols_out,
                           const char** end) {
  const int threads = codec_threads();
  if (threads < 2 || limit - first < (1 << 18)) return false;
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
  if (rows < 2 * threads) return false;
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
  std::vector<char> ok((size_t)threads, 1);
  auto work = [&](int t) {
    const int64_t r0 = rows * t / threads, r1 = rows * (t + 1) / threads;
    for (int64_t r = r0; r < r1; ++r) {
      Cur c{begin[(size_t)r], stop[(size_t)r]};
      float* dst = out + r * cols;
      int64_t w = 0;
      c.ws();
      if (c.p < c.end) {
        for (;;) {
          float v;
          if (!parse_number(c, &v)) {
            ok[(size_t)t] = 0;
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
        ok[(size_t)t] = 0;
        return;
      }
    }
  };
  std::vector<std::thread> pool;
  int started = 1;  // range 0 runs on this thread
  try {
    for (int t = 1; t < threads; ++t) {
      pool.emplace_back(work, t);
      ++started;
    }
  } catch (const std::exception&) {  // no more threads to be had: the ranges that got none are parsed here
  }
  work(0);
  for (int t = started; t < threads; ++t) work(t);
  for (auto& th : pool) th.join();
  for (char good : ok)
    if (!good) return false;
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
      if (is_int) o += snprintf(o, 16, "%d", iv[r * n_cols + k]);
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

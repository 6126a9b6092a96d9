// csvsink.cpp — the data format on the far side of the hot path: physim's `csvsink` renderer
// (utilities/src/csvsink.rs:44-80), the headless sink a CPU run of the README pipeline uses.
// Every printed state is ONE line: "x,y,z," per entity, then '\n'; the first state received (the
// initial one, pipeline.rs:129-131) is always printed, state k >= 1 iff k % print_n == 0
// (csvsink.rs:60-70).  Numbers are Rust's `{}` for f64: the shortest digits that round-trip, in
// positional notation (never an exponent), "NaN", "inf", "-inf".
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "host_pool.hpp"

namespace {

struct CsvSink {
  std::FILE* f = nullptr;
  size_t print_n = 1, iteration = 0;
};

// Rust `format!("{}", v)`; returns the number of characters written (cap >= 400 always fits)
size_t format_f64(double v, char* out, size_t cap) {
  if (std::isnan(v)) {
    const size_t k = cap < 3 ? cap : 3;
    std::memcpy(out, "NaN", k);
    return k;
  }
  if (std::isinf(v)) {
    const char* s = v < 0 ? "-inf" : "inf";
    const size_t k = std::min(cap, std::strlen(s));
    std::memcpy(out, s, k);
    return k;
  }
  // shortest round-trip digits (d.ddd e±xx), then written out positionally with zero padding, as
  // Rust does (f64::MAX prints as 17976931348623157 followed by 292 zeros, not its exact integer)
  char sci[64];
  const std::to_chars_result r = std::to_chars(sci, sci + sizeof sci, v, std::chars_format::scientific);
  if (r.ec != std::errc()) return 0;
  const char* p = sci;
  size_t k = 0;
  auto put = [&](char c) {
    if (k < cap) out[k] = c;
    ++k;
  };
  if (*p == '-') {
    put('-');
    ++p;
  }
  char digits[32];
  int nd = 0;
  for (; p < r.ptr && *p != 'e'; ++p)
    if (*p != '.') digits[nd++] = *p;
  int exp10 = 0;
  if (p < r.ptr && *p == 'e') {
    ++p;
    const bool neg = *p == '-';
    if (*p == '-' || *p == '+') ++p;
    for (; p < r.ptr; ++p) exp10 = exp10 * 10 + (*p - '0');
    if (neg) exp10 = -exp10;
  }
  while (nd > 1 && digits[nd - 1] == '0') --nd;  // (shortest form has none, kept for safety)
  if (nd == 1 && digits[0] == '0') {
    put('0');
    return k <= cap ? k : 0;
  }
  const int point = exp10 + 1;  // digits before the decimal point
  if (point <= 0) {
    put('0');
    put('.');
    for (int i = 0; i < -point; ++i) put('0');
    for (int i = 0; i < nd; ++i) put(digits[i]);
  } else {
    for (int i = 0; i < point; ++i) put(i < nd ? digits[i] : '0');
    if (nd > point) {
      put('.');
      for (int i = point; i < nd; ++i) put(digits[i]);
    }
  }
  return k <= cap ? k : 0;
}

}  // namespace

extern "C" {

size_t pb200_csv_format_f64(double v, char* buf, size_t cap) { return format_f64(v, buf, cap); }

void* pb200_csvsink_create(const char* file, size_t print_n) {
  if (print_n == 0) {  // the reference panics (rem_euclid by zero)
    pb200::set_error("csvsink: print_n must be >= 1");
    return nullptr;
  }
  // csvsink.rs:45-57: create + truncate; default name csvsink.csv (csvsink.rs:31-34)
  std::FILE* f = std::fopen(file && *file ? file : "csvsink.csv", "wb");
  if (!f) {
    pb200::set_error("Error opening %s: %s", file ? file : "csvsink.csv", std::strerror(errno));
    return nullptr;
  }
  CsvSink* s = new CsvSink();
  s->f = f;
  s->print_n = print_n;
  return s;
}

int pb200_csvsink_push(void* sink, const Entity* state, size_t n) {
  if (!sink || (n && !state)) return -1;
  CsvSink& s = *static_cast<CsvSink*>(sink);
  const size_t k = s.iteration++;
  if (k != 0 && k % s.print_n != 0) return 0;
  // format in parallel into per-chunk strings, write them in order
  const int parts = std::max(1, std::min<int>(pb200::HostPool::instance().threads(), int(n / 4096) + 1));
  std::vector<std::string> chunks(parts);
  const size_t per = (n + parts - 1) / parts;
  pb200::HostPool::instance().parallel_for(size_t(parts), 1, [&](size_t b, size_t e) {
    char num[512];
    for (size_t c = b; c < e; ++c) {
      std::string& out = chunks[c];
      const size_t i0 = std::min(n, c * per), i1 = std::min(n, i0 + per);
      out.reserve((i1 - i0) * 60);
      for (size_t i = i0; i < i1; ++i) {
        const double v[3] = {state[i].x, state[i].y, state[i].z};
        for (int a = 0; a < 3; ++a) {
          out.append(num, format_f64(v[a], num, sizeof num));
          out.push_back(',');
        }
      }
    }
  });
  for (const std::string& c : chunks)
    if (!c.empty() && std::fwrite(c.data(), 1, c.size(), s.f) != c.size()) return -1;
  if (std::fputc('\n', s.f) == EOF) return -1;
  return 0;
}

int pb200_csvsink_next_is_printed(void* sink) {
  if (!sink) return 0;
  const CsvSink& s = *static_cast<CsvSink*>(sink);
  return (s.iteration == 0 || s.iteration % s.print_n == 0) ? 1 : 0;
}

size_t pb200_csvsink_count(void* sink) { return sink ? static_cast<CsvSink*>(sink)->iteration : 0; }

int pb200_csvsink_skip(void* sink) {
  if (!sink) return -1;
  ++static_cast<CsvSink*>(sink)->iteration;
  return 0;
}

void pb200_csvsink_destroy(void* sink) {
  if (!sink) return;
  CsvSink* s = static_cast<CsvSink*>(sink);
  if (s->f) std::fclose(s->f);
  delete s;
}

}  // extern "C"

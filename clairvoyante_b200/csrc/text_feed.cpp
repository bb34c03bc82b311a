// text_feed.cpp -- native parser for the candidate-tensor text stream (host code, no CUDA).
//
// Replaces the per-row `row.split()` + `np.array(528 strings)` + channel subtract of the reference feed
// (clairvoyante/utils_v2.py:29-47; row format `chrom pos refseq v0 .. v527`, values printed "%0.1f",
// dataPrepScripts/CreateTensor.py:56).  SURVEY.md 8(f) rank 1: once the network runs at 10^7 sites/s the CPython
// tokeniser (10^3..10^4 rows/s) is the wall of callVar.py.
//
// One call parses up to max_lines complete lines of a byte buffer:
//   pass 1 (caller thread)  finds the line boundaries,
//   pass 2 (worker threads) tokenises each line, converts the 528 values, subtracts channel 0 from channels 1..3
//                           (utils_v2.py:46) and writes the row at its line index,
//   pass 3 (caller thread)  compacts the kept rows to the front of x (drops are rare).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/cvb200.h"

void cvb_internal_set_error(const char* msg);  // cvb200.cu

namespace {

int fail(const char* msg) { cvb_internal_set_error(msg); return 1; }

struct SpaceTable {
  uint8_t t[256];
  constexpr SpaceTable() : t() {
    for (int i = 0; i < 256; ++i) t[i] = 0;
    t[(int)' '] = t[(int)'\t'] = t[(int)'\r'] = t[(int)'\n'] = t[(int)'\v'] = t[(int)'\f'] = 1;
    t[0x1c] = t[0x1d] = t[0x1e] = t[0x1f] = 1;  // str.split() also breaks on the ASCII separators
  }
};
constexpr SpaceTable kSpace;
inline bool is_space(char c) { return kSpace.t[(uint8_t)c] != 0; }

// any token NumPy's string -> float32 conversion accepts (exponents, inf, nan ...): via double like NumPy
bool parse_value_slow(const char* p, const char* e, float* out) {
  char tmp[64];
  const size_t n = (size_t)(e - p);
  if (n == 0 || n >= sizeof(tmp)) return false;
  memcpy(tmp, p, n);
  tmp[n] = 0;
  char* endp = nullptr;
  const double v = strtod(tmp, &endp);
  if (endp != tmp + n) return false;
  *out = (float)v;
  return true;
}

// One field starting at p (not a blank): returns the end of the token; *ok = false if it is not a number.
// Fast path "[+-]digits[.digits]" with at most 15 significant digits: the integer numerator and the power of ten are
// exact doubles, so the single division rounds like strtod; "%0.1f" rows of counts (".0") skip the division altogether.
inline const char* parse_field(const char* p, const char* e, float* out, bool* ok) {
  const char* q = p;
  bool neg = false;
  if (*q == '-' || *q == '+') { neg = *q == '-'; ++q; }
  uint64_t ip = 0, fp = 0;
  int nd = 0, nf = 0;
  while (q < e && (unsigned)(*q - '0') < 10u && nd < 15) { ip = ip * 10 + (uint64_t)(*q - '0'); ++q; ++nd; }
  if (q < e && *q == '.') {
    ++q;
    while (q < e && (unsigned)(*q - '0') < 10u && nd + nf < 15) { fp = fp * 10 + (uint64_t)(*q - '0'); ++q; ++nf; }
  }
  if ((q == e || is_space(*q)) && nd + nf > 0) {
    static const double p10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    const double v = fp == 0 ? (double)ip : ((double)ip * p10[nf] + (double)fp) / p10[nf];
    *out = (float)(neg ? -v : v);
    return q;
  }
  while (q < e && !is_space(*q)) ++q;
  if (!parse_value_slow(p, q, out)) *ok = false;
  return q;
}

struct LineMeta {  // mirrors the int64[10] record documented in cvb200.h
  int64_t status, line_off, line_len, chrom_off, chrom_len, pos_off, pos_len, seq_off, seq_len, reserved;
};

void parse_line(const char* buf, LineMeta* m, float* row) {
  const char* p = buf + m->line_off;
  const char* e = p + m->line_len;
  const char* tokb[3];
  const char* toke[3];
  int nt = 0;
  bool numbers_ok = true;
  while (p < e) {
    while (p < e && is_space(*p)) ++p;
    if (p >= e) break;
    if (nt < 3) {
      tokb[nt] = p;
      while (p < e && !is_space(*p)) ++p;
      toke[nt] = p;
    } else if (nt < 531) {
      p = parse_field(p, e, row + (nt - 3), &numbers_ok);
    } else {
      while (p < e && !is_space(*p)) ++p;  // too many fields
    }
    ++nt;
  }
  const bool bad_number = !numbers_ok;
  if (nt == 0) { m->status = CVB_LINE_BLANK; return; }
  if (nt != 531) { m->status = CVB_LINE_MALFORMED; return; }
  m->chrom_off = tokb[0] - buf; m->chrom_len = toke[0] - tokb[0];
  m->pos_off = tokb[1] - buf;   m->pos_len = toke[1] - tokb[1];
  m->seq_off = tokb[2] - buf;   m->seq_len = toke[2] - tokb[2];
  if (m->seq_len <= 16) { m->status = CVB_LINE_MALFORMED; return; }
  const char c = (char)(tokb[2][16] & ~0x20);  // ASCII upper-case of the centre reference base (utils_v2.py:38-39)
  if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) { m->status = CVB_LINE_SKIPPED; return; }
  if (bad_number) { m->status = CVB_LINE_MALFORMED; return; }
  for (int i = 0; i < 528; i += 4) {  // utils_v2.py:46: channels 1..3 relative to the reference channel
    const float r = row[i];
    row[i + 1] -= r; row[i + 2] -= r; row[i + 3] -= r;
  }
  m->status = CVB_LINE_KEPT;
}

}  // namespace

extern "C" int cvb_parse_tensor_text(const char* buf, int64_t len, int final_chunk, int64_t max_lines, int threads, float* x,
                                     int64_t* meta, int64_t* lines, int64_t* kept, int64_t* consumed) {
  if (!lines || !kept || !consumed) return fail("cvb_parse_tensor_text: NULL result pointer");
  *lines = 0; *kept = 0; *consumed = 0;
  if (len < 0 || max_lines < 0) return fail("cvb_parse_tensor_text: negative size");
  if (len == 0 || max_lines == 0) return 0;
  if (!buf || !x || !meta) return fail("cvb_parse_tensor_text: NULL buffer");
  LineMeta* M = reinterpret_cast<LineMeta*>(meta);
  int64_t n = 0, off = 0;
  while (n < max_lines && off < len) {
    const char* nl = static_cast<const char*>(memchr(buf + off, '\n', (size_t)(len - off)));
    int64_t end;
    if (nl) end = nl - buf;
    else if (final_chunk) end = len;  // last line of the stream without a newline
    else break;                       // incomplete line: the caller re-submits it with more data
    memset(&M[n], 0, sizeof(LineMeta));
    M[n].line_off = off;
    M[n].line_len = end - off;
    ++n;
    off = nl ? end + 1 : end;
  }
  *lines = n;
  *consumed = off;
  if (n == 0) return 0;
  int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(T, 64), n / 64));  // >= 64 lines (~170 KB) per thread
  auto work = [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) parse_line(buf, &M[i], x + i * 528);
  };
  if (T == 1) {
    work(0, n);
  } else {
    std::vector<std::thread> pool;
    pool.reserve((size_t)T - 1);
    const int64_t per = (n + T - 1) / T;
    for (int t = 1; t < T; ++t) pool.emplace_back(work, std::min(n, t * per), std::min(n, (t + 1) * per));
    work(0, std::min(n, per));
    for (auto& th : pool) th.join();
  }
  int64_t k = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (M[i].status != CVB_LINE_KEPT) continue;
    if (k != i) memcpy(x + k * 528, x + i * 528, 528 * sizeof(float));
    ++k;
  }
  *kept = k;
  return 0;
}

// "chrom:pos:SEQ\n" (SEQ upper-cased, utils_v2.py:38,42) for every KEPT line of a cvb_parse_tensor_text result: the position
// strings GetTensor yields, assembled here because building them row by row in Python was the slowest part of the feed.
extern "C" int64_t cvb_tensor_text_positions(const char* buf, const int64_t* meta, int64_t lines, char* out, int64_t cap) {
  if (lines < 0 || (lines > 0 && (!buf || !meta || !out))) {
    cvb_internal_set_error("cvb_tensor_text_positions: bad argument");
    return -1;
  }
  const LineMeta* M = reinterpret_cast<const LineMeta*>(meta);
  char* o = out;
  char* const oe = out + cap;
  for (int64_t i = 0; i < lines; ++i) {
    const LineMeta& m = M[i];
    if (m.status != CVB_LINE_KEPT) continue;
    if (m.chrom_len + m.pos_len + m.seq_len + 3 > oe - o) {
      cvb_internal_set_error("cvb_tensor_text_positions: output buffer too small");
      return -1;
    }
    memcpy(o, buf + m.chrom_off, (size_t)m.chrom_len);
    o += m.chrom_len;
    *o++ = ':';
    memcpy(o, buf + m.pos_off, (size_t)m.pos_len);
    o += m.pos_len;
    *o++ = ':';
    const char* s = buf + m.seq_off;
    for (int64_t k = 0; k < m.seq_len; ++k) {
      const char ch = s[k];
      *o++ = (ch >= 'a' && ch <= 'z') ? (char)(ch - 32) : ch;
    }
    *o++ = '\n';
  }
  return o - out;
}

// Narrow feed (cvb_predict_host_counts_i16 / _u8): fp32 candidate tensors -> the RAW counts they were built from, as int16
// (and, when the caller asks, uint8).  x holds n_pos (row, base) positions of 4 channels; subtracted != 0 means channels
// 1..3 are already relative to channel 0 (what GetTensor yields, utils_v2.py:46) and the raw count is x_i + x_0.  The
// conversion is only valid for tensors of non-negative integers: *exact = 0 if any value is not an integer in [0, 32767]
// (the caller then keeps the fp32 feed), *max_count = the largest count seen (<= 255: the uint8 buffer is usable).
extern "C" int cvb_pack_counts(const float* x, int64_t n_pos, int subtracted, int threads, int16_t* out_i16, uint8_t* out_u8,
                               int* max_count, int* exact) {
  if (n_pos < 0 || (n_pos > 0 && (!x || !out_i16))) return fail("cvb_pack_counts: bad argument");
  int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(T, 64), n_pos / 65536));
  std::vector<int> tmax((size_t)T, 0), tok((size_t)T, 1);
  auto work = [&](int t, int64_t a, int64_t b) {
    int mx = 0, ok = 1;
    for (int64_t i = a; i < b; ++i) {
      const float* p = x + 4 * i;
      const float c0 = p[0];
      const float v[4] = {c0, subtracted ? p[1] + c0 : p[1], subtracted ? p[2] + c0 : p[2], subtracted ? p[3] + c0 : p[3]};
      for (int k = 0; k < 4; ++k) {
        const float f = v[k];
        const int c = (f >= 0.f && f <= 32767.f) ? (int)f : -1;
        if (c < 0 || (float)c != f) { ok = 0; out_i16[4 * i + k] = 0; if (out_u8) out_u8[4 * i + k] = 0; continue; }
        mx = c > mx ? c : mx;
        out_i16[4 * i + k] = (int16_t)c;
        if (out_u8) out_u8[4 * i + k] = (uint8_t)(c > 255 ? 255 : c);
      }
    }
    tmax[(size_t)t] = mx;
    tok[(size_t)t] = ok;
  };
  if (T == 1) {
    work(0, 0, n_pos);
  } else {
    std::vector<std::thread> pool;
    const int64_t per = (n_pos + T - 1) / T;
    for (int t = 1; t < T; ++t) pool.emplace_back(work, t, std::min(n_pos, t * per), std::min(n_pos, (t + 1) * per));
    work(0, 0, std::min(n_pos, per));
    for (auto& th : pool) th.join();
  }
  int mx = 0, ok = 1;
  for (int t = 0; t < T; ++t) { mx = std::max(mx, tmax[(size_t)t]); ok &= tok[(size_t)t]; }
  if (max_count) *max_count = mx;
  if (exact) *exact = ok;
  return 0;
}

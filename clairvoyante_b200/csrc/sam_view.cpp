// sam_view.cpp -- what `samtools view -F <mask> <file> ctg[:start-end]` keeps, applied to SAM TEXT (host code, no CUDA).
//
// Both alignment stages of the reference read their records from that command (dataPrepScripts/CreateTensor.py:134-136,
// ExtractVariantCandidates.py:107-109: `-F 2308` drops unmapped, secondary and supplementary records; the region limits
// the records to those overlapping it).  When samtools is not installed the package's command lines accept a .sam / .sam.gz
// text file instead; this filter gives that path the same record set: header lines dropped, other contigs dropped, flag
// mask applied, 1-based inclusive overlap of [POS, POS + reference span - 1] with [start, end] (reference span = sum of
// M / D / N / = / X lengths, at least 1).  Rows with fewer than six fields or non-numeric FLAG / POS are passed on (the
// stages count them as malformed).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/cvb200.h"

void cvb_internal_set_error(const char* msg);  // cvb200.cu

namespace {

inline const char* next_tab(const char* p, const char* e) {
  const char* t = static_cast<const char*>(memchr(p, '\t', (size_t)(e - p)));
  return t ? t : e;
}

// decimal field [p, e): returns false when it is empty or holds anything but digits
inline bool parse_u(const char* p, const char* e, int64_t* v) {
  if (p >= e) return false;
  int64_t x = 0;
  for (; p < e; ++p) {
    if (*p < '0' || *p > '9') return false;
    x = x * 10 + (*p - '0');
    if (x > (int64_t)1 << 60) return false;
  }
  *v = x;
  return true;
}

bool keep_line(const char* p, const char* e, const char* ctg, size_t ctg_len, int flag_mask, int64_t start, int64_t end) {
  if (p < e && *p == '@') return false;
  const char* f[6];
  const char* g[6];
  const char* q = p;
  int nf = 0;
  while (nf < 6 && q <= e) {
    const char* t = next_tab(q, e);
    f[nf] = q;
    g[nf] = t;
    ++nf;
    if (t >= e) break;
    q = t + 1;
  }
  if (nf < 6) {  // malformed (or blank): blank lines are dropped, anything else goes on to the stage
    for (const char* c = p; c < e; ++c)
      if (*c != ' ' && *c != '\t' && *c != '\r') return true;
    return false;
  }
  int64_t flag = 0, pos = 0;
  if (!parse_u(f[1], g[1], &flag)) return true;
  if (flag & flag_mask) return false;
  if ((size_t)(g[2] - f[2]) != ctg_len || memcmp(f[2], ctg, ctg_len) != 0) return false;
  if (start < 0) return true;
  if (!parse_u(f[3], g[3], &pos)) return true;
  if (pos > end) return false;
  if (pos >= start) return true;
  int64_t span = 0, n = 0;
  bool have = false;
  for (const char* c = f[5]; c < g[5]; ++c) {
    if (*c >= '0' && *c <= '9') {
      n = n * 10 + (*c - '0');
      have = true;
      if (n > (int64_t)1 << 40) n = (int64_t)1 << 40;
    } else {
      if (have && (*c == 'M' || *c == 'D' || *c == 'N' || *c == '=' || *c == 'X')) span += n;
      n = 0;
      have = false;
    }
  }
  if (span < 1) span = 1;
  return pos + span - 1 >= start;
}

}  // namespace

extern "C" int cvb_sam_view(const char* in, int64_t len, int final_chunk, const char* ctg_name, int flag_mask, int64_t start,
                            int64_t end, char* out, int64_t* out_len, int64_t* consumed) {
  if (!out_len || !consumed || len < 0 || !ctg_name || (len > 0 && (!in || !out))) {
    cvb_internal_set_error("cvb_sam_view: bad argument");
    return 1;
  }
  const size_t ctg_len = strlen(ctg_name);
  // filter [p, e) (whole lines, except possibly the last) into o; returns the end of what was written
  auto run = [&](const char* p, const char* e, bool last_may_lack_newline, char* o, const char** stopped) -> char* {
    while (p < e) {
      const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(e - p)));
      if (!nl && !last_may_lack_newline) break;
      const char* le = nl ? nl : e;
      if (keep_line(p, le, ctg_name, ctg_len, flag_mask, start, end)) {
        memcpy(o, p, (size_t)(le - p));
        o += le - p;
        *o++ = '\n';  // (out needs len + 1 bytes when the final line has no newline)
      }
      p = nl ? nl + 1 : e;
    }
    *stopped = p;
    return o;
  };
  const char* const e = in + len;
  int T = (int)std::min<int64_t>(std::min<unsigned>(std::thread::hardware_concurrency(), 8u), len >> 21);  // >= 2 MB per thread
  if (T < 2) {
    const char* stopped = in;
    char* o = run(in, e, final_chunk != 0, out, &stopped);
    *out_len = o - out;
    *consumed = stopped - in;
    return 0;
  }
  // cut at line starts; every part writes to the same offset of `out` it has in `in` (output never exceeds input), then the
  // parts are moved together
  std::vector<const char*> cut((size_t)T + 1);
  cut[0] = in;
  cut[(size_t)T] = e;
  for (int t = 1; t < T; ++t) {
    const char* guess = in + len * t / T;
    if (guess < cut[(size_t)t - 1]) guess = cut[(size_t)t - 1];
    const char* nl = static_cast<const char*>(memchr(guess, '\n', (size_t)(e - guess)));
    cut[(size_t)t] = nl ? nl + 1 : e;
  }
  while (T > 1 && cut[(size_t)T - 1] >= e) --T;  // (no line start left for the trailing parts: the part before them reaches the end)
  cut[(size_t)T] = e;
  std::vector<char*> wrote((size_t)T);
  std::vector<const char*> stopped((size_t)T);
  auto work = [&](int t) {
    const bool is_last = t == T - 1;
    wrote[(size_t)t] = run(cut[(size_t)t], cut[(size_t)t + 1], is_last ? final_chunk != 0 : true, out + (cut[(size_t)t] - in),
                           &stopped[(size_t)t]);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  char* o = wrote[0];
  for (int t = 1; t < T; ++t) {
    char* b = out + (cut[(size_t)t] - in);
    const size_t n = (size_t)(wrote[(size_t)t] - b);
    memmove(o, b, n);
    o += n;
  }
  *out_len = o - out;
  *consumed = stopped[(size_t)T - 1] - in;
  return 0;
}

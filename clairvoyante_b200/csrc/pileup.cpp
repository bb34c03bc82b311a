// pileup.cpp -- native alignment pile-up: SAM records + candidate positions + reference sequence -> (33,4,4) count tensors
// (host code, no CUDA).  SURVEY.md 8(f) rank 4: `dataPrepScripts/CreateTensor.py` is what feeds callVar.py in the
// reference's own pipelines (callVarBam.py:61) and, at ~10^3 sites/s in CPython, is their wall once the network runs at
// 10^7 sites/s.
//
// Follows CreateTensor.py step by step (line numbers of /root/reference/dataPrepScripts/CreateTensor.py):
//   :148-181  per SAM row: MAPQ filter, depth cap (`dcov` reads per identical POS), then the CIGAR walk
//   :183-235  CIGAR ops M/=/X, I, D advance (refPos, queryPos); S advances queryPos; N/H/P are ignored (the reference does
//             NOT advance refPos on N -- kept); a candidate centre c is (re)activated for a read while
//             c-17 <= refPos < c+17 (`considerleftedge`, :76-81; otherwise only at refPos == c-17, :74-75) and
//             deactivated after refPos == c+17 (:203-205,232-234); every base of an active centre is one tuple
//   :23-52    GenerateTensor: a tuple contributes only inside -17 <= refPos-c < 16 and only for upper-case A/C/G/T;
//             match: depth++, channels 0 and 2 at the reference base, 1 and 3 at the read base; insertion: channel 1 at
//             min(offset + queryAdv, 32); deletion: channel 2 at the reference base
//   :237-246  a centre is emitted once a read with a NEW start position lies beyond c+17 (and at end of input);
//             :54-57 it is kept only if its window starts inside the fetched reference and depth[16] >= minCoverage
// The reference stores every tuple in Python lists and sums them at flush time; here the sums are accumulated directly.
// Its 10,000,000 outstanding-tuple budget (`availableSlots`, :99,197-199,245) is accounted identically.
// Documented differences: (1) centres flushed together leave in ascending position order (the reference: Python-2 dict
// order); (2) a reference / read index outside the supplied strings counts as 'N' (the reference raises IndexError or,
// for negative indices, wraps around); (3) rows with fewer than 10 fields are counted in `malformed` and skipped.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/cvb200.h"

void cvb_internal_set_error(const char* msg);  // cvb200.cu

namespace {

int fail(const char* msg) { cvb_internal_set_error(msg); return 1; }

constexpr int F = 16, W33 = 2 * F + 1, SITE = W33 * 16;  // param.flankingBaseNum, rows, floats per tensor

inline int base_code(char c) {  // upper-case ACGT only (CreateTensor.py:28-31 tests membership in "ACGT-")
  switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
  }
}

struct Center {
  int64_t slots = 0;   // tuples held (returned to the budget at flush, :245)
  int32_t depth16 = 0; // depth[flankingBaseNum]
  float code[SITE];
  Center() { memset(code, 0, sizeof(code)); }
};

struct Done {
  int64_t center;
  float code[SITE];
};

}  // namespace

struct cvb_pileup {
  std::string ref;            // reference bases of the fetched region
  int64_t ref_off = 0;        // refSeq index = refPos - ref_off  (ref_off = refStart - 1, or 0)
  std::vector<int64_t> cand;  // sorted unique candidate positions (1-based, as in the candidate list)
  int min_mq = 0, dcov = 250, min_cov = 0;
  bool left_edge = true;
  std::map<int64_t, Center> live;  // centerToAln
  std::deque<Done> done;
  int64_t available = 10000000;
  int64_t previous_pos = 0, depth_cap = 0;
  int64_t reads = 0, reads_used = 0, malformed = 0;
  std::string carry;  // incomplete last line of the previous feed

  inline char ref_at(int64_t ref_pos) const {
    const int64_t i = ref_pos - ref_off;
    return (i >= 0 && i < (int64_t)ref.size()) ? ref[(size_t)i] : 'N';
  }
  void emit(int64_t c, const Center& ce) {  // GenerateTensor's tail (:54-59)
    const int64_t new_ref_pos = c - ref_off;
    if (new_ref_pos - (F + 1) >= 0 && ce.depth16 >= min_cov) {
      done.emplace_back();
      done.back().center = c;
      memcpy(done.back().code, ce.code, sizeof(ce.code));
    }
  }
  void flush_before(int64_t pos0) {  // :237-246
    while (!live.empty() && live.begin()->first + (F + 1) < pos0) {
      emit(live.begin()->first, live.begin()->second);
      available += live.begin()->second.slots;
      live.erase(live.begin());
    }
  }
  void flush_all() {  // :248-252 (the reference does not return the slots here either; nothing follows)
    for (auto& kv : live) emit(kv.first, kv.second);
    live.clear();
  }
  void read_line(const char* p, const char* e);
};

namespace {

struct Active {
  int64_t c;
  Center* ce;
};

inline bool is_ws(char ch) { return ch == ' ' || ch == '\t' || ch == '\r' || ch == '\n' || ch == '\v' || ch == '\f'; }

// one tuple (:196-199 and the summation of :25-52)
inline void tuple_match(cvb_pileup* s, const Active& a, int64_t ref_pos, char rb, char qb) {
  if (s->available == 0) return;
  --s->available;
  ++a.ce->slots;
  const int r = base_code(rb), q = base_code(qb);
  if (r < 0 || q < 0) return;
  const int64_t d = ref_pos - a.c;
  if (d < -(F + 1) || d >= F) return;
  const int off = (int)(d + F + 1);
  if (off == F) ++a.ce->depth16;
  float* row = a.ce->code + off * 16;
  row[r * 4 + 0] += 1.f;
  row[q * 4 + 1] += 1.f;
  row[r * 4 + 2] += 1.f;
  row[q * 4 + 3] += 1.f;
}
inline void tuple_ins(cvb_pileup* s, const Active& a, int64_t ref_pos, int64_t query_adv, char qb) {
  if (s->available == 0) return;
  --s->available;
  ++a.ce->slots;
  const int q = base_code(qb);
  if (q < 0) return;
  const int64_t d = ref_pos - a.c;
  if (d < -(F + 1) || d >= F) return;
  int64_t idx = d + F + 1 + query_adv;
  if (idx > 2 * F) idx = 2 * F;
  a.ce->code[idx * 16 + q * 4 + 1] += 1.f;
}
inline void tuple_del(cvb_pileup* s, const Active& a, int64_t ref_pos, char rb) {
  if (s->available == 0) return;
  --s->available;
  ++a.ce->slots;
  const int r = base_code(rb);
  if (r < 0) return;
  const int64_t d = ref_pos - a.c;
  if (d < -(F + 1) || d >= F) return;
  a.ce->code[(d + F + 1) * 16 + r * 4 + 2] += 1.f;
}

}  // namespace

void cvb_pileup::read_line(const char* p, const char* e) {
  // ---- split on whitespace, keep fields 0..9 (:150-160)
  const char* fb[10];
  const char* fe[10];
  int nf = 0;
  while (p < e && nf < 10) {
    while (p < e && is_ws(*p)) ++p;
    if (p >= e) break;
    fb[nf] = p;
    while (p < e && !is_ws(*p)) ++p;
    fe[nf++] = p;
  }
  if (nf == 0) return;          // blank line
  if (fb[0][0] == '@') return;  // header
  ++reads;
  if (nf < 10) { ++malformed; return; }
  const int64_t POS = strtoll(std::string(fb[3], fe[3]).c_str(), nullptr, 10) - 1;
  const long MQ = strtol(std::string(fb[4], fe[4]).c_str(), nullptr, 10);
  const char* cig = fb[5];
  const char* cig_e = fe[5];
  const char* seq = fb[9];
  const int64_t seq_len = fe[9] - fb[9];
  if (MQ < min_mq) return;  // :165-166
  // depth cap (:174-181)
  if (previous_pos != POS) {
    previous_pos = POS;
    depth_cap = 0;
  } else {
    ++depth_cap;
    if (depth_cap >= dcov) return;
  }
  ++reads_used;
  auto query = [&](int64_t i) -> char { return (i >= 0 && i < seq_len) ? seq[i] : 'N'; };

  int64_t ref_pos = POS, query_pos = 0;
  std::vector<Active> active;  // activeSet, ascending centre (centres activate and retire in position order)
  size_t na = 0;               // next candidate this read has not activated yet
  bool na_init = false;
  auto activate = [&](int64_t rp) {  // `if refPos in beginToEnd` (:187-194, :224-231)
    if (!na_init) {              // candidates whose activation window [c-17, c+17) ended at or before rp never activate
      na = (size_t)(std::upper_bound(cand.begin(), cand.end(), rp - (F + 1)) - cand.begin());
      na_init = true;
    }
    while (na < cand.size() && cand[na] - (F + 1) <= rp) {
      const int64_t c = cand[na++];
      if (!left_edge && c - (F + 1) != rp) continue;  // without considerleftedge only the window start activates
      active.push_back(Active{c, &live[c]});
    }
  };
  auto retire = [&](int64_t rp) {  // `if refPos in endToCenter` (:203-205, :232-234)
    for (size_t i = 0; i < active.size(); ++i)
      if (active[i].c + (F + 1) == rp) {
        active.erase(active.begin() + (ptrdiff_t)i);
        break;
      }
  };

  while (cig < cig_e) {  // re.finditer(r"(\d+)([MIDNSHP=X])", CIGAR)
    if (*cig < '0' || *cig > '9') { ++cig; continue; }
    int64_t adv = 0;
    const char* q = cig;
    while (q < cig_e && *q >= '0' && *q <= '9') adv = adv * 10 + (*q++ - '0');
    if (q >= cig_e) break;
    const char op = *q;
    if (!strchr("MIDNSHP=X", op)) { cig = q; continue; }  // digits not followed by an op: the regex would retry later
    cig = q + 1;
    if (available == 0) break;  // :184-185
    if (op == 'S') {
      query_pos += adv;
    } else if (op == 'M' || op == '=' || op == 'X') {
      for (int64_t i = 0; i < adv;) {
        activate(ref_pos);
        if (active.empty()) {
          // nothing open, and every window starting at or before ref_pos has been consumed: nothing can happen before the
          // next candidate's window start -- jump there (or to the end of the run) in one step
          const int64_t next_start = na < cand.size() ? cand[na] - (F + 1) : INT64_MAX;
          const int64_t skip = std::max<int64_t>(1, std::min<int64_t>(adv - i, next_start - ref_pos));
          ref_pos += skip;
          query_pos += skip;
          i += skip;
          continue;
        }
        const char rb = ref_at(ref_pos), qb = query(query_pos);
        for (const Active& a : active) tuple_match(this, a, ref_pos, rb, qb);
        retire(ref_pos);
        ++ref_pos;
        ++query_pos;
        ++i;
      }
    } else if (op == 'I') {
      for (int64_t i = 0; i < adv; ++i) {
        if (!active.empty()) {
          const char qb = query(query_pos);
          for (const Active& a : active) tuple_ins(this, a, ref_pos, i, qb);
        }
        ++query_pos;
      }
    } else if (op == 'D') {
      for (int64_t i = 0; i < adv;) {
        if (active.empty()) {
          activate(ref_pos);  // (a centre opened on a deleted base gets no tuple for it: the reference appends first, :218-223)
          if (active.empty()) {
            const int64_t next_start = na < cand.size() ? cand[na] - (F + 1) : INT64_MAX;
            const int64_t skip = std::max<int64_t>(1, std::min<int64_t>(adv - i, next_start - ref_pos));
            ref_pos += skip;
            i += skip;
          } else {
            ++ref_pos;
            ++i;
          }
          continue;
        }
        const char rb = ref_at(ref_pos);
        for (const Active& a : active) tuple_del(this, a, ref_pos, rb);
        activate(ref_pos);
        retire(ref_pos);
        ++ref_pos;
        ++i;
      }
    }  // N, H, P: no effect (the reference has no branch for them)
  }
  if (depth_cap == 0) flush_before(POS);
}

extern "C" int cvb_pileup_create(const char* ref_seq, int64_t ref_len, int64_t ref_start, const int64_t* cand_pos, int64_t n_cand,
                                 int min_mq, int dcov, int min_coverage, int consider_left_edge, cvb_pileup** out) {
  if (!out || (!ref_seq && ref_len > 0) || (!cand_pos && n_cand > 0) || ref_len < 0 || n_cand < 0)
    return fail("cvb_pileup_create: bad argument");
  cvb_pileup* s = new (std::nothrow) cvb_pileup();
  if (!s) return fail("cvb_pileup_create: out of memory");
  s->ref.assign(ref_seq ? ref_seq : "", (size_t)ref_len);
  s->ref_off = ref_start > 0 ? ref_start - 1 : 0;
  s->cand.assign(cand_pos, cand_pos + n_cand);
  std::sort(s->cand.begin(), s->cand.end());
  s->cand.erase(std::unique(s->cand.begin(), s->cand.end()), s->cand.end());
  s->min_mq = min_mq;
  s->dcov = dcov;
  s->min_cov = min_coverage;
  s->left_edge = consider_left_edge != 0;
  *out = s;
  return 0;
}

extern "C" int cvb_pileup_destroy(cvb_pileup* s) {
  delete s;
  return 0;
}

extern "C" int cvb_pileup_feed(cvb_pileup* s, const char* sam, int64_t len, int final_chunk) {
  if (!s || (!sam && len > 0) || len < 0) return fail("cvb_pileup_feed: bad argument");
  const char* p = sam;
  const char* e = sam + len;
  if (!s->carry.empty()) {  // finish the line started in the previous chunk
    const char* nl = p < e ? (const char*)memchr(p, '\n', (size_t)(e - p)) : nullptr;
    if (!nl && !final_chunk) { s->carry.append(p, (size_t)(e - p)); return 0; }
    const char* stop = nl ? nl : e;
    s->carry.append(p, (size_t)(stop - p));
    s->read_line(s->carry.data(), s->carry.data() + s->carry.size());
    s->carry.clear();
    p = nl ? nl + 1 : e;
  }
  while (p < e) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
    if (!nl) {
      if (final_chunk) s->read_line(p, e);
      else s->carry.assign(p, (size_t)(e - p));
      break;
    }
    s->read_line(p, nl);
    p = nl + 1;
  }
  if (final_chunk) s->flush_all();
  return 0;
}

extern "C" int64_t cvb_pileup_ready(const cvb_pileup* s) { return s ? (int64_t)s->done.size() : 0; }

extern "C" int cvb_pileup_take(cvb_pileup* s, int64_t max_sites, float* x, int64_t* center, int64_t* n_out) {
  if (!s || !n_out || max_sites < 0 || (max_sites > 0 && (!x || !center))) return fail("cvb_pileup_take: bad argument");
  int64_t n = 0;
  while (n < max_sites && !s->done.empty()) {
    memcpy(x + n * SITE, s->done.front().code, sizeof(float) * SITE);
    center[n] = s->done.front().center;
    s->done.pop_front();
    ++n;
  }
  *n_out = n;
  return 0;
}

extern "C" int cvb_pileup_stats(const cvb_pileup* s, int64_t stats[4]) {
  if (!s || !stats) return fail("cvb_pileup_stats: bad argument");
  stats[0] = s->reads;
  stats[1] = s->reads_used;
  stats[2] = s->malformed;
  stats[3] = (int64_t)s->live.size();
  return 0;
}

// ---- the text rows of CreateTensor.py:56: "ctg pos refseq33 v0 .. v527\n" with the values printed "%0.1f"
extern "C" int64_t cvb_pileup_format_rows(const char* ctg, const int64_t* center, const float* x, int64_t n, const char* ref_seq,
                                          int64_t ref_len, int64_t ref_start, char* out, int64_t cap) {
  if (!ctg || n < 0 || (n > 0 && (!center || !x || !out)) || (!ref_seq && ref_len > 0)) {
    cvb_internal_set_error("cvb_pileup_format_rows: bad argument");
    return -1;
  }
  const int64_t ref_off = ref_start > 0 ? ref_start - 1 : 0;
  const size_t lc = strlen(ctg);
  char* p = out;
  char* const end = out + cap;
  for (int64_t i = 0; i < n; ++i) {
    if ((int64_t)(end - p) < (int64_t)lc + 64 + SITE * 16) {
      cvb_internal_set_error("cvb_pileup_format_rows: output buffer too small");
      return -1;
    }
    memcpy(p, ctg, lc);
    p += lc;
    p += snprintf(p, 32, " %lld ", (long long)center[i]);
    const int64_t nrp = center[i] - ref_off;  // refSeq[newRefPos-17 : newRefPos+16], Python slice semantics
    int64_t a = nrp - (F + 1), b = nrp + F;
    if (a < 0) a = 0;  // (emitted centres have a >= 0; a negative start would wrap in Python)
    if (b > ref_len) b = ref_len;
    if (b > a) { memcpy(p, ref_seq + a, (size_t)(b - a)); p += b - a; }
    const float* v = x + i * SITE;
    for (int k = 0; k < SITE; ++k) {
      *p++ = ' ';
      const float f = v[k];
      const int32_t q = (int32_t)f;
      if (f >= 0.f && f < 1e6f && (float)q == f) {  // a count: digits + ".0"
        char tmp[8];
        int nd = 0;
        int32_t t = q;
        do { tmp[nd++] = (char)('0' + t % 10); t /= 10; } while (t);
        while (nd) *p++ = tmp[--nd];
        *p++ = '.';
        *p++ = '0';
      } else {
        p += snprintf(p, 16, "%0.1f", (double)f);
      }
    }
    *p++ = '\n';
  }
  return (int64_t)(p - out);
}

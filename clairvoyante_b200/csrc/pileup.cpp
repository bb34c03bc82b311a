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
#include <thread>
#include <utility>
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

// one SAM record that passed the row filters, ready for the CIGAR walk
struct ReadRec {
  int64_t pos;          // 0-based POS
  const char* cig;
  const char* cig_e;
  const char* seq;
  int64_t seq_len;
  int64_t end;          // POS + reference span of the CIGAR (M, D, =, X)
  int64_t ord;          // ordinal inside the feed call
  bool flush;           // depthCap == 0 after this read: finished centres are flushed (:237)
};

// what a CIGAR walk works on: all of the handle's state (serial), or one thread's share of it (a contiguous range of the
// candidate list with its own centres, output list and tuple counter)
struct WalkCtx {
  std::map<int64_t, Center>* live;
  std::vector<std::pair<int64_t, Done>>* done;  // (ordinal of the read that flushed it, tensor)
  int64_t* available;
  size_t c_lo, c_hi;                            // candidate indices this context may activate
  int64_t start = 0, peak = 0;                  // tuples outstanding beyond `start` at most (threads only)
};

struct RawRec;

}  // namespace

struct cvb_pileup {
  std::string ref;            // reference bases of the fetched region
  int64_t ref_off = 0;        // refSeq index = refPos - ref_off  (ref_off = refStart - 1, or 0)
  std::vector<int64_t> cand;  // sorted unique candidate positions (1-based, as in the candidate list)
  int min_mq = 0, dcov = 250, min_cov = 0;
  bool left_edge = true;
  int threads = 1;
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
  void emit(int64_t c, const Center& ce, int64_t ord, std::vector<std::pair<int64_t, Done>>* out) const {  // GenerateTensor's tail (:54-59)
    const int64_t new_ref_pos = c - ref_off;
    if (new_ref_pos - (F + 1) >= 0 && ce.depth16 >= min_cov) {
      out->emplace_back();
      out->back().first = ord;
      out->back().second.center = c;
      memcpy(out->back().second.code, ce.code, sizeof(ce.code));
    }
  }
  void flush_all() {  // :248-252 (the reference does not return the slots here either; nothing follows)
    std::vector<std::pair<int64_t, Done>> out;
    for (auto& kv : live) emit(kv.first, kv.second, 0, &out);
    for (auto& d : out) done.push_back(d.second);
    live.clear();
  }
  bool admit(const RawRec& t, int64_t ord, ReadRec* r);
  void walk(const ReadRec& r, WalkCtx& w) const;
  void flush_before(const ReadRec& r, WalkCtx& w) const;
  void run_serial(const std::vector<ReadRec>& recs);
  bool run_parallel(const std::vector<ReadRec>& recs);
  void process(const std::vector<ReadRec>& recs);
};

namespace {

struct Active {
  int64_t c;
  Center* ce;
};

inline bool is_ws(char ch) { return ch == ' ' || ch == '\t' || ch == '\r' || ch == '\n' || ch == '\v' || ch == '\f'; }

inline void note_use(WalkCtx& w) {  // one tuple taken from the budget
  --*w.available;
  const int64_t out = w.start - *w.available;
  if (out > w.peak) w.peak = out;
}

// one tuple (:196-199 and the summation of :25-52)
inline void tuple_match(WalkCtx& w, const Active& a, int64_t ref_pos, char rb, char qb) {
  if (*w.available == 0) return;
  note_use(w);
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
inline void tuple_ins(WalkCtx& w, const Active& a, int64_t ref_pos, int64_t query_adv, char qb) {
  if (*w.available == 0) return;
  note_use(w);
  ++a.ce->slots;
  const int q = base_code(qb);
  if (q < 0) return;
  const int64_t d = ref_pos - a.c;
  if (d < -(F + 1) || d >= F) return;
  int64_t idx = d + F + 1 + query_adv;
  if (idx > 2 * F) idx = 2 * F;
  a.ce->code[idx * 16 + q * 4 + 1] += 1.f;
}
inline void tuple_del(WalkCtx& w, const Active& a, int64_t ref_pos, char rb) {
  if (*w.available == 0) return;
  note_use(w);
  ++a.ce->slots;
  const int r = base_code(rb);
  if (r < 0) return;
  const int64_t d = ref_pos - a.c;
  if (d < -(F + 1) || d >= F) return;
  a.ce->code[(d + F + 1) * 16 + r * 4 + 2] += 1.f;
}

}  // namespace

// The read loop's row handling (:150-181) in two steps, so that a feed call can tokenise its lines on several threads:
// tokenize() is a pure function of the line; admit() applies what depends on the order of the rows (counters, MAPQ filter,
// the depth-cap run).
namespace {
struct RawRec {
  int kind;  // 0 = blank or header (ignored), 1 = fewer than ten fields, 2 = a record
  int64_t pos;
  long mq;
  const char* cig;
  const char* cig_e;
  const char* seq;
  int64_t seq_len, end;
};

void tokenize(const char* p, const char* e, RawRec* r) {
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
  if (nf == 0 || fb[0][0] == '@') { r->kind = 0; return; }  // blank line, header
  if (nf < 10) { r->kind = 1; return; }
  r->kind = 2;
  char num[24];  // (strtoll semantics on a bounded copy: the fields are not NUL-terminated)
  size_t ln = (size_t)std::min<ptrdiff_t>(fe[3] - fb[3], 23);
  memcpy(num, fb[3], ln);
  num[ln] = 0;
  r->pos = strtoll(num, nullptr, 10) - 1;
  ln = (size_t)std::min<ptrdiff_t>(fe[4] - fb[4], 23);
  memcpy(num, fb[4], ln);
  num[ln] = 0;
  r->mq = strtol(num, nullptr, 10);
  r->cig = fb[5];
  r->cig_e = fe[5];
  r->seq = fb[9];
  r->seq_len = fe[9] - fb[9];
  int64_t span = 0, n = 0;
  for (const char* c = r->cig; c < r->cig_e; ++c) {
    if (*c >= '0' && *c <= '9') {
      n = n * 10 + (*c - '0');
      if (n > ((int64_t)1 << 40)) n = (int64_t)1 << 40;
    } else {
      if (*c == 'M' || *c == 'D' || *c == '=' || *c == 'X') span += n;
      n = 0;
    }
  }
  r->end = r->pos + span;
}
}  // namespace

bool cvb_pileup::admit(const RawRec& t, int64_t ord, ReadRec* r) {
  if (t.kind == 0) return false;
  ++reads;
  if (t.kind == 1) { ++malformed; return false; }
  if (t.mq < min_mq) return false;  // :165-166
  // depth cap (:174-181)
  if (previous_pos != t.pos) {
    previous_pos = t.pos;
    depth_cap = 0;
  } else {
    ++depth_cap;
    if (depth_cap >= dcov) return false;
  }
  ++reads_used;
  r->pos = t.pos;
  r->cig = t.cig;
  r->cig_e = t.cig_e;
  r->seq = t.seq;
  r->seq_len = t.seq_len;
  r->end = t.end;
  r->ord = ord;
  r->flush = depth_cap == 0;
  return true;
}

// the CIGAR walk of one read (:183-246) over the candidates w.c_lo .. w.c_hi
void cvb_pileup::walk(const ReadRec& rec, WalkCtx& w) const {
  const int64_t POS = rec.pos;
  const char* cig = rec.cig;
  const char* cig_e = rec.cig_e;
  const char* seq = rec.seq;
  const int64_t seq_len = rec.seq_len;
  auto query = [&](int64_t i) -> char { return (i >= 0 && i < seq_len) ? seq[i] : 'N'; };

  int64_t ref_pos = POS, query_pos = 0;
  std::vector<Active> active;  // activeSet, ascending centre (centres activate and retire in position order)
  size_t na = 0;               // next candidate this read has not activated yet
  bool na_init = false;
  auto activate = [&](int64_t rp) {  // `if refPos in beginToEnd` (:187-194, :224-231)
    if (!na_init) {              // candidates whose activation window [c-17, c+17) ended at or before rp never activate
      na = (size_t)(std::upper_bound(cand.begin() + (ptrdiff_t)w.c_lo, cand.begin() + (ptrdiff_t)w.c_hi, rp - (F + 1)) - cand.begin());
      na_init = true;
    }
    while (na < w.c_hi && cand[na] - (F + 1) <= rp) {
      const int64_t c = cand[na++];
      if (!left_edge && c - (F + 1) != rp) continue;  // without considerleftedge only the window start activates
      active.push_back(Active{c, &(*w.live)[c]});
    }
  };
  auto retire = [&](int64_t rp) {  // `if refPos in endToCenter` (:203-205, :232-234)
    for (size_t i = 0; i < active.size(); ++i)
      if (active[i].c + (F + 1) == rp) {
        active.erase(active.begin() + (ptrdiff_t)i);
        break;
      }
  };
  auto next_start = [&]() -> int64_t { return na < w.c_hi ? cand[na] - (F + 1) : INT64_MAX / 4; };  // (room for the subtraction below)

  while (cig < cig_e) {  // re.finditer(r"(\d+)([MIDNSHP=X])", CIGAR)
    if (*cig < '0' || *cig > '9') { ++cig; continue; }
    int64_t adv = 0;
    const char* q = cig;
    while (q < cig_e && *q >= '0' && *q <= '9') {
      adv = adv * 10 + (*q++ - '0');
      if (adv > ((int64_t)1 << 40)) adv = (int64_t)1 << 40;  // (absurd lengths saturate instead of overflowing)
    }
    if (q >= cig_e) break;
    const char op = *q;
    if (!strchr("MIDNSHP=X", op)) { cig = q; continue; }  // digits not followed by an op: the regex would retry later
    cig = q + 1;
    if (*w.available == 0) break;  // :184-185
    if (op == 'S') {
      query_pos += adv;
    } else if (op == 'M' || op == '=' || op == 'X') {
      for (int64_t i = 0; i < adv;) {
        activate(ref_pos);
        if (active.empty()) {
          // nothing open, and every window starting at or before ref_pos has been consumed: nothing can happen before the
          // next candidate's window start -- jump there (or to the end of the run) in one step
          const int64_t skip = std::max<int64_t>(1, std::min<int64_t>(adv - i, next_start() - ref_pos));
          ref_pos += skip;
          query_pos += skip;
          i += skip;
          continue;
        }
        const char rb = ref_at(ref_pos), qb = query(query_pos);
        for (const Active& a : active) tuple_match(w, a, ref_pos, rb, qb);
        retire(ref_pos);
        ++ref_pos;
        ++query_pos;
        ++i;
      }
    } else if (op == 'I') {
      for (int64_t i = 0; i < adv; ++i) {
        if (!active.empty()) {
          const char qb = query(query_pos);
          for (const Active& a : active) tuple_ins(w, a, ref_pos, i, qb);
        }
        ++query_pos;
      }
    } else if (op == 'D') {
      for (int64_t i = 0; i < adv;) {
        if (active.empty()) {
          activate(ref_pos);  // (a centre opened on a deleted base gets no tuple for it: the reference appends first, :218-223)
          if (active.empty()) {
            const int64_t skip = std::max<int64_t>(1, std::min<int64_t>(adv - i, next_start() - ref_pos));
            ref_pos += skip;
            i += skip;
          } else {
            ++ref_pos;
            ++i;
          }
          continue;
        }
        const char rb = ref_at(ref_pos);
        for (const Active& a : active) tuple_del(w, a, ref_pos, rb);
        activate(ref_pos);
        retire(ref_pos);
        ++ref_pos;
        ++i;
      }
    }  // N, H, P: no effect (the reference has no branch for them)
  }
  if (rec.flush) flush_before(rec, w);
}

void cvb_pileup::flush_before(const ReadRec& rec, WalkCtx& w) const {  // :237-246
  std::map<int64_t, Center>& lv = *w.live;
  while (!lv.empty() && lv.begin()->first + (F + 1) < rec.pos) {
    emit(lv.begin()->first, lv.begin()->second, rec.ord, w.done);
    *w.available += lv.begin()->second.slots;
    lv.erase(lv.begin());
  }
}

void cvb_pileup::run_serial(const std::vector<ReadRec>& recs) {
  std::vector<std::pair<int64_t, Done>> out;
  WalkCtx w{&live, &out, &available, 0, cand.size()};
  w.start = available;
  for (const ReadRec& r : recs) {
    walk(r, w);
    for (auto& d : out) done.push_back(d.second);
    out.clear();
  }
}

// The CIGAR walks of one feed call on `threads` threads: thread t owns a contiguous range of the candidate list, walks EVERY
// read against it (reads that do not reach the range cost a few jumps) and keeps its centres, finished tensors and tuple
// count to itself; afterwards the finished tensors are merged in the order the serial loop would have flushed them
// (ordinal of the flushing read, then position).  The 10 M outstanding-tuple budget is the one thing the centres share in the
// serial loop: the result is accepted only if the budget provably never ran out (sum of the threads' peaks <= what was
// available), otherwise the state is rolled back and the call runs serially.
bool cvb_pileup::run_parallel(const std::vector<ReadRec>& recs) {
  if (threads < 2 || recs.size() < 64 || cand.empty()) return false;
  // candidates this call can touch: from the oldest open centre / the first window that reaches the first read, to the last
  // window that starts before the end of the longest read
  int64_t lo_pos = INT64_MAX, hi_pos = INT64_MIN;
  for (const ReadRec& r : recs) {
    lo_pos = std::min(lo_pos, r.pos);
    hi_pos = std::max(hi_pos, r.end);
  }
  size_t lo = (size_t)(std::upper_bound(cand.begin(), cand.end(), lo_pos - (F + 1)) - cand.begin());
  if (!live.empty()) lo = std::min(lo, (size_t)(std::lower_bound(cand.begin(), cand.end(), live.begin()->first) - cand.begin()));
  const size_t hi = (size_t)(std::upper_bound(cand.begin(), cand.end(), hi_pos + (F + 1)) - cand.begin());
  if (hi <= lo) return false;
  const int T = (int)std::min<size_t>((size_t)threads, (hi - lo) / 8);
  if (T < 2) return false;
  const std::map<int64_t, Center> snapshot = live;  // (the few centres open across the call boundary)
  const int64_t avail0 = available;
  struct Part {
    std::map<int64_t, Center> live;
    std::vector<std::pair<int64_t, Done>> done;
    int64_t available = 0;
    WalkCtx w;
  };
  std::vector<Part> parts((size_t)T);
  for (int t = 0; t < T; ++t) {
    Part& pt = parts[(size_t)t];
    pt.available = avail0;
    pt.w = WalkCtx{&pt.live, &pt.done, &pt.available, lo + (hi - lo) * (size_t)t / (size_t)T, lo + (hi - lo) * (size_t)(t + 1) / (size_t)T};
    pt.w.start = avail0;
    if (t == 0) pt.w.c_lo = 0;  // (stray centres outside [lo, hi) stay with the edge threads)
    if (t == T - 1) pt.w.c_hi = cand.size();
  }
  while (!live.empty()) {  // hand every open centre to the thread that owns its candidate
    auto node = live.extract(live.begin());
    const size_t idx = (size_t)(std::lower_bound(cand.begin(), cand.end(), node.key()) - cand.begin());
    int t = 0;
    while (t + 1 < T && idx >= parts[(size_t)t].w.c_hi) ++t;
    parts[(size_t)t].live.insert(std::move(node));
  }
  auto work = [&](int t) {
    Part& pt = parts[(size_t)t];
    // reference positions at which this thread's candidates can be touched: [first window start, last window end]
    const int64_t w_lo = pt.w.c_lo < pt.w.c_hi ? cand[pt.w.c_lo] - (F + 1) : INT64_MAX;
    const int64_t w_hi = pt.w.c_lo < pt.w.c_hi ? cand[pt.w.c_hi - 1] + (F + 1) : INT64_MIN;
    for (const ReadRec& r : recs) {
      if (r.end < w_lo || r.pos > w_hi) {  // the read cannot open or extend any of them: only its flush matters
        if (r.flush) flush_before(r, pt.w);
        continue;
      }
      walk(r, pt.w);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  int64_t peaks = 0, net = 0;
  for (const Part& pt : parts) {
    peaks += pt.w.peak;
    net += avail0 - pt.available;
  }
  if (peaks > avail0) {  // the shared budget might have run out in the serial order: redo serially from the saved state
    live = snapshot;
    available = avail0;
    return false;
  }
  available = avail0 - net;
  std::vector<std::pair<int64_t, Done>*> all;
  for (Part& pt : parts) {
    live.merge(pt.live);
    for (auto& d : pt.done) all.push_back(&d);
  }
  std::stable_sort(all.begin(), all.end(), [](const std::pair<int64_t, Done>* a, const std::pair<int64_t, Done>* b) {
    return a->first != b->first ? a->first < b->first : a->second.center < b->second.center;
  });
  for (auto* d : all) done.push_back(d->second);
  return true;
}

void cvb_pileup::process(const std::vector<ReadRec>& recs) {
  if (recs.empty()) return;
  if (!run_parallel(recs)) run_serial(recs);
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

extern "C" int cvb_pileup_set_threads(cvb_pileup* s, int threads) {
  if (!s) return fail("cvb_pileup_set_threads: NULL handle");
  s->threads = threads < 1 ? 1 : (threads > 64 ? 64 : threads);
  return 0;
}

extern "C" int cvb_pileup_feed(cvb_pileup* s, const char* sam, int64_t len, int final_chunk) {
  if (!s || (!sam && len > 0) || len < 0) return fail("cvb_pileup_feed: bad argument");
  const char* p = sam;
  const char* e = sam + len;
  std::vector<std::pair<const char*, const char*>> lines;
  std::string first;  // the line completed from the previous chunk (its record points into this copy)
  if (!s->carry.empty()) {  // finish the line started in the previous chunk
    const char* nl = p < e ? (const char*)memchr(p, '\n', (size_t)(e - p)) : nullptr;
    if (!nl && !final_chunk) { s->carry.append(p, (size_t)(e - p)); return 0; }
    const char* stop = nl ? nl : e;
    s->carry.append(p, (size_t)(stop - p));
    first.swap(s->carry);
    lines.emplace_back(first.data(), first.data() + first.size());
    p = nl ? nl + 1 : e;
  }
  while (p < e) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
    if (!nl) {
      if (final_chunk) lines.emplace_back(p, e);
      else s->carry.assign(p, (size_t)(e - p));
      break;
    }
    lines.emplace_back(p, nl);
    p = nl + 1;
  }
  // tokenise (any order, threads), then admit in row order
  std::vector<RawRec> raw(lines.size());
  const int T = (int)std::min<size_t>((size_t)s->threads, lines.size() / 256);
  auto tok = [&](size_t a, size_t b) {
    for (size_t i = a; i < b; ++i) tokenize(lines[i].first, lines[i].second, &raw[i]);
  };
  if (T < 2) {
    tok(0, lines.size());
  } else {
    std::vector<std::thread> pool;
    const size_t per = (lines.size() + (size_t)T - 1) / (size_t)T;
    for (int t = 1; t < T; ++t) pool.emplace_back(tok, std::min(lines.size(), (size_t)t * per), std::min(lines.size(), (size_t)(t + 1) * per));
    tok(0, std::min(lines.size(), per));
    for (auto& th : pool) th.join();
  }
  std::vector<ReadRec> recs;
  recs.reserve(raw.size());
  ReadRec r;
  for (const RawRec& t : raw)
    if (s->admit(t, (int64_t)recs.size(), &r)) recs.push_back(r);
  s->process(recs);
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

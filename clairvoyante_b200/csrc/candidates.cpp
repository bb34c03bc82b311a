// candidates.cpp -- native variant-candidate extraction: SAM records + reference sequence -> candidate positions (host code,
// no CUDA).  First stage of the reference's calling pipeline `ExtractVariantCandidates | CreateTensor | callVar`
// (callVarBam.py:56-66); like CreateTensor.py it is a per-base CPython loop there.
//
// Follows dataPrepScripts/ExtractVariantCandidates.py (line numbers of /root/reference/dataPrepScripts/...):
//   :127-152  per SAM row: RNAME must equal ctgName, MAPQ filter, reads less than 55 % aligned
//             (1 - softclipped / (sum of all CIGAR lengths + 1) < 0.55) are skipped
//   :154-176  CIGAR walk: M/=/X count the read base at every reference position; I / D add one "I" / "D" at refPos-1 (D then
//             advances refPos); S advances the query; N/H/P do nothing (N does NOT advance refPos in the reference -- kept)
//   :178-213  positions left of the current read's start are final: region / BED filter, optional training subsample,
//             OutputCandidate, delete;  :215-243 the remaining positions in ascending order at end of input
//   :22-42    OutputCandidate: total = all seven counts; needs total >= minCoverage; counts sorted descending (stable);
//             candidate if (p0 <= 1 - threshold and p1 >= threshold) or the top key differs from the reference base;
//             row "ctg pos+1 refBase total k0 n0 .. k6 n6"
// The seven counters are the reference's dict {"A","C","G","T","I","D","N"}; ties in the stable sort keep the dict's
// iteration order, which for CPython 2.7 (64-bit, no hash randomisation) is A, C, D, G, I, N, T: a 7-item literal is
// presized to 8 slots and grows to 32 on the 6th insertion, and a one-character string c hashes to a slot (c ^ 1) & 31.
// Documented differences: (1) a read base that is not one of A C G T N raises KeyError in the reference; here it is counted
// as N; (2) out-of-range reference indices read as 'N' (reference: IndexError / negative-index wrap); (3) the training
// subsample (`--gen4Training`, random.uniform, unseeded in the reference) uses a counter-based hash of (seed, position);
// (4) rows with fewer than 10 fields are counted as malformed and skipped.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
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

// counter order = iteration order of the reference's dict (see header): A C D G I N T
enum { kA = 0, kC = 1, kD = 2, kG = 3, kI = 4, kN = 5, kT = 6 };
const char kKeys[7] = {'A', 'C', 'D', 'G', 'I', 'N', 'T'};

inline int read_base_slot(char c) {
  switch (c) {
    case 'A': return kA;
    case 'C': return kC;
    case 'G': return kG;
    case 'T': return kT;
    default: return kN;  // 'N'; anything else would be a KeyError in the reference
  }
}
constexpr int64_t kMaxWindow = (int64_t)1 << 27;  // open positions one handle will hold (reads are position-sorted: a real window is a read length)
inline bool is_ws(char ch) { return ch == ' ' || ch == '\t' || ch == '\r' || ch == '\n' || ch == '\v' || ch == '\f'; }

struct Counts { int32_t n[7]; int32_t touched; Counts() { memset(n, 0, sizeof(n)); touched = 0; } };

inline double hash_uniform01(uint64_t seed, uint64_t idx) {  // splitmix64 finaliser -> [0, 1)
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace

struct cvb_candidates {
  std::string ctg, ref;
  int64_t ref_off = 0;                 // refSeq index = pos - ref_off
  bool has_region = false;
  int64_t ctg_start = 0, ctg_end = 0;  // as compared by the reference: 0-based sweep against (ctgStart + 1, ctgEnd)
  bool has_bed = false;
  std::vector<std::pair<int64_t, int64_t>> bed;  // half-open [begin, end) of this contig, sorted by begin
  std::vector<int64_t> bed_max_end;              // prefix maximum of end (point query)
  int min_mq = 0;
  double min_cov = 4, threshold = 0.125;
  bool subsample = false;
  double output_prob = 1.0;
  uint64_t seed = 0;
  // the reference's `pileup` dict: positions >= `sweep` live in a window indexed by (position - sweep) -- reads arrive in
  // position order, so the open positions are a dense run -- positions created behind the sweep in a small ordered map
  std::vector<Counts> ring;      // power-of-two capacity; slots outside [head, head + count) are all-zero
  size_t head = 0, count = 0;    // the window covers positions [sweep, sweep + count)
  inline Counts& slot(size_t i) { return ring[(head + i) & (ring.size() - 1)]; }
  void reserve_window(size_t n) {  // make positions sweep .. sweep + n - 1 addressable
    if (n > ring.size()) {
      size_t cap = ring.empty() ? 1024 : ring.size();
      while (cap < n) cap <<= 1;
      std::vector<Counts> bigger(cap);
      for (size_t i = 0; i < count; ++i) bigger[i] = slot(i);
      ring.swap(bigger);
      head = 0;
    }
    if (n > count) count = n;
  }
  std::map<int64_t, Counts> late;
  std::string out_text;
  std::vector<int64_t> out_pos;
  int64_t reads = 0, processed = 0, malformed = 0;
  std::string carry;

  inline char ref_at(int64_t pos) const {
    const int64_t i = pos - ref_off;
    return (i >= 0 && i < (int64_t)ref.size()) ? ref[(size_t)i] : 'N';
  }
  bool in_bed(int64_t p) const {  // len(tree[ctg].search(p)) != 0
    // intervals with begin <= p: indices [0, k); any of them with end > p?
    const size_t k = (size_t)(std::upper_bound(bed.begin(), bed.end(), std::make_pair(p, INT64_MAX)) - bed.begin());
    return k > 0 && bed_max_end[k - 1] > p;
  }
  int threads = 1;
  void format_position(int64_t p, const Counts& c, std::string* text, std::vector<int64_t>* pos) const;
  void finish_position(int64_t p, const Counts& c) { format_position(p, c, &out_text, &out_pos); }
  void read_line(const char* p, const char* e);
  void feed_parallel(const std::vector<std::pair<const char*, const char*>>& lines);
  int64_t sweep = 0;
  inline Counts& at(int64_t pos) {  // pileup.setdefault(pos, {...})
    if (pos < sweep) { Counts& c = late[pos]; c.touched = 1; return c; }
    const size_t i = (size_t)(pos - sweep);
    if (i >= count) reserve_window(i + 1);
    Counts& c = slot(i);
    c.touched = 1;
    return c;
  }
  // `while sweep < POS` (:178-213): positions in [sweep, POS) are final.  A position BEHIND the sweep can still be created
  // later (an insertion / deletion is booked at refPos-1, which may lie left of the read's own start); the reference never
  // revisits it in this loop -- it stays in the dict until the end-of-input pass, and so it does here (`late`).
  void sweep_before(int64_t pos0) {
    while (sweep < pos0 && count > 0) {
      Counts& c = slot(0);
      if (c.touched) finish_position(sweep, c);
      c = Counts();
      head = (head + 1) & (ring.size() - 1);
      --count;
      ++sweep;
    }
    if (pos0 > sweep) sweep = pos0;
  }
  void finish_all() {  // :215-243: the remaining positions in ascending order
    for (auto& kv : late) finish_position(kv.first, kv.second);
    late.clear();
    for (size_t i = 0; i < count; ++i) {
      Counts& c = slot(i);
      if (c.touched) finish_position(sweep + (int64_t)i, c);
      c = Counts();
    }
    count = 0;
  }
  int64_t open_positions() const {
    int64_t n = (int64_t)late.size();
    for (const Counts& c : ring) n += c.touched;
    return n;
  }
};

void cvb_candidates::format_position(int64_t p, const Counts& c, std::string* text, std::vector<int64_t>* pos) const {
  // ---- :185-204 region / BED / subsample
  bool flag = false;
  if (has_region) {
    if (p >= ctg_start && p <= ctg_end) flag = has_bed ? in_bed(p) : true;
  } else if (has_bed) {
    flag = in_bed(p);
  } else {
    flag = true;
  }
  if (flag && subsample && hash_uniform01(seed, (uint64_t)p) > output_prob) flag = false;
  if (!flag) return;
  // ---- OutputCandidate (:22-42)
  int64_t total = 0;
  for (int i = 0; i < 7; ++i) total += c.n[i];
  if ((double)total < min_cov) return;
  const int64_t denom = total == 0 ? 1 : total;
  int order[7] = {0, 1, 2, 3, 4, 5, 6};
  std::stable_sort(order, order + 7, [&](int a, int b) { return c.n[a] > c.n[b]; });
  const double p0 = (double)c.n[order[0]] / (double)denom, p1 = (double)c.n[order[1]] / (double)denom;
  const char ref_base = ref_at(p);
  if (!((p0 <= 1.0 - threshold && p1 >= threshold) || kKeys[order[0]] != ref_base)) return;
  char buf[256];
  int n = snprintf(buf, sizeof(buf), "%s %lld %c %lld", ctg.c_str(), (long long)(p + 1), ref_base, (long long)total);
  text->append(buf, (size_t)n);
  for (int i = 0; i < 7; ++i) {
    n = snprintf(buf, sizeof(buf), " %c %d", kKeys[order[i]], c.n[order[i]]);
    text->append(buf, (size_t)n);
  }
  text->push_back('\n');
  pos->push_back(p + 1);
}

void cvb_candidates::read_line(const char* p, const char* e) {
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
  if (nf == 0) return;
  if (fb[0][0] == '@') return;
  ++reads;
  if (nf < 10) { ++malformed; return; }
  if ((size_t)(fe[2] - fb[2]) != ctg.size() || memcmp(fb[2], ctg.data(), ctg.size()) != 0) return;  // :133-135
  const int64_t POS = strtoll(std::string(fb[3], fe[3]).c_str(), nullptr, 10) - 1;
  const long MQ = strtol(std::string(fb[4], fe[4]).c_str(), nullptr, 10);
  if (MQ < min_mq) return;
  const char* const cig0 = fb[5];
  const char* const cig_e = fe[5];
  const char* seq = fb[9];
  const int64_t seq_len = fe[9] - fb[9];

  auto next_op = [&](const char*& c, int64_t& adv, char& op) -> bool {  // re.finditer(r"(\d+)([MIDNSHP=X])", CIGAR)
    while (c < cig_e) {
      if (*c < '0' || *c > '9') { ++c; continue; }
      int64_t a = 0;
      const char* q = c;
      while (q < cig_e && *q >= '0' && *q <= '9') {
        a = a * 10 + (*q++ - '0');
        if (a > ((int64_t)1 << 40)) a = (int64_t)1 << 40;  // (absurd lengths saturate instead of overflowing)
      }
      if (q >= cig_e) { c = cig_e; return false; }
      if (!strchr("MIDNSHP=X", *q)) { c = q; continue; }
      adv = a;
      op = *q;
      c = q + 1;
      return true;
    }
    return false;
  };
  int64_t skip_base = 0, total_aln = 0, adv = 0, ref_adv = 0;
  char op = 0;
  for (const char* c = cig0; next_op(c, adv, op);) {
    total_aln += adv;
    if (op == 'S') skip_base += adv;
    if (op == 'M' || op == '=' || op == 'X' || op == 'D') ref_adv += adv;
  }
  if (1.0 - (double)skip_base / (double)(total_aln + 1) < 0.55) return;  // :150-151
  if (POS - sweep > kMaxWindow || ref_adv > kMaxWindow) { ++malformed; return; }  // (a row that would open a window of > 128 M positions)
  ++processed;
  int64_t ref_pos = POS, query_pos = 0;
  for (const char* c = cig0; next_op(c, adv, op);) {
    if (op == 'S') {
      query_pos += adv;
    } else if (op == 'M' || op == '=' || op == 'X') {
      if (ref_pos >= sweep && adv > 0) {  // the whole run lies in the window: one bounds check for the run
        const size_t i0 = (size_t)(ref_pos - sweep);
        if (i0 + (size_t)adv > count) reserve_window(i0 + (size_t)adv);
        for (int64_t i = 0; i < adv; ++i) {
          const char b = (query_pos + i >= 0 && query_pos + i < seq_len) ? seq[query_pos + i] : 'N';
          Counts& c = slot(i0 + (size_t)i);
          c.touched = 1;
          ++c.n[read_base_slot(b)];
        }
        ref_pos += adv;
        query_pos += adv;
      } else {
        for (int64_t i = 0; i < adv; ++i) {
          const char b = (query_pos >= 0 && query_pos < seq_len) ? seq[query_pos] : 'N';
          ++at(ref_pos).n[read_base_slot(b)];
          ++ref_pos;
          ++query_pos;
        }
      }
    } else if (op == 'I') {
      ++at(ref_pos - 1).n[kI];
      query_pos += adv;
    } else if (op == 'D') {
      ++at(ref_pos - 1).n[kD];
      ref_pos += adv;
    }
  }
  sweep_before(POS);  // :178-213 (`while sweep < POS`)
}

// ---------------------------------------------------------------------------------------------------------------------
// Several host threads per feed call (cvb_candidates_set_threads).  The serial loop above is the definition; this path
// reproduces it exactly.  What couples the reads is only the `sweep` value each read sees (the largest POS of the reads
// processed before it) -- known after a serial pass over the row filters -- so the counting can be split by POSITION: thread t
// owns a contiguous range of reference positions, walks every read, books the contributions that fall into its range (into
// its copy of the open window when the position is at or beyond that read's sweep, into a `late` map otherwise) and reports
// the positions the sweep passes; the reports are merged in the order the serial loop produces them (ordinal of the
// sweeping read, then position).
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct CRaw {  // pure function of one line
  int kind;    // 0 blank / header, 1 fewer than ten fields, 2 another contig, 3 a record of this contig
  int64_t pos, end, seq_len;
  long mq;
  bool aligned55;
  const char* cig0;
  const char* cig_e;
  const char* seq;
};

struct CRec {  // an admitted read
  int64_t pos, end, seq_len, sweep, ord;
  const char* cig0;
  const char* cig_e;
  const char* seq;
};

inline bool next_cigar_op(const char*& c, const char* cig_e, int64_t& adv, char& op) {  // re.finditer(r"(\d+)([MIDNSHP=X])", CIGAR)
  while (c < cig_e) {
    if (*c < '0' || *c > '9') { ++c; continue; }
    int64_t a = 0;
    const char* q = c;
    while (q < cig_e && *q >= '0' && *q <= '9') {
      a = a * 10 + (*q++ - '0');
      if (a > ((int64_t)1 << 40)) a = (int64_t)1 << 40;
    }
    if (q >= cig_e) { c = cig_e; return false; }
    if (!strchr("MIDNSHP=X", *q)) { c = q; continue; }
    adv = a;
    op = *q;
    c = q + 1;
    return true;
  }
  return false;
}

void tokenize_candidate_row(const char* p, const char* e, const std::string& ctg, CRaw* r) {
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
  if (nf == 0 || fb[0][0] == '@') { r->kind = 0; return; }
  if (nf < 10) { r->kind = 1; return; }
  if ((size_t)(fe[2] - fb[2]) != ctg.size() || memcmp(fb[2], ctg.data(), ctg.size()) != 0) { r->kind = 2; return; }
  r->kind = 3;
  char num[24];
  size_t ln = (size_t)std::min<ptrdiff_t>(fe[3] - fb[3], 23);
  memcpy(num, fb[3], ln);
  num[ln] = 0;
  r->pos = strtoll(num, nullptr, 10) - 1;
  ln = (size_t)std::min<ptrdiff_t>(fe[4] - fb[4], 23);
  memcpy(num, fb[4], ln);
  num[ln] = 0;
  r->mq = strtol(num, nullptr, 10);
  r->cig0 = fb[5];
  r->cig_e = fe[5];
  r->seq = fb[9];
  r->seq_len = fe[9] - fb[9];
  int64_t skip_base = 0, total_aln = 0, adv = 0, ref_adv = 0;
  char op = 0;
  for (const char* c = r->cig0; next_cigar_op(c, r->cig_e, adv, op);) {
    total_aln += adv;
    if (op == 'S') skip_base += adv;
    if (op == 'M' || op == '=' || op == 'X' || op == 'D') ref_adv += adv;
  }
  r->aligned55 = !(1.0 - (double)skip_base / (double)(total_aln + 1) < 0.55);  // :150-151
  r->end = r->pos + ref_adv;  // one past the last reference position the walk can reach
}

struct Emit {
  int64_t ord, pos;
  size_t off, len;
};

struct CPart {
  int64_t lo = 0, hi = 0;          // positions [lo, hi)
  std::vector<Counts> win;         // its share of the open window
  std::map<int64_t, Counts> late;
  std::string text;
  std::vector<Emit> emits;
  int64_t cursor = 0;              // positions below it have been swept
};

}  // namespace

void cvb_candidates::feed_parallel(const std::vector<std::pair<const char*, const char*>>& lines) {
  // ---- tokenise on the threads, admit in row order (counters, filters, the sweep value every read sees)
  std::vector<CRaw> raw(lines.size());
  const int TT = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, lines.size() / 256));
  auto tok = [&](size_t a, size_t b) {
    for (size_t i = a; i < b; ++i) tokenize_candidate_row(lines[i].first, lines[i].second, ctg, &raw[i]);
  };
  {
    std::vector<std::thread> pool;
    const size_t per = (lines.size() + (size_t)TT - 1) / (size_t)TT;
    for (int t = 1; t < TT; ++t) pool.emplace_back(tok, std::min(lines.size(), (size_t)t * per), std::min(lines.size(), (size_t)(t + 1) * per));
    tok(0, std::min(lines.size(), per));
    for (auto& th : pool) th.join();
  }
  std::vector<CRec> recs;
  recs.reserve(raw.size());
  int64_t sw = sweep, hi_pos = sweep + (int64_t)count;
  for (const CRaw& t : raw) {
    if (t.kind == 0) continue;
    ++reads;
    if (t.kind == 1) { ++malformed; continue; }
    if (t.kind == 2 || t.mq < min_mq || !t.aligned55) continue;
    if (t.pos - sw > kMaxWindow || t.end - t.pos > kMaxWindow) { ++malformed; continue; }
    ++processed;
    recs.push_back(CRec{t.pos, t.end, t.seq_len, sw, (int64_t)recs.size(), t.cig0, t.cig_e, t.seq});
    if (t.pos > sw) sw = t.pos;  // sweep_before(POS)
    hi_pos = std::max(hi_pos, t.end + 1);
  }
  if (recs.empty()) return;
  const int64_t lo_pos = sweep;
  if (hi_pos <= lo_pos) hi_pos = lo_pos + 1;
  const int T = (int)std::max<int64_t>(1, std::min<int64_t>(threads, (hi_pos - lo_pos) / 2048));
  std::vector<CPart> parts((size_t)T);
  for (int t = 0; t < T; ++t) {
    CPart& pt = parts[(size_t)t];
    pt.lo = lo_pos + (hi_pos - lo_pos) * t / T;
    pt.hi = lo_pos + (hi_pos - lo_pos) * (t + 1) / T;
    pt.win.assign((size_t)(pt.hi - pt.lo), Counts());
    pt.cursor = pt.lo;
    for (int64_t p = std::max(pt.lo, sweep); p < std::min(pt.hi, sweep + (int64_t)count); ++p) pt.win[(size_t)(p - pt.lo)] = slot((size_t)(p - sweep));
  }
  auto work = [&](int t) {
    CPart& pt = parts[(size_t)t];
    const bool first = t == 0;
    std::vector<int64_t> no_pos;
    auto book = [&](int64_t q, int s, int64_t sweep_i) {  // one contribution at position q (owned by this thread)
      Counts& c = q >= sweep_i ? pt.win[(size_t)(q - pt.lo)] : pt.late[q];
      c.touched = 1;
      ++c.n[s];
    };
    auto owns = [&](int64_t q) { return (q >= pt.lo && q < pt.hi) || (first && q < pt.lo); };
    for (const CRec& r : recs) {
      if (!(r.end < pt.lo - 1 || r.pos - 1 >= pt.hi) || (first && r.pos - 1 < pt.lo)) {
        int64_t ref_pos = r.pos, query_pos = 0, adv = 0;
        char op = 0;
        for (const char* c = r.cig0; next_cigar_op(c, r.cig_e, adv, op);) {
          if (op == 'S') {
            query_pos += adv;
          } else if (op == 'M' || op == '=' || op == 'X') {
            int64_t a = std::max<int64_t>(ref_pos, first ? INT64_MIN / 4 : pt.lo), b = std::min<int64_t>(ref_pos + adv, pt.hi);
            for (int64_t q = a; q < b; ++q) {
              const int64_t qi = query_pos + (q - ref_pos);
              const char base = (qi >= 0 && qi < r.seq_len) ? r.seq[qi] : 'N';
              book(q, read_base_slot(base), r.sweep);
            }
            ref_pos += adv;
            query_pos += adv;
          } else if (op == 'I') {
            if (owns(ref_pos - 1)) book(ref_pos - 1, kI, r.sweep);
            query_pos += adv;
          } else if (op == 'D') {
            if (owns(ref_pos - 1)) book(ref_pos - 1, kD, r.sweep);
            ref_pos += adv;
          }
        }
      }
      // `while sweep < POS`: the positions [sweep_i, POS) of this range are final
      const int64_t upto = std::min(r.pos, pt.hi);
      for (int64_t p = std::max(pt.cursor, r.sweep); p < upto; ++p) {
        Counts& c = pt.win[(size_t)(p - pt.lo)];
        if (c.touched) {
          const size_t off = pt.text.size();
          no_pos.clear();
          format_position(p, c, &pt.text, &no_pos);
          if (!no_pos.empty()) pt.emits.push_back(Emit{r.ord, p, off, pt.text.size() - off});
          c = Counts();
        }
      }
      if (upto > pt.cursor) pt.cursor = upto;
    }
  };
  {
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
  }
  // ---- merge: reports in serial order, then the open window and the late map
  std::vector<std::pair<const Emit*, const CPart*>> all;
  for (const CPart& pt : parts)
    for (const Emit& em : pt.emits) all.emplace_back(&em, &pt);
  std::stable_sort(all.begin(), all.end(), [](const std::pair<const Emit*, const CPart*>& a, const std::pair<const Emit*, const CPart*>& b) {
    return a.first->ord != b.first->ord ? a.first->ord < b.first->ord : a.first->pos < b.first->pos;
  });
  for (auto& em : all) {
    out_text.append(em.second->text, em.first->off, em.first->len);
    out_pos.push_back(em.first->pos + 1);
  }
  std::fill(ring.begin(), ring.end(), Counts());
  head = 0;
  count = 0;
  sweep = sw;
  for (CPart& pt : parts) {
    for (int64_t p = std::max(pt.lo, sweep); p < pt.hi; ++p) {
      const Counts& c = pt.win[(size_t)(p - pt.lo)];
      if (c.touched) at(p) = c;
    }
    for (auto& kv : pt.late) {
      Counts& d = late[kv.first];
      d.touched = 1;
      for (int i = 0; i < 7; ++i) d.n[i] += kv.second.n[i];
    }
  }
}

extern "C" int cvb_candidates_create(const char* ctg_name, const char* ref_seq, int64_t ref_len, int64_t ref_start, int64_t ctg_start,
                                     int64_t ctg_end, const int64_t* bed_begin, const int64_t* bed_end, int64_t n_bed, int min_mq,
                                     double min_coverage, double threshold, double output_prob, uint64_t seed,
                                     cvb_candidates** out) {
  if (!out || !ctg_name || (!ref_seq && ref_len > 0) || ref_len < 0 || n_bed < -1 || (n_bed > 0 && (!bed_begin || !bed_end)))
    return fail("cvb_candidates_create: bad argument");
  cvb_candidates* s = new (std::nothrow) cvb_candidates();
  if (!s) return fail("cvb_candidates_create: out of memory");
  s->ctg = ctg_name;
  s->ref.assign(ref_seq ? ref_seq : "", (size_t)ref_len);
  s->ref_off = ref_start > 0 ? ref_start - 1 : 0;
  s->has_region = ctg_start >= 0 && ctg_end >= 0;
  s->ctg_start = ctg_start;
  s->ctg_end = ctg_end;
  s->has_bed = n_bed >= 0;
  for (int64_t i = 0; i < n_bed; ++i) s->bed.emplace_back(bed_begin[i], bed_end[i]);
  std::sort(s->bed.begin(), s->bed.end());
  int64_t mx = INT64_MIN;
  for (auto& iv : s->bed) { mx = std::max(mx, iv.second); s->bed_max_end.push_back(mx); }
  s->min_mq = min_mq;
  s->min_cov = min_coverage;
  s->threshold = threshold;
  s->subsample = output_prob >= 0.0;
  s->output_prob = output_prob;
  s->seed = seed;
  *out = s;
  return 0;
}

extern "C" int cvb_candidates_destroy(cvb_candidates* s) {
  delete s;
  return 0;
}

extern "C" int cvb_candidates_set_threads(cvb_candidates* s, int threads) {
  if (!s) return fail("cvb_candidates_set_threads: NULL handle");
  s->threads = threads < 1 ? 1 : (threads > 64 ? 64 : threads);
  return 0;
}

extern "C" int cvb_candidates_feed(cvb_candidates* s, const char* sam, int64_t len, int final_chunk) {
  if (!s || (!sam && len > 0) || len < 0) return fail("cvb_candidates_feed: bad argument");
  const char* p = sam;
  const char* e = sam + len;
  std::vector<std::pair<const char*, const char*>> lines;
  std::string first;  // the line completed from the previous chunk
  if (!s->carry.empty()) {
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
  if (s->threads >= 2 && lines.size() >= 512) {
    s->feed_parallel(lines);
  } else {
    for (auto& ln : lines) s->read_line(ln.first, ln.second);
  }
  if (final_chunk) s->finish_all();
  return 0;
}

extern "C" int64_t cvb_candidates_pending_bytes(const cvb_candidates* s) { return s ? (int64_t)s->out_text.size() : 0; }
extern "C" int64_t cvb_candidates_pending(const cvb_candidates* s) { return s ? (int64_t)s->out_pos.size() : 0; }

extern "C" int cvb_candidates_take(cvb_candidates* s, char* text, int64_t text_cap, int64_t* text_len, int64_t* pos, int64_t pos_cap,
                                   int64_t* n_pos) {
  if (!s || !text_len || !n_pos) return fail("cvb_candidates_take: bad argument");
  if ((int64_t)s->out_text.size() > text_cap || (int64_t)s->out_pos.size() > pos_cap || (text_cap > 0 && !text) || (pos_cap > 0 && !pos))
    return fail("cvb_candidates_take: buffers too small (see cvb_candidates_pending_bytes / _pending)");
  if (!s->out_text.empty()) memcpy(text, s->out_text.data(), s->out_text.size());
  if (!s->out_pos.empty()) memcpy(pos, s->out_pos.data(), s->out_pos.size() * sizeof(int64_t));
  *text_len = (int64_t)s->out_text.size();
  *n_pos = (int64_t)s->out_pos.size();
  s->out_text.clear();
  s->out_pos.clear();
  return 0;
}

extern "C" int cvb_candidates_stats(const cvb_candidates* s, int64_t stats[4]) {
  if (!s || !stats) return fail("cvb_candidates_stats: bad argument");
  stats[0] = s->reads;
  stats[1] = s->processed;
  stats[2] = s->malformed;
  stats[3] = s->open_positions();
  return 0;
}

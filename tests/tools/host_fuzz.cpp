// host_fuzz.cpp -- memory-safety harness for the host-side C++ stages of libcvb200 (no CUDA, no GPU).
//
// Built by tests/test_host_fuzz_cpu.py with -fsanitize=address,undefined from the stage sources where they lie
// (blosc_frame.cpp, text_feed.cpp, pileup.cpp, candidates.cpp, crc32c.cpp, sam_view.cpp, vcf_text.cpp) and run with a fixed seed: valid inputs,
// then the same inputs with bytes flipped / cut / duplicated.  The checks are "returns, leaks nothing, touches nothing
// outside its buffers" plus the round trips that must still hold (blosc encode -> decode).
//   host_fuzz <seed> <iterations>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/cvb200.h"

static char g_last[512];
void cvb_internal_set_error(const char* msg) { snprintf(g_last, sizeof(g_last), "%s", msg); }  // lives in cvb200.cu in the product

namespace {

uint64_t g_state = 1;
uint64_t rnd() {  // splitmix64
  uint64_t z = (g_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
int64_t rnd_below(int64_t n) { return n > 0 ? (int64_t)(rnd() % (uint64_t)n) : 0; }

void die(const char* what) {
  fprintf(stderr, "host_fuzz: %s (last error: %s)\n", what, g_last);
  exit(2);
}

// byte-level damage: flips, cuts, duplicated stretches, inserted garbage
void mutate(std::vector<uint8_t>& v) {
  const int ops = 1 + (int)rnd_below(4);
  for (int o = 0; o < ops && !v.empty(); ++o) {
    const int kind = (int)rnd_below(6);
    const int64_t at = rnd_below((int64_t)v.size());
    if (kind == 0) v[at] ^= (uint8_t)(1u << rnd_below(8));
    else if (kind == 1) v[at] = (uint8_t)rnd();
    else if (kind == 2) v.resize((size_t)at);
    else if (kind == 3) v.erase(v.begin() + at, v.begin() + at + rnd_below((int64_t)v.size() - at));
    else if (kind == 4) {
      const int64_t len = rnd_below(64);
      std::vector<uint8_t> junk((size_t)len);
      for (auto& b : junk) b = (uint8_t)rnd();
      v.insert(v.begin() + at, junk.begin(), junk.end());
    } else {
      const int64_t len = rnd_below(((int64_t)v.size() - at) < 200 ? (int64_t)v.size() - at : 200);
      std::vector<uint8_t> dup(v.begin() + at, v.begin() + at + len);
      v.insert(v.begin() + at, dup.begin(), dup.end());
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- blosc
std::vector<uint8_t> tensor_like(int64_t n) {  // count-tensor-like bytes: runs of zeros, small floats, some noise
  std::vector<uint8_t> v((size_t)n, 0);
  const int style = (int)rnd_below(4);
  if (style == 0) return v;
  if (style == 1) { for (auto& b : v) b = (uint8_t)rnd(); return v; }
  for (int64_t i = 0; i + 4 <= n; i += 4) {
    if (rnd_below(style == 2 ? 4 : 40) == 0) {
      const float f = (float)rnd_below(60);
      memcpy(&v[(size_t)i], &f, 4);
    }
  }
  return v;
}

void fuzz_blosc() {
  static const int sizes[] = {0, 1, 3, 4, 15, 16, 17, 100, 511, 512, 513, 4096, 70000, 300000, 600001};
  const int64_t n = sizes[rnd_below(sizeof(sizes) / sizeof(sizes[0]))] + rnd_below(3);
  static const int ts[] = {1, 2, 4, 8, 3, 16, 17, 255};
  const int typesize = ts[rnd_below(8)];
  const int shuf = (int)rnd_below(2);
  std::vector<uint8_t> src = tensor_like(n);
  std::vector<uint8_t> frame((size_t)cvb_blosc_compress_bound(n));
  int64_t fn = 0;
  if (cvb_blosc_compress(src.data(), n, typesize, shuf, frame.data(), (int64_t)frame.size(), &fn)) die("blosc compress failed");
  if (fn > (int64_t)frame.size()) die("blosc compress overran its bound");
  frame.resize((size_t)fn);
  int64_t nb = -1, cb = -1;
  int tsz = -1, fl = -1;
  if (cvb_blosc_info(frame.data(), fn, &nb, &cb, &tsz, &fl) || nb != n || cb != fn) die("blosc info mismatch");
  {
    std::vector<uint8_t> exact(frame);  // heap copy with exact bounds so that ASan sees over-reads of the frame
    std::vector<uint8_t> out((size_t)n);
    int64_t on = -1;
    if (cvb_blosc_decompress(exact.data(), fn, out.data(), n, &on) || on != n) die("blosc round trip failed");
    if (n && memcmp(out.data(), src.data(), (size_t)n)) die("blosc round trip differs");
    if (n > 0 && cvb_blosc_decompress(exact.data(), fn, out.data(), n - 1, &on) == 0) die("blosc accepted a short destination");
  }
  for (int k = 0; k < 8; ++k) {  // damaged frames: any return value, no crash; output buffer sized from the header as callers do
    std::vector<uint8_t> bad(frame);
    mutate(bad);
    if (rnd_below(3) == 0 && bad.size() >= 16) {  // target the header fields
      const int64_t at = 2 + rnd_below(14);
      bad[(size_t)at] = (uint8_t)rnd();
    }
    int64_t hb = 0;
    if (cvb_blosc_info(bad.data(), (int64_t)bad.size(), &hb, nullptr, nullptr, nullptr)) continue;
    if (hb > (2 << 20)) hb = 2 << 20;  // (callers allocate what the header says; keep the harness small)
    std::vector<uint8_t> out((size_t)hb);
    int64_t on = 0;
    std::vector<uint8_t> exact(bad);
    cvb_blosc_decompress(exact.data(), (int64_t)exact.size(), out.data(), hb, &on);
  }
}

// ------------------------------------------------------------------------------------------------------------ text feed
std::string tensor_row() {
  std::string r = "chr" + std::to_string(rnd_below(22)) + " " + std::to_string(rnd_below(100000000)) + " ";
  static const char B[] = "ACGTNacgt";
  for (int i = 0; i < 33; ++i) r += B[rnd_below(i == 16 ? 5 : 9)];
  for (int i = 0; i < 528; ++i) {
    r += ' ';
    const int s = (int)rnd_below(20);
    if (s < 14) r += std::to_string(rnd_below(80)) + ".0";
    else if (s < 17) r += std::to_string(rnd_below(80));
    else if (s == 17) r += "1e1";
    else if (s == 18) r += "-3.25";
    else r += "0.5";
  }
  r += rnd_below(8) ? "\n" : "\r\n";
  return r;
}

void fuzz_text() {
  std::string text;
  const int rows = 1 + (int)rnd_below(6);
  for (int i = 0; i < rows; ++i) text += tensor_row();
  std::vector<uint8_t> v(text.begin(), text.end());
  if (rnd_below(4)) mutate(v);
  const int64_t max_lines = 1 + rnd_below(10);
  std::vector<float> x((size_t)max_lines * 528);
  std::vector<int64_t> meta((size_t)max_lines * 10);
  std::vector<uint8_t> exact(v);
  int64_t lines = 0, kept = 0, consumed = 0, off = 0;
  for (int pass = 0; pass < 20 && off < (int64_t)exact.size(); ++pass) {
    const int fin = (int)rnd_below(2);
    if (cvb_parse_tensor_text((const char*)exact.data() + off, (int64_t)exact.size() - off, fin, max_lines, 1 + (int)rnd_below(3),
                              x.data(), meta.data(), &lines, &kept, &consumed))
      die("text parser reported an error on plain bytes");
    if (lines < 0 || lines > max_lines || kept < 0 || kept > lines || consumed < 0 || consumed > (int64_t)exact.size() - off)
      die("text parser returned counts outside their ranges");
    for (int64_t i = 0; i < lines; ++i) {
      const int64_t* m = &meta[(size_t)i * 10];
      if (m[1] < 0 || m[2] < 0 || m[1] + m[2] > (int64_t)exact.size() - off) die("text parser: line span outside the buffer");
      if (m[0] == CVB_LINE_KEPT)
        for (int f = 3; f < 9; f += 2)
          if (m[f] < 0 || m[f + 1] < 0 || m[f] + m[f + 1] > (int64_t)exact.size() - off) die("text parser: field span outside the buffer");
    }
    {  // the position strings of the kept rows: capacity = the consumed bytes, as the Python side sizes it
      std::vector<char> names((size_t)consumed + 16);
      const int64_t w = cvb_tensor_text_positions((const char*)exact.data() + off, meta.data(), lines, names.data(), (int64_t)names.size());
      if (w < 0 || w > (int64_t)names.size()) die("tensor_text_positions: capacity rule violated");
      int64_t nl = 0;
      for (int64_t i = 0; i < w; ++i) nl += names[(size_t)i] == '\n';
      if (nl != kept) die("tensor_text_positions: one line per kept row expected");
      if (kept > 0 && cvb_tensor_text_positions((const char*)exact.data() + off, meta.data(), lines, names.data(), 3) >= 0)
        die("tensor_text_positions ignored its capacity");
    }
    if (consumed == 0) break;
    off += consumed;
  }
}

// ------------------------------------------------------------------------------------------------- pile-up and candidates
std::string make_ref(int64_t n) {
  static const char B[] = "ACGTACGTACGTNacgt";
  std::string r((size_t)n, 'A');
  for (auto& c : r) c = B[rnd_below(17)];
  return r;
}

std::string sam_rows(const std::string& ref, int reads, const char* ctg, bool sorted) {
  std::string out;
  int64_t pos = 1 + rnd_below(20);
  static const char OPS[] = "MMMMMMIDSN=XHP";
  for (int r = 0; r < reads; ++r) {
    pos += rnd_below(12);
    std::string cigar, seq;
    const int nops = 1 + (int)rnd_below(6);
    for (int o = 0; o < nops; ++o) {
      const int64_t len = 1 + rnd_below(o == 0 || rnd_below(3) ? 60 : 8);
      const char op = OPS[rnd_below(14)];
      cigar += std::to_string(len) + op;
      if (op == 'M' || op == 'I' || op == 'S' || op == '=' || op == 'X')
        for (int64_t k = 0; k < len; ++k) seq += "ACGTNacgt"[rnd_below(rnd_below(10) ? 4 : 9)];
    }
    if (rnd_below(30) == 0) cigar = "*";
    out += "r" + std::to_string(r) + "\t" + std::to_string(rnd_below(4) ? 0 : 16) + "\t" + (rnd_below(20) ? ctg : "other") + "\t" +
           std::to_string(pos) + "\t" + std::to_string(rnd_below(61)) + "\t" + cigar + "\t*\t0\t0\t" + (seq.empty() ? "*" : seq) + "\t*\n";
    if (!sorted && pos > (int64_t)ref.size() + 50) pos = 1;  // (unsorted input: must not crash either)
  }
  return out;
}

void feed_chunks(const std::vector<uint8_t>& sam, int (*feed)(void*, const char*, int64_t, int), void* h) {
  int64_t off = 0;
  const int64_t n = (int64_t)sam.size();
  if (n > 0 && rnd_below(3) == 0) {  // everything in one call (the path that may use several threads)
    std::vector<uint8_t> whole(sam);
    feed(h, (const char*)whole.data(), n, 0);
    feed(h, nullptr, 0, 1);
    return;
  }
  while (off < n) {
    int64_t len = 1 + rnd_below(rnd_below(3) ? 4000 : 40);
    if (len > n - off) len = n - off;
    std::vector<uint8_t> piece(sam.begin() + off, sam.begin() + off + len);  // exact-size heap copy
    feed(h, (const char*)piece.data(), len, 0);
    off += len;
  }
  feed(h, nullptr, 0, 1);
}

int feed_pileup(void* h, const char* s, int64_t n, int fin) { return cvb_pileup_feed((cvb_pileup*)h, s, n, fin); }
int feed_cand(void* h, const char* s, int64_t n, int fin) { return cvb_candidates_feed((cvb_candidates*)h, s, n, fin); }

void fuzz_alignments() {
  const int64_t ref_len = 200 + rnd_below(1500);
  const std::string ref = make_ref(ref_len);
  const int64_t ref_start = rnd_below(3) ? 0 : rnd_below(100);
  const bool sorted = rnd_below(4) != 0;
  std::string sam = sam_rows(ref, 5 + (int)rnd_below(rnd_below(3) ? 120 : (rnd_below(2) ? 500 : 1500)), "ctg", sorted);
  std::vector<uint8_t> v(sam.begin(), sam.end());
  const bool intact = rnd_below(3) == 0;
  if (!intact) mutate(v);

  {  // samtools-view filter on text: chunked, exact-size buffers, with and without a region
    const bool region = rnd_below(2);
    const int64_t a = region ? rnd_below(ref_len) : -1, b = region ? a + rnd_below(600) : -1;
    std::vector<uint8_t> pending;
    int64_t off = 0, kept = 0;
    const int64_t n = (int64_t)v.size();
    while (off < n || !pending.empty()) {
      int64_t len = 1 + rnd_below(rnd_below(3) ? 3000 : 50);
      if (len > n - off) len = n - off;
      pending.insert(pending.end(), v.begin() + off, v.begin() + off + len);
      off += len;
      const int fin = off >= n;
      std::vector<uint8_t> in(pending);  // exact-size heap copies
      std::vector<char> out(in.size() + 1);
      int64_t on = -1, used = -1;
      if (cvb_sam_view((const char*)in.data(), (int64_t)in.size(), fin, "ctg", 2308, a, b, out.data(), &on, &used)) die("sam_view failed");
      if (on < 0 || on > (int64_t)in.size() + 1 || used < 0 || used > (int64_t)in.size()) die("sam_view: counts outside their ranges");
      if (fin && used != (int64_t)in.size()) die("sam_view left input behind on the final chunk");
      kept += on;
      pending.erase(pending.begin(), pending.begin() + used);
      if (fin) break;
    }
    if (kept > n + 1) die("sam_view produced more than it was given");
  }
  {  // candidates
    std::vector<int64_t> bb, be;
    if (rnd_below(2)) {
      int64_t p = rnd_below(100);
      for (int i = 0; i < 4; ++i) { bb.push_back(p); p += 1 + rnd_below(300); be.push_back(p); p += rnd_below(100); }
    }
    const bool region = rnd_below(2);
    cvb_candidates* c = nullptr;
    if (cvb_candidates_create("ctg", ref.data(), ref_len, ref_start, region ? rnd_below(300) : -1, region ? 300 + rnd_below(2000) : -1,
                              bb.empty() ? nullptr : bb.data(), bb.empty() ? nullptr : be.data(), (int64_t)bb.size(), (int)rnd_below(30),
                              (double)(1 + rnd_below(6)), 0.05 + 0.1 * (double)rnd_below(5), rnd_below(2) ? 0.0 : 0.3, rnd(), &c))
      die("candidates_create failed");
    cvb_candidates_set_threads(c, 1 + (int)rnd_below(4));
    feed_chunks(v, feed_cand, c);
    const int64_t bytes = cvb_candidates_pending_bytes(c), np = cvb_candidates_pending(c);
    if (bytes < 0 || np < 0) die("candidates: negative pending counts");
    std::vector<char> text((size_t)bytes + 1);
    std::vector<int64_t> pos((size_t)np + 1);
    int64_t tl = 0, nn = 0;
    if (cvb_candidates_take(c, text.data(), bytes, &tl, pos.data(), np, &nn)) die("candidates_take failed");
    if (tl != bytes || nn != np) die("candidates_take moved a different amount than pending reported");
    int64_t st[4];
    cvb_candidates_stats(c, st);
    cvb_candidates_destroy(c);
  }
  {  // pile-up over arbitrary candidate positions (sorted, some outside the reference)
    std::vector<int64_t> cand;
    int64_t p = rnd_below(50) - 10;
    const int nc = (int)rnd_below(40);
    for (int i = 0; i < nc; ++i) { p += 1 + rnd_below(80); cand.push_back(p); }
    cvb_pileup* h = nullptr;
    if (cvb_pileup_create(ref.data(), ref_len, ref_start, cand.empty() ? nullptr : cand.data(), (int64_t)cand.size(), (int)rnd_below(30),
                          (int)rnd_below(3) ? 1000 : 3, (int)rnd_below(5), (int)rnd_below(2), &h))
      return;  // (a refused configuration is fine)
    cvb_pileup_set_threads(h, 1 + (int)rnd_below(4));
    feed_chunks(v, feed_pileup, h);
    const int64_t ready = cvb_pileup_ready(h);
    if (ready < 0 || (sorted && intact && ready > (int64_t)cand.size()))  // (unsorted input can re-open a centre, as in the reference)
      die("pile-up: more tensors than candidates");
    std::vector<float> x((size_t)(ready + 1) * 528);
    std::vector<int64_t> center((size_t)ready + 1);
    int64_t got = 0;
    if (cvb_pileup_take(h, ready, x.data(), center.data(), &got) || got != ready) die("pileup_take failed");
    for (int64_t i = 1; i < got && sorted && intact; ++i)  // (the order is promised for position-sorted input only)
      if (center[(size_t)i] <= center[(size_t)i - 1]) die("pile-up: centres not ascending");
    for (int64_t i = 0; i < got * 528; ++i)
      if (!(x[(size_t)i] >= 0.0f && x[(size_t)i] <= 1e6f)) die("pile-up: count outside [0, 1e6]");
    if (got) {
      const int64_t cap = got * 9000 + 64;
      std::vector<char> rows((size_t)cap);
      const int64_t w = cvb_pileup_format_rows("ctg", center.data(), x.data(), got, ref.data(), ref_len, ref_start, rows.data(), cap);
      if (w < 0 || w > cap) die("pileup_format_rows overran");
      if (cvb_pileup_format_rows("ctg", center.data(), x.data(), got, ref.data(), ref_len, ref_start, rows.data(), 100) >= 0)
        die("pileup_format_rows accepted a tiny buffer");
    }
    int64_t st[4];
    cvb_pileup_stats(h, st);
    cvb_pileup_destroy(h);
  }
}

// ---------------------------------------------------------------------------------------------------------- VCF records
void fuzz_vcf() {
  const int64_t n = rnd_below(40);
  std::vector<float> x((size_t)n * 528), base((size_t)n * 4), z((size_t)n * 2), t((size_t)n * 4), l((size_t)n * 6);
  for (auto& v : x) v = rnd_below(6) == 0 ? (float)rnd_below(60) : 0.f;
  auto fill = [](std::vector<float>& a) {
    for (auto& v : a) v = (float)rnd_below(1000) / 1000.f;
    for (size_t i = 0; i + 1 < a.size(); i += 7) a[i + 1] = a[i];  // ties
  };
  fill(base); fill(z); fill(t); fill(l);
  std::string pos;
  for (int64_t j = 0; j < n; ++j) {
    pos += (rnd_below(9) ? "chr1" : "HLA:A*01") + std::string(":") + std::to_string(rnd_below(1000000)) + ":";
    for (int k = 0; k < 33; ++k) pos += "ACGTN"[rnd_below(k == 16 ? 4 : 5)];
    if (j + 1 < n || rnd_below(2)) pos += "\n";
  }
  std::vector<uint8_t> pv(pos.begin(), pos.end());
  const bool intact = rnd_below(2);
  if (!intact) mutate(pv);
  const int64_t cap = (int64_t)pv.size() + 256 * n + 64;
  std::vector<char> out((size_t)cap);
  const int64_t w = cvb_vcf_records(x.data(), (const char*)pv.data(), (int64_t)pv.size(), base.data(), z.data(), t.data(), l.data(), n,
                                    (int)rnd_below(2), rnd_below(2) ? -1 : (int)rnd_below(300), out.data(), cap);
  if (w > cap) die("vcf_records overran");
  if (intact && w < 0) die("vcf_records refused well-formed input");
  if (n > 3 && cvb_vcf_records(x.data(), (const char*)pv.data(), (int64_t)pv.size(), base.data(), z.data(), t.data(), l.data(), n, 1, -1,
                               out.data(), 50) > 50)
    die("vcf_records ignored its capacity");
}

void fuzz_crc() {
  const int64_t n = rnd_below(5000);
  std::vector<uint8_t> v((size_t)n);
  for (auto& b : v) b = (uint8_t)rnd();
  const int64_t cut = rnd_below(n + 1);
  const uint32_t whole = cvb_crc32c(0, v.data(), n);
  const uint32_t parts = cvb_crc32c(cvb_crc32c(0, v.data(), cut), v.data() + cut, n - cut);
  if (whole != parts) die("crc32c is not incremental");
}

}  // namespace

int main(int argc, char** argv) {
  g_state = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
  const int iters = argc > 2 ? atoi(argv[2]) : 200;
  const char* only = argc > 3 ? argv[3] : "";
  for (int i = 0; i < iters; ++i) {
    if (!*only || !strcmp(only, "blosc")) fuzz_blosc();
    if (!*only || !strcmp(only, "text")) fuzz_text();
    if (!*only || !strcmp(only, "aln")) fuzz_alignments();
    if (!*only || !strcmp(only, "crc")) fuzz_crc();
    if (!*only || !strcmp(only, "vcf")) fuzz_vcf();
  }
  if (strcmp(cvb_crc32c(0, "123456789", 9) == 0xE3069283u ? "ok" : "bad", "ok")) die("crc32c known answer");
  printf("host_fuzz: %d iterations clean\n", iters);
  return 0;
}

// vcf_text.cpp -- the VCF records of one batch of calls (host code, no CUDA): native form of the per-site loop of the
// reference's `Output` (clairvoyante/callVar.py:58-153), which formats one record at a time in Python.
//
// Per site j (line numbers of /root/reference/clairvoyante/callVar.py):
//   :59       skipped unless showRef or argmax(varType) != REF
//   :61-65    varType / zygosity / indel length = first maximum of each head
//   :69-72    QUAL = int(-4.343 * log((t2 z2 l2 + 1e-300) / (t1 z1 l1 + 1e-300))): the products of the largest (x1) and second
//             largest (x2) outputs are float32 products taken left to right, the sums, ratio and log are double, int()
//             truncates toward zero
//   :74-79    FILTER "." or PASS / LowQual against --qual
//   :81-83    candidate bases = the two largest outputs of the base head; argsort()[::-1], i.e. among equals the HIGHER index
//             first (ascending stable order reversed)
//   :86-89    DP = sum(x[F,:,0]) + sum(x[F+1,:,1]) + sum(x[F+1,:,2]) + sum(x[F,:,3]) in float32, left to right; DP == 0: no record
//   :92-100   REF / SNP: ALT = best base that is not the reference base (REF call: the reference base), AF = x[F,alt,3] / DP
//   :101-124  INS: length 0 reads as 1; AF = sum(x[F+1,:,1]) / DP; inserted bases = argmax of the insertion channel at
//             F+1 ..; for the ">4" class the length is inferred: positions are taken while k < F+5 or
//             sum(ins) >= 0.125 * sum(ref); inferred >= F -> "<INS>" + SVTYPE=INS
//   :125-146  DEL: same inference on the deletion channel; REF = refSeq[F : F+len+1], ALT = refSeq[F]; inferred >= F -> "<DEL>"
//   :147-153  LENGUESS, genotype 0/0 | 0/1 | 1/1, the record "%s\t%d\t.\t%s\t%s\t%d\t%s\t%s\tGT:GQ:DP:AF\t%s:%d:%d:%.4f"
// Checked against oracle/callvar_output.py (scalar restatement) and against the VCF files the reference's own Output wrote
// (tests/golden/reference_run.npz).  NaN outputs are not ordered the way NumPy orders them (the model never produces them).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/cvb200.h"

void cvb_internal_set_error(const char* msg);  // cvb200.cu

namespace {

constexpr int F = 16, H = 33, MAXLEN = 5;
const char BASES[] = "ACGT";

inline int argmax(const float* v, int n) {
  int b = 0;
  for (int i = 1; i < n; ++i)
    if (v[i] > v[b]) b = i;
  return b;
}

// largest and second largest value (np.sort(v)[::-1][:2])
inline void top2(const float* v, int n, float* a, float* b) {
  float m1 = v[0], m2 = v[1];
  if (m2 > m1) { const float t = m1; m1 = m2; m2 = t; }
  for (int i = 2; i < n; ++i) {
    if (v[i] > m1) { m2 = m1; m1 = v[i]; }
    else if (v[i] > m2) m2 = v[i];
  }
  *a = m1;
  *b = m2;
}

inline float sum4(const float* v, int stride) {  // python sum(): ((0 + v0) + v1) + v2) + v3 in float32
  volatile float s = 0.f;
  for (int k = 0; k < 4; ++k) s = s + v[k * stride];
  return s;
}

inline int base_index(char c) {
  switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
  }
}

}  // namespace

extern "C" int64_t cvb_vcf_records(const float* x, const char* pos, int64_t pos_len, const float* base, const float* z,
                                   const float* t, const float* l, int64_t n, int show_ref, int qual_cut, char* out, int64_t cap) {
  if (n < 0 || (n > 0 && (!x || !pos || !base || !z || !t || !l || !out))) {
    cvb_internal_set_error("cvb_vcf_records: bad argument");
    return -1;
  }
  const char* p = pos;
  const char* const pe = pos + pos_len;
  char* o = out;
  char* const oe = out + cap;
  for (int64_t j = 0; j < n; ++j) {
    // "chrom:coord:SEQ\n"
    if (p >= pe) {
      cvb_internal_set_error("cvb_vcf_records: fewer position strings than sites");
      return -1;
    }
    const char* le = static_cast<const char*>(memchr(p, '\n', (size_t)(pe - p)));
    if (!le) le = pe;
    const char* line = p;
    p = le < pe ? le + 1 : pe;
    const float* tj = t + j * 4;
    int var_type = argmax(tj, 4);
    if (!show_ref && var_type == 0) continue;
    const char* c2 = le;  // last ':'
    while (c2 > line && c2[-1] != ':') --c2;
    const char* c1 = c2 > line ? c2 - 1 : line;  // -> the ':' before SEQ
    const char* q = c1;
    while (q > line && q[-1] != ':') --q;  // start of coord
    if (c2 == line || q == line) {
      cvb_internal_set_error("cvb_vcf_records: position string is not chrom:pos:seq");
      return -1;
    }
    const char* seq = c2;
    const int seq_len = (int)(le - c2);
    const int chrom_len = (int)(q - 1 - line);
    if (seq_len < H) {
      cvb_internal_set_error("cvb_vcf_records: reference context shorter than 33 bases");
      return -1;
    }
    const long long coord = strtoll(q, nullptr, 10);
    const float* xj = x + j * (H * 16);
    const int zyg = argmax(z + j * 2, 2);
    int var_len = argmax(l + j * 6, 6);
    float t1, t2, z1, z2, l1, l2;
    top2(tj, 4, &t1, &t2);
    top2(z + j * 2, 2, &z1, &z2);
    top2(l + j * 6, 6, &l1, &l2);
    volatile float num_f = t2 * z2;  // float32 products, left to right
    num_f = num_f * l2;
    volatile float den_f = t1 * z1;
    den_f = den_f * l1;
    const double ratio = ((double)num_f + 1e-300) / ((double)den_f + 1e-300);
    const long long qual = (long long)(-4.343 * log(ratio));
    const char* filt = ".";
    if (qual_cut >= 0) filt = qual >= qual_cut ? "PASS" : "LowQual";
    // DP
    volatile float dp = sum4(xj + F * 16 + 0, 4);
    dp = dp + sum4(xj + (F + 1) * 16 + 1, 4);
    dp = dp + sum4(xj + (F + 1) * 16 + 2, 4);
    dp = dp + sum4(xj + F * 16 + 3, 4);
    if (dp == 0.f) continue;
    char ref_s[40], alt_s[48], info[48];
    int ref_n = 0, alt_n = 0, inferred = 0;
    info[0] = 0;
    float af = 0.f;
    if (var_type <= 1) {
      ref_s[ref_n++] = seq[F];
      char alt = seq[F];
      if (var_type == 1) {
        // argsort()[::-1] of the base head: ascending stable order reversed
        int idx[4] = {0, 1, 2, 3};
        const float* bj = base + j * 4;
        for (int a = 1; a < 4; ++a) {
          const int k = idx[a];
          int b = a - 1;
          while (b >= 0 && bj[idx[b]] > bj[k]) { idx[b + 1] = idx[b]; --b; }
          idx[b + 1] = k;
        }
        const char b1 = BASES[idx[3]], b2 = BASES[idx[2]];
        alt = b1 != seq[F] ? b1 : b2;
      }
      alt_s[alt_n++] = alt;
      const int bi = base_index(alt);
      if (bi < 0) {
        cvb_internal_set_error("cvb_vcf_records: reference base at the centre is not A, C, G or T");
        return -1;
      }
      af = xj[F * 16 + bi * 4 + 3] / dp;
    } else if (var_type == 2) {
      if (var_len == 0) var_len = 1;
      af = sum4(xj + (F + 1) * 16 + 1, 4) / dp;
      char ins[40];
      int ins_n = 0;
      if (var_len != MAXLEN) {
        for (int k = F + 1; k < F + var_len + 1; ++k) {
          const float v[4] = {xj[k * 16 + 1], xj[k * 16 + 5], xj[k * 16 + 9], xj[k * 16 + 13]};
          ins[ins_n++] = BASES[argmax(v, 4)];
        }
      } else {
        for (int k = F + 1; k < 2 * F + 1; ++k) {
          const float v[4] = {xj[k * 16 + 1], xj[k * 16 + 5], xj[k * 16 + 9], xj[k * 16 + 13]};
          if (k < F + MAXLEN || (double)sum4(xj + k * 16 + 1, 4) >= 0.125 * (double)sum4(xj + k * 16 + 0, 4)) {
            ++inferred;
            ins[ins_n++] = BASES[argmax(v, 4)];
          } else {
            break;
          }
        }
      }
      ref_s[ref_n++] = seq[F];
      if (inferred >= F) {
        memcpy(alt_s, "<INS>", 5);
        alt_n = 5;
        strcpy(info, "SVTYPE=INS");
      } else {
        alt_s[alt_n++] = seq[F];
        memcpy(alt_s + alt_n, ins, (size_t)ins_n);
        alt_n += ins_n;
      }
    } else {
      if (var_len == 0) var_len = 1;
      af = sum4(xj + (F + 1) * 16 + 2, 4) / dp;
      if (var_len == MAXLEN) {
        for (int k = F + 1; k < 2 * F + 1; ++k) {
          if (k < F + MAXLEN || (double)sum4(xj + k * 16 + 2, 4) >= 0.125 * (double)sum4(xj + k * 16 + 0, 4)) ++inferred;
          else break;
        }
      }
      if (inferred >= F) {
        ref_s[ref_n++] = seq[F];
        memcpy(alt_s, "<DEL>", 5);
        alt_n = 5;
        strcpy(info, "SVTYPE=DEL");
      } else {
        int take = (var_len != MAXLEN ? var_len : inferred) + 1;  // refSeq[F : F + len + 1] (a Python slice: clipped at the end)
        if (take > seq_len - F) take = seq_len - F;
        if (take > 32) take = 32;
        memcpy(ref_s, seq + F, (size_t)take);
        ref_n = take;
        alt_s[alt_n++] = seq[F];
      }
    }
    if (inferred > 0 && inferred < F) {
      char tmp[24];
      snprintf(tmp, sizeof(tmp), "%sLENGUESS=%d", info[0] ? ";" : "", inferred);
      strcat(info, tmp);
    }
    const char* gt = var_type == 0 ? "0/0" : (zyg == 0 ? "0/1" : "1/1");
    const int64_t room = oe - o;
    if (room < chrom_len + 220) {
      cvb_internal_set_error("cvb_vcf_records: output buffer too small");
      return -1;
    }
    memcpy(o, line, (size_t)chrom_len);
    o += chrom_len;
    o += snprintf(o, (size_t)(oe - o), "\t%lld\t.\t%.*s\t%.*s\t%lld\t%s\t%s\tGT:GQ:DP:AF\t%s:%lld:%lld:%.4f\n", coord, ref_n, ref_s, alt_n,
                  alt_s, qual, filt, info[0] ? info : ".", gt, qual, (long long)dp, (double)af);
  }
  return o - out;
}

"""Per-site restatement of the reference's VCF record logic (clairvoyante/callVar.py:50-153).
TEST INFRASTRUCTURE ONLY (see oracle/cv_oracle.py header).  One candidate at a time, scalar Python, following the
reference's decision order; tests compare clairvoyante_b200.callVar.Output (batched NumPy) against it.
Pinned against the reference itself: the VCF records the reference's own callVar.Test / Output wrote for 630 tensors and a
fixed probability table (tests/golden/reference_run.npz, made by tests/golden/make_golden_reference_run.py) are reproduced
line for line (tests/test_reference_run_cpu.py)."""
from math import log

import numpy as np

F = 16                      # param.flankingBaseNum
BASES = "ACGT"
MAXLEN = 5                  # callVar.py:18
MIN_AF = 0.125              # callVar.py:19


def vcf_line(x, pos, base, z, t, l, show_ref=False, qual_cut=None):
    """x (33,4,4) float32; pos 'chrom:coord:seq'; base/z/t/l the four head outputs of one site.  Returns the VCF
    line or None when nothing is printed."""
    var_type = int(np.argmax(t))                                   # :58-60
    if not show_ref and var_type == 0:
        return None
    zyg = int(np.argmax(z)); var_len = int(np.argmax(l))           # :62-64
    chrom, coord, ref_seq = pos.split(":")                         # :66
    st, sz, sl = np.sort(t)[::-1], np.sort(z)[::-1], np.sort(l)[::-1]
    # :72 -- under the reference's NumPy 1.x the products stay in the inputs' dtype (float32 from TensorFlow) and adding the
    # Python float promotes to float64; NumPy 2 (NEP 50) would keep float32 + 1e-300 = float32, hence the explicit float64
    qual = int(-4.343 * log((np.float64(st[1] * sz[1] * sl[1]) + 1e-300) / (np.float64(st[0] * sz[0] * sl[0]) + 1e-300)))
    filt = "."
    if qual_cut is not None:
        filt = "PASS" if qual >= qual_cut else "LowQual"            # :75-79
    order = base.argsort()[::-1]                                   # :81
    b1, b2 = BASES[order[0]], BASES[order[1]]
    dp = sum(x[F, :, 0]) + sum(x[F + 1, :, 1]) + sum(x[F + 1, :, 2]) + sum(x[F, :, 3])   # :88-89
    if dp == 0:
        return None
    ref = alt = ""; guess = 0; info = []; af = 0.0
    coord = int(coord)
    if var_type in (0, 1):                                         # :92-100
        ref = ref_seq[F]
        alt = (b1 if b1 != ref else b2) if var_type == 1 else ref
        af = x[F, BASES.index(alt), 3] / dp
    elif var_type == 2:                                            # :101-124
        if var_len == 0:
            var_len = 1
        af = sum(x[F + 1, :, 1]) / dp
        if var_len != MAXLEN:
            for k in range(F + 1, F + var_len + 1):
                alt += BASES[int(np.argmax(x[k, :, 1]))]
        else:
            for k in range(F + 1, 2 * F + 1):
                if k < F + MAXLEN or sum(x[k, :, 1]) >= MIN_AF * sum(x[k, :, 0]):
                    guess += 1
                    alt += BASES[int(np.argmax(x[k, :, 1]))]
                else:
                    break
        ref = ref_seq[F]
        if guess >= F:
            alt = "<INS>"; info.append("SVTYPE=INS")
        else:
            alt = ref + alt
    else:                                                          # :125-146
        if var_len == 0:
            var_len = 1
        af = sum(x[F + 1, :, 2]) / dp
        if var_len == MAXLEN:
            for k in range(F + 1, 2 * F + 1):
                if k < F + MAXLEN or sum(x[k, :, 2]) >= MIN_AF * sum(x[k, :, 0]):
                    guess += 1
                else:
                    break
        if guess >= F:
            ref = ref_seq[F]; alt = "<DEL>"; info.append("SVTYPE=DEL")
        elif var_len != MAXLEN:
            ref = ref_seq[F:F + var_len + 1]; alt = ref_seq[F]
        else:
            ref = ref_seq[F:F + guess + 1]; alt = ref_seq[F]
    if 0 < guess < F:
        info.append("LENGUESS=%d" % guess)                         # :147
    gt = "0/0" if var_type == 0 else ("0/1" if zyg == 0 else "1/1")  # :151-153
    return "%s\t%d\t.\t%s\t%s\t%d\t%s\t%s\tGT:GQ:DP:AF\t%s:%d:%d:%.4f" % (
        chrom, coord, ref, alt, qual, filt, ";".join(info) if info else ".", gt, qual, dp, af)

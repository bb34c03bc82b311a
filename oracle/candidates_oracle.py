"""CPU restatement of the reference's candidate extraction (dataPrepScripts/ExtractVariantCandidates.py) -- TEST
INFRASTRUCTURE ONLY (see oracle/cv_oracle.py header for who may import oracle/).

Follows MakeCandidates (:53-253) and OutputCandidate (:22-42) statement by statement with the reference's `pileup` dict of
per-position counters and its `sweep` loop.  The native extractor (clairvoyante_b200/csrc/candidates.cpp) is checked against
this on the same SAM text.

Pinned against the reference itself: tests/golden/reference_run.npz holds the rows the reference's own script printed for
three scenarios (defaults; region + BED + MAPQ / threshold / coverage options; low threshold) when it was executed in the
build container with stand-ins for samtools / gzip / intervaltree and a replayed CPython-2.7 dict order
(tests/golden/make_golden_reference_run.py); this restatement and the native stage reproduce them byte for byte
(tests/test_reference_run_cpu.py).  What had to be restated rather than copied: (a) the counters are a Python-2 dict literal {"A","C","G","T","I","D","N"} whose `.items()` order decides ties in the
stable descending sort (:33) -- for CPython 2.7 (64-bit, no hash randomisation) that order is A, C, D, G, I, N, T (an
8-slot presized table grows to 32 slots on the sixth insertion; a one-character string hashes to slot (ord(c) ^ 1) & 31);
under PyPy, which the reference also supports, the order would be the insertion order; (b) the BED lookup
`len(tree.search(p)) != 0` is a point query on half-open intervals (intervaltree 2.0.1).  Stated differences (shared with
the product): read bases outside "ACGTN" count as N (reference: KeyError); out-of-range reference indices read as 'N';
rows with fewer than ten fields are skipped; the --gen4Training subsample draws from a counter-based hash instead of
`random.uniform` (unseeded in the reference, so not reproducible there either).
"""
import re

cigarRe = r"(\d+)([MIDNSHP=X])"
KEY_ORDER = ("A", "C", "D", "G", "I", "N", "T")      # CPython-2.7 iteration order of the reference's counter dict


def _new_counts():
    return {k: 0 for k in KEY_ORDER}                   # (insertion-ordered in Python 3: .items() follows KEY_ORDER)


def hash_uniform01(seed, idx):
    """splitmix64 finaliser of (seed, idx) -> [0, 1); same as candidates.cpp"""
    m = (1 << 64) - 1
    z = (seed + idx * 0x9E3779B97F4A7C15 + 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    z = z ^ (z >> 31)
    return (z >> 11) / 9007199254740992.0


def output_candidate(ctgName, pos, baseCount, refBase, minCoverage, threshold):
    """OutputCandidate (:22-42) -> output row or None"""
    totalCount = 0
    totalCount += sum(x[1] for x in baseCount)
    if totalCount < minCoverage:
        return None
    denominator = totalCount
    if denominator == 0:
        denominator = 1
    baseCount.sort(key=lambda x: -x[1])
    p0 = float(baseCount[0][1]) / denominator
    p1 = float(baseCount[1][1]) / denominator
    if (p0 <= 1.0 - threshold and p1 >= threshold) or baseCount[0][0] != refBase:
        output = [ctgName, pos + 1, refBase, totalCount]
        output.extend(["%s %d" % x for x in baseCount])
        return " ".join([str(x) for x in output])
    return None


def make_candidates(sam_text, ctgName, ref_seq, ref_start=None, ctgStart=None, ctgEnd=None, bed=None, minMQ=0,
                    minCoverage=4, threshold=0.125, output_prob=None, seed=0):
    """MakeCandidates' main loop (:127-243).  ctgStart is the value the reference compares with (its --ctgStart + 1, :63);
    bed = list of half-open (begin, end) or None.  Returns the output rows."""
    ref_off = 0 if ref_start is None else ref_start - 1

    def ref_at(p):
        i = p - ref_off
        return ref_seq[i] if 0 <= i < len(ref_seq) else "N"

    def in_bed(p):
        return any(b <= p < e for b, e in bed)

    def flag_for(p):
        outputFlag = 0
        if ctgStart is not None and ctgEnd is not None:
            if p >= ctgStart and p <= ctgEnd:
                if bed is not None:
                    if in_bed(p):
                        outputFlag = 1
                else:
                    outputFlag = 1
        elif bed is not None:
            if in_bed(p):
                outputFlag = 1
        else:
            outputFlag = 1
        if output_prob is not None and outputFlag == 1:
            if hash_uniform01(seed, p) > output_prob:
                outputFlag = 0
        return outputFlag

    out = []
    pileup = {}
    sweep = 0
    for line in sam_text.split("\n"):
        l = line.strip().split()
        if not l or l[0][0] == "@":
            continue
        if len(l) < 10:
            continue
        RNAME = l[2]
        if RNAME != ctgName:
            continue
        POS = int(l[3]) - 1
        MQ = int(l[4])
        CIGAR = l[5]
        SEQ = l[9]
        refPos = POS
        queryPos = 0
        if MQ < minMQ:
            continue
        skipBase = 0
        totalAlnPos = 0
        for m in re.finditer(cigarRe, CIGAR):
            advance = int(m.group(1))
            totalAlnPos += advance
            if m.group(2) == "S":
                skipBase += advance
        if 1.0 - float(skipBase) / (totalAlnPos + 1) < 0.55:
            continue
        for m in re.finditer(cigarRe, CIGAR):
            advance = int(m.group(1))
            if m.group(2) == "S":
                queryPos += advance
                continue
            if m.group(2) in ("M", "=", "X"):
                matches = []
                for i in range(advance):
                    matches.append((refPos, SEQ[queryPos] if 0 <= queryPos < len(SEQ) else "N"))
                    refPos += 1
                    queryPos += 1
                for pos, base in matches:
                    pileup.setdefault(pos, _new_counts())
                    pileup[pos][base if base in "ACGT" else "N"] += 1
            elif m.group(2) == "I":
                pileup.setdefault(refPos - 1, _new_counts())
                pileup[refPos - 1]["I"] += 1
                for i in range(advance):
                    queryPos += 1
            elif m.group(2) == "D":
                pileup.setdefault(refPos - 1, _new_counts())
                pileup[refPos - 1]["D"] += 1
                for i in range(advance):
                    refPos += 1
        while sweep < POS:
            flag = pileup.get(sweep)
            if flag is None:
                sweep += 1
                continue
            baseCount = list(pileup[sweep].items())
            refBase = ref_at(sweep)
            row = None
            if flag_for(sweep) == 1:
                row = output_candidate(ctgName, sweep, baseCount, refBase, minCoverage, threshold)
            if row is not None:
                out.append(row)
            del pileup[sweep]
            sweep += 1
    remainder = sorted(pileup.keys())
    for pos in remainder:
        baseCount = list(pileup[pos].items())
        refBase = ref_at(pos)
        row = None
        if flag_for(pos) == 1:
            row = output_candidate(ctgName, pos, baseCount, refBase, minCoverage, threshold)
        if row is not None:
            out.append(row)
    return out

"""CPU restatement of the reference's alignment pile-up (dataPrepScripts/CreateTensor.py) -- TEST INFRASTRUCTURE ONLY
(see oracle/cv_oracle.py header for who may import oracle/).

Follows the reference statement by statement with its own data structures -- `beginToEnd`, `endToCenter`, `activeSet`,
`centerToAln` lists of (refPos, queryAdv, refBase, queryBase) tuples, the 10,000,000-tuple `availableSlots` budget -- and
sums the tuples only at flush time in `generate_tensor`, like GenerateTensor (:23-59).  The native pile-up
(clairvoyante_b200/csrc/pileup.cpp) accumulates on the fly and is checked against this on the same SAM text.

Pinned against the reference itself: tests/golden/reference_run.npz holds the tensor rows the reference's own script printed
for three scenarios (defaults; region, MAPQ, depth cap and coverage options; considerleftedge off) when it was executed in
the build container with stand-ins for samtools / gzip (tests/golden/make_golden_reference_run.py); this restatement and the
native stage reproduce them -- same centres, same order, byte-identical "%0.1f" text (tests/test_reference_run_cpu.py).
The SAM rows are synthetic, in the column layout `samtools view` prints.

Restated with these stated differences (shared with the product): (1) centres flushed together are emitted in ascending
position order -- the reference iterates a Python-2 dict (:238, :248); (2) a reference / read index outside the supplied
strings reads as 'N' -- the reference raises IndexError, or wraps around for a negative index; (3) SAM rows with fewer than
ten fields are skipped.  Input plumbing (samtools faidx / view subprocesses, gzip, argparse: :96-146, :260-312) is replaced
by plain arguments.
"""
import re

import numpy as np

F = 16                                   # param.flankingBaseNum (dataPrepScripts/param.py:1)
MATRIX = 4                               # param.matrixNum
cigarRe = r"(\d+)([MIDNSHP=X])"          # CreateTensor.py:17
base2num = dict(zip("ACGT", (0, 1, 2, 3)))
stripe2 = 4 * MATRIX
stripe1 = MATRIX


def generate_tensor(alns, center, ref_start, min_coverage):
    """GenerateTensor (:23-59) -> (33,4,4) float32 counts or None"""
    alnCode = [0.0] * ((2 * F + 1) * 4 * MATRIX)
    depth = [0] * (2 * F + 1)
    for aln in alns:
        for refPos, queryAdv, refBase, queryBase in aln:
            if str(refBase) not in "ACGT-":
                continue
            if str(queryBase) not in "ACGT-":
                continue
            if refPos - center >= -(F + 1) and refPos - center < F:
                offset = refPos - center + (F + 1)
                if queryBase != "-":
                    if refBase != "-":
                        depth[offset] = depth[offset] + 1
                        alnCode[stripe2 * offset + stripe1 * base2num[refBase] + 0] += 1.0
                        alnCode[stripe2 * offset + stripe1 * base2num[queryBase] + 1] += 1.0
                        alnCode[stripe2 * offset + stripe1 * base2num[refBase] + 2] += 1.0
                        alnCode[stripe2 * offset + stripe1 * base2num[queryBase] + 3] += 1.0
                    elif refBase == "-":
                        idx = min(offset + queryAdv, 2 * F + 1 - 1)
                        alnCode[stripe2 * idx + stripe1 * base2num[queryBase] + 1] += 1.0
                elif queryBase == "-":
                    if refBase != "-":
                        alnCode[stripe2 * offset + stripe1 * base2num[refBase] + 2] += 1.0
    newRefPos = center - (0 if ref_start is None else (ref_start - 1))
    if (newRefPos - (F + 1) >= 0) and depth[F] >= min_coverage:
        return np.array(alnCode, dtype=np.float32).reshape(2 * F + 1, 4, MATRIX)
    return None


def create_tensors(sam_text, ref_seq, candidates, ref_start=None, min_mq=0, dcov=250, min_coverage=0,
                   consider_left_edge=True):
    """OutputAlnTensor's main loop (:148-252).  Returns [(center, tensor)] in emission order."""
    ref_off = 0 if ref_start is None else ref_start - 1

    def ref_at(p):
        i = p - ref_off
        return ref_seq[i] if 0 <= i < len(ref_seq) else "N"

    beginToEnd = {}
    for pos in candidates:                                   # GetCandidate (:61-83), eagerly
        if not consider_left_edge:
            beginToEnd[pos - (F + 1)] = [(pos + (F + 1), pos)]
        else:
            for i in range(pos - (F + 1), pos + (F + 1)):
                beginToEnd.setdefault(i, []).append((pos + (F + 1), pos))

    out = []
    availableSlots = 10000000
    centerToAln = {}
    previousPos = 0
    depthCap = 0
    for line in sam_text.split("\n"):
        l = line.split()
        if not l or l[0][0] == "@":
            continue
        if len(l) < 10:
            continue
        POS = int(l[3]) - 1
        MQ = int(l[4])
        CIGAR = l[5]
        SEQ = l[9]
        refPos = POS
        queryPos = 0
        if MQ < min_mq:
            continue
        endToCenter = {}
        activeSet = set()
        if previousPos != POS:
            previousPos = POS
            depthCap = 0
        else:
            depthCap += 1
            if depthCap >= dcov:
                continue

        def query(i):
            return SEQ[i] if 0 <= i < len(SEQ) else "N"

        for m in re.finditer(cigarRe, CIGAR):
            if availableSlots == 0:
                break
            advance = int(m.group(1))
            if m.group(2) == "S":
                queryPos += advance
            if m.group(2) in ("M", "=", "X"):
                for i in range(advance):
                    if refPos in beginToEnd:
                        for rEnd, rCenter in beginToEnd[refPos]:
                            if rCenter in activeSet:
                                continue
                            endToCenter[rEnd] = rCenter
                            activeSet.add(rCenter)
                            centerToAln.setdefault(rCenter, [])
                            centerToAln[rCenter].append([])
                    for center in list(activeSet):
                        if availableSlots != 0:
                            availableSlots -= 1
                            centerToAln[center][-1].append((refPos, 0, ref_at(refPos), query(queryPos)))
                    if refPos in endToCenter:
                        center = endToCenter[refPos]
                        activeSet.remove(center)
                    refPos += 1
                    queryPos += 1
            elif m.group(2) == "I":
                queryAdv = 0
                for i in range(advance):
                    for center in list(activeSet):
                        if availableSlots != 0:
                            availableSlots -= 1
                            centerToAln[center][-1].append((refPos, queryAdv, "-", query(queryPos)))
                    queryPos += 1
                    queryAdv += 1
            elif m.group(2) == "D":
                for i in range(advance):
                    for center in list(activeSet):
                        if availableSlots != 0:
                            availableSlots -= 1
                            centerToAln[center][-1].append((refPos, 0, ref_at(refPos), "-"))
                    if refPos in beginToEnd:
                        for rEnd, rCenter in beginToEnd[refPos]:
                            if rCenter in activeSet:
                                continue
                            endToCenter[rEnd] = rCenter
                            activeSet.add(rCenter)
                            centerToAln.setdefault(rCenter, [])
                            centerToAln[rCenter].append([])
                    if refPos in endToCenter:
                        center = endToCenter[refPos]
                        activeSet.remove(center)
                    refPos += 1

        if depthCap == 0:
            for center in sorted(centerToAln.keys()):          # (difference 1: sorted)
                if center + (F + 1) < POS:
                    t = generate_tensor(centerToAln[center], center, ref_start, min_coverage)
                    if t is not None:
                        out.append((center, t))
                    availableSlots += sum(len(i) for i in centerToAln[center])
                    del centerToAln[center]

    for center in sorted(centerToAln.keys()):
        t = generate_tensor(centerToAln[center], center, ref_start, min_coverage)
        if t is not None:
            out.append((center, t))
    return out


def tensor_line(ctg, center, ref_seq, ref_start, tensor):
    """the output row of GenerateTensor (:56)"""
    newRefPos = center - (0 if ref_start is None else (ref_start - 1))
    return "%s %d %s %s" % (ctg, center, ref_seq[newRefPos - (F + 1):newRefPos + F],
                            " ".join("%0.1f" % x for x in tensor.reshape(-1)))

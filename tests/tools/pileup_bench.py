"""Throughput of the native alignment pile-up (csrc/pileup.cpp) on this host beside the statement-by-statement Python
restatement of dataPrepScripts/CreateTensor.py (oracle/createtensor_oracle.py).
   python tests/tools/pileup_bench.py [ref_kb] [coverage] [bases per candidate]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from clairvoyante_b200 import CreateTensor as CT, ExtractVariantCandidates as EVC   # noqa: E402
from oracle import createtensor_oracle as O, candidates_oracle as OC                  # noqa: E402


def make(ref_len, coverage, read_len=150, seed=0, spacing=50):
    rng = np.random.default_rng(seed)
    ref = "".join(rng.choice(list("ACGT"), size=ref_len))
    n_reads = ref_len * coverage // read_len
    starts = np.sort(rng.integers(0, ref_len - read_len - 10, size=n_reads))
    rows = []
    snps = {int(q): "ACGT"[("ACGT".index(ref[int(q)]) + 1 + int(rng.integers(0, 3))) % 4]      # heterozygous SNP every ~300 bases
            for q in rng.integers(0, ref_len, size=ref_len // 300)}
    hap = "".join(snps.get(i, c) for i, c in enumerate(ref))
    for i, st in enumerate(starts.tolist()):
        r = rng.random()
        src = hap if rng.random() < 0.5 else ref
        if r < 0.85:
            cigar, seq = "%dM" % read_len, src[st:st + read_len]
        elif r < 0.93:
            a = int(rng.integers(20, 100)); k = int(rng.integers(1, 5))
            cigar, seq = "%dM%dI%dM" % (a, k, read_len - a - k), src[st:st + a] + "ACGT"[:k] + src[st + a:st + read_len - k]
        else:
            a = int(rng.integers(20, 100)); k = int(rng.integers(1, 5))
            cigar, seq = "%dM%dD%dM" % (a, k, read_len - a), src[st:st + a] + src[st + a + k:st + read_len + k]
        if rng.random() < 0.3:                              # a sequencing error
            j = int(rng.integers(0, len(seq))); seq = seq[:j] + "ACGT"[(("ACGT".index(seq[j]) + 1) % 4)] + seq[j + 1:]
        rows.append("r%d\t0\tctg\t%d\t60\t%s\t*\t0\t0\t%s\t*" % (i, st + 1, cigar, seq))
    cands = sorted(set(int(c) for c in rng.integers(20, ref_len - 20, size=ref_len // spacing)))
    return ref, "\n".join(rows) + "\n", cands, n_reads


def main():
    ref_kb = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    cov = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    spacing = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    ref, sam, cands, n_reads = make(ref_kb * 1000, cov, spacing=spacing)
    b = sam.encode()
    out = dict(host_cores=os.cpu_count(), ref_bases=len(ref), reads=n_reads, coverage=cov, candidates=len(cands), sam_bytes=len(b))
    best = None
    for _ in range(3):
        t = time.perf_counter()
        p = CT.Pileup(ref, cands, threads=1)
        p.feed(b, final=True)
        c, x = p.take()
        dt = time.perf_counter() - t
        p.close()
        best = dt if best is None else min(best, dt)
    nt = min(16, os.cpu_count() or 1)
    best_mt = None
    for _ in range(3):
        t = time.perf_counter()
        p = CT.Pileup(ref, cands, threads=nt)
        for i in range(0, len(b), 8 << 20):
            p.feed(b[i:i + (8 << 20)])
        p.feed(b"", final=True)
        cm, xm = p.take()
        dt = time.perf_counter() - t
        p.close()
        best_mt = dt if best_mt is None else min(best_mt, dt)
    assert np.array_equal(cm, c) and np.array_equal(xm, x)
    out["native_threads"] = dict(seconds=round(best_mt, 4), sites_per_s=round(len(c) / best_mt), reads_per_s=round(n_reads / best_mt),
                                 threads=nt)
    out["native"] = dict(seconds=round(best, 4), sites=len(c), sites_per_s=round(len(c) / best), reads_per_s=round(n_reads / best),
                         sam_mb_per_s=round(len(b) / best / 1e6, 1), threads=1)
    rows = 0
    t = time.perf_counter()
    for cc, xx in CT.pileup_tensors(b, ref, cands):
        rows += CT.tensor_text("ctg", cc, xx, ref).count(b"\n")
    out["native_with_text_rows"] = dict(sites_per_s=round(rows / (time.perf_counter() - t)))
    t = time.perf_counter()
    n = sum(g[1] for g in CT.GetTensorFromAlignments(b, ref, cands, "ctg", 1000))
    out["native_in_process_batches"] = dict(sites_per_s=round(n / (time.perf_counter() - t)))
    # the restatement on the first ~4 % of the reads
    sub = "\n".join(sam.split("\n")[:max(200, n_reads // 25)]) + "\n"
    last = int(sub.strip().split("\n")[-1].split("\t")[3])
    subc = [q for q in cands if q < last]
    t = time.perf_counter()
    w = O.create_tensors(sub, ref, subc)
    dt = time.perf_counter() - t
    out["python_restatement"] = dict(sites=len(w), sites_per_s=round(len(w) / dt), reads_per_s=round(sub.count("\n") / dt))
    out["speedup_sites_per_s"] = round(out["native"]["sites_per_s"] / max(out["python_restatement"]["sites_per_s"], 1), 1)
    # stage 1 of the pipeline: candidate extraction on the same reads
    best = None
    for _ in range(3):
        t = time.perf_counter()
        c = EVC.Candidates("ctg", ref, threads=1)
        c.feed(b, final=True)
        text, pos = c.take()
        dt = time.perf_counter() - t
        c.close()
        best = dt if best is None else min(best, dt)
    best_mt = None
    for _ in range(3):
        t = time.perf_counter()
        c = EVC.Candidates("ctg", ref, threads=nt)
        c.feed(b, final=True)
        text_mt, pos_mt = c.take()
        dt = time.perf_counter() - t
        c.close()
        best_mt = dt if best_mt is None else min(best_mt, dt)
    assert text_mt == text
    out["candidates_native_threads"] = dict(seconds=round(best_mt, 4), reads_per_s=round(n_reads / best_mt), threads=nt)
    t = time.perf_counter()
    w = OC.make_candidates(sub, "ctg", ref)
    dt = time.perf_counter() - t
    out["candidates_native"] = dict(seconds=round(best, 4), candidates=len(pos), reads_per_s=round(n_reads / best),
                                    sam_mb_per_s=round(len(b) / best / 1e6, 1), threads=1)
    out["candidates_python_restatement"] = dict(reads_per_s=round(sub.count("\n") / dt))
    out["candidates_speedup_reads_per_s"] = round(out["candidates_native"]["reads_per_s"] /
                                                  max(out["candidates_python_restatement"]["reads_per_s"], 1), 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

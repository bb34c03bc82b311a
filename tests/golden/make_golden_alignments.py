"""Generates tests/golden/alignments.npz: a small synthetic SAM text, its reference sequence, and what the ORACLE
restatements of the reference's two alignment stages produce from them -- the candidate rows of
dataPrepScripts/ExtractVariantCandidates.py (oracle/candidates_oracle.py) and the count tensors of
dataPrepScripts/CreateTensor.py at those candidates (oracle/createtensor_oracle.py).  Like forward_*.npz these are oracle
outputs (the reference needs Python 2 + samtools): they pin the oracles and the native stages against accidental change.
    python tests/golden/make_golden_alignments.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_pileup_cpu import synth_alignments        # noqa: E402
from oracle import candidates_oracle as OC, createtensor_oracle as OT   # noqa: E402

rng = np.random.default_rng(2024)
ref, sam, _ = synth_alignments(rng, ref_len=1500, n_reads=220)
rows = OC.make_candidates(sam, "ctg", ref, None, minCoverage=4, threshold=0.125)
pos = [int(r.split()[1]) for r in rows]
tens = OT.create_tensors(sam, ref, pos)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "alignments.npz"), sam=np.array(sam), ref=np.array(ref),
                    candidate_rows=np.array(rows), centers=np.array([c for c, _ in tens], np.int64),
                    tensors=np.stack([t for _, t in tens]).astype(np.int16))
print(len(rows), "candidate rows,", len(tens), "tensors written")

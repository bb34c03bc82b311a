"""GPU: callVarBam end to end on the command line -- alignments (.sam) + FASTA + checkpoint -> VCF -- against the same
chain assembled from its parts on the CPU: native candidates and pile-up (checked bit-exactly in the CPU suite), the NumPy
oracle forward pass, the per-site VCF restatement.  (Named to run after the kernel parity tests.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

from clairvoyante_b200 import CreateTensor as CT, ExtractVariantCandidates as EVC, initializers as I
from oracle import callvar_output as CO, cv_oracle as O

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_pileup_cpu import synth_alignments   # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("slim", [False, True])
def test_callvarbam_cli(tmp_path, slim):
    variant = "v3_slim" if slim else "v3"
    W = I.init_weights(variant, 4)
    rng = np.random.default_rng(31)
    ref, sam, _ = synth_alignments(rng, ref_len=6000, n_reads=1500, lower=0.0)
    # the pipeline sees a reference that differs from the reads' source at every 25th base: ~240 homozygous "variants"
    ref = "".join("ACGT"[("ACGT".index(c) + 1) % 4] if (i % 25 == 7 and c in "ACGT") else c for i, c in enumerate(ref))
    (tmp_path / "a.sam").write_text(sam)
    (tmp_path / "ref.fa").write_text(">ctg\n" + "\n".join(ref[i:i + 60] for i in range(0, len(ref), 60)) + "\n")
    ck, out = str(tmp_path / "model"), str(tmp_path / "o.vcf")
    np.savez(ck + ".cvb.npz", **W)
    cmd = [sys.executable, "-m", "clairvoyante_b200.callVarBam", "--chkpnt_fn", ck, "--bam_fn", str(tmp_path / "a.sam"), "--ref_fn",
           str(tmp_path / "ref.fa"), "--ctgName", "ctg", "--call_fn", out, "--sampleName", "HG001", "--samtools", "/nonexistent/samtools",
           "--delay", "0"] + (["--slim"] if slim else [])
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    body = [ln for ln in open(out).read().splitlines() if not ln.startswith("#")]
    # ---- the same chain from its parts
    c = EVC.Candidates("ctg", ref)
    c.feed(sam, final=True)
    _, pos = c.take()
    c.close()
    xs, ps = [], []
    for _, n, X, p in CT.GetTensorFromAlignments(sam, ref, pos, "ctg", 1000):
        xs.append(X[:n]); ps += p
    x = np.concatenate(xs)
    assert len(x) > 200
    o = O.forward(W, x, variant, dtype=np.float32)
    exp = [CO.vcf_line(x[j], ps[j], o["base"][j], o["zygosity"][j], o["varType"][j], o["indelLength"][j], False, None)
           for j in range(len(x))]
    exp = [e for e in exp if e is not None]
    got_by_pos = {b.split("\t")[1]: b for b in body}
    exp_by_pos = {e.split("\t")[1]: e for e in exp}
    # an oracle-side near-tie can flip a REF / non-REF decision (the site appears or disappears) or a field of a record
    assert len(set(got_by_pos) ^ set(exp_by_pos)) <= 0.03 * max(len(exp_by_pos), 1) + 2
    common = sorted(set(got_by_pos) & set(exp_by_pos), key=int)
    same = sum(got_by_pos[k] == exp_by_pos[k] for k in common)
    assert len(common) > 20 and same >= 0.95 * len(common), "%d of %d records identical" % (same, len(common))
    assert [b.split("\t")[1] for b in body] == sorted(got_by_pos, key=int)          # ascending, one record per site

"""CPU: the in-process callVarBam (candidates -> tensors -> model -> VCF, no text stream) writes the same VCF as the
reference's three-stage arrangement built from the stage CLIs and a tensor file (callVarBam.py:113-131), with a stub model
standing where the GPU model would be."""
import os
import subprocess
import sys
import types

import numpy as np
import pytest

from clairvoyante_b200 import callVar, callVarBam, param, utils_v2

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_pileup_cpu import synth_alignments       # noqa: E402
from test_callvar_cpu import _StubModel             # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _args(tmp_path, **kw):
    d = dict(chkpnt_fn=None, ref_fn=str(tmp_path / "ref.fa"), bed_fn=None, bam_fn=str(tmp_path / "a.sam"), call_fn=str(tmp_path / "o.vcf"),
             vcf_fn=None, threshold=0.125, minCoverage=4, qual=None, sampleName="S", ctgName="ctg", ctgStart=None, ctgEnd=None,
             considerleftedge=True, dcov=250, samtools="/nonexistent/samtools", pypy="pypy", v3=True, v2=False, slim=False,
             threads=None, delay=0)
    d.update(kw)
    return types.SimpleNamespace(**d)


@pytest.mark.parametrize("region", [None, (400, 2200)])
def test_in_process_pipeline_equals_staged_pipeline(tmp_path, monkeypatch, region):
    monkeypatch.setattr(param, "predictBatchSize", 16)
    rng = np.random.default_rng(21)
    ref, sam, _ = synth_alignments(rng, n_reads=500)
    (tmp_path / "a.sam").write_text(sam)
    (tmp_path / "ref.fa").write_text(">ctg\n" + "\n".join(ref[i:i + 60] for i in range(0, len(ref), 60)) + "\n")
    kw = dict(ctgStart=region[0], ctgEnd=region[1]) if region else {}
    n = callVarBam.Run(_args(tmp_path, **kw), model=_StubModel())
    assert n > 30
    got = [l for l in open(tmp_path / "o.vcf") if not l.startswith("#")]
    # ---- the staged arrangement: stage CLIs joined by files, then callVar.Test on the tensor file
    rng_args = ["--ctgStart", str(region[0]), "--ctgEnd", str(region[1])] if region else []
    common = ["--bam_fn", str(tmp_path / "a.sam"), "--ref_fn", str(tmp_path / "ref.fa"), "--ctgName", "ctg", "--samtools",
              "/nonexistent/samtools"] + rng_args
    r1 = subprocess.run([sys.executable, "-m", "clairvoyante_b200.ExtractVariantCandidates", "--can_fn", str(tmp_path / "can.gz")] + common,
                        cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r1.returncode == 0, r1.stderr
    r2 = subprocess.run([sys.executable, "-m", "clairvoyante_b200.CreateTensor", "--can_fn", str(tmp_path / "can.gz"), "--tensor_fn",
                         str(tmp_path / "t.gz")] + common, cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r2.returncode == 0, r2.stderr
    a2 = types.SimpleNamespace(tensor_fn=str(tmp_path / "t.gz"), call_fn=str(tmp_path / "o2.vcf"), showRef=False, qual=None, ref_fn=None,
                               sampleName="S")
    err, sys.stderr = sys.stderr, open(os.devnull, "w")
    try:
        callVar.Test(a2, _StubModel(), utils_v2)
    finally:
        sys.stderr = err
    want = [l for l in open(tmp_path / "o2.vcf") if not l.startswith("#")]
    assert got == want and len(got) > 5


def test_candidate_sites_from_vcf(tmp_path, monkeypatch):
    monkeypatch.setattr(param, "predictBatchSize", 16)
    rng = np.random.default_rng(22)
    ref, sam, cands = synth_alignments(rng, n_reads=300)
    (tmp_path / "a.sam").write_text(sam)
    (tmp_path / "ref.fa").write_text(">ctg\n" + ref + "\n")
    (tmp_path / "sites.vcf").write_text("##fileformat=VCFv4.1\n#CHROM\tPOS\n" + "".join("ctg\t%d\t.\tA\tC\t.\t.\t.\tGT\t0/1\n" % c for c in cands[:40])
                                        + "other\t7\t.\tA\tC\t.\t.\t.\tGT\t0/1\n")
    n = callVarBam.Run(_args(tmp_path, vcf_fn=str(tmp_path / "sites.vcf")), model=_StubModel())
    assert n == 40
    body = [l for l in open(tmp_path / "o.vcf") if not l.startswith("#")]
    assert set(int(l.split("\t")[1]) for l in body) <= set(cands[:40])


def test_parallel_command_generator(tmp_path):
    """callVarBamParallel.py:66-89: chunks of refChunkSize over the major contigs of the .fai, BED-less chunks skipped"""
    from clairvoyante_b200 import callVarBamParallel as P
    (tmp_path / "ref.fa").write_text(">chr1\nACGT\n")
    (tmp_path / "ref.fa.fai").write_text("chr1\t2500\t6\t60\t61\nchrUn_x\t900\t9\t60\t61\n21\t1000\t9\t60\t61\n")
    (tmp_path / "a.bam").write_text("")
    (tmp_path / "r.bed").write_text("chr1\t1200\t1300\n21\t0\t10\n")
    a = types.SimpleNamespace(chkpnt_fn="model", ref_fn=str(tmp_path / "ref.fa"), bed_fn=None, refChunkSize=1000, bam_fn=str(tmp_path / "a.bam"),
                              vcf_fn=None, output_prefix="out/p", includingAllContigs=False, threshold=0.2, minCoverage=4, qual=None,
                              sampleName="S", considerleftedge=True, samtools="samtools", slim=False)
    cmds = P.commands(a)
    regions = [(c.split("--ctgName ")[1].split()[0], int(c.split("--ctgStart ")[1].split()[0]), int(c.split("--ctgEnd ")[1].split()[0])) for c in cmds]
    assert regions == [("chr1", 0, 1000), ("chr1", 1000, 2000), ("chr1", 2000, 2500), ("21", 0, 1000)]
    assert all("--call_fn out/p.%s_%d_%d.vcf" % r in c for r, c in zip(regions, cmds))
    a.bed_fn = str(tmp_path / "r.bed")
    a.includingAllContigs = True
    cmds = P.commands(a)
    assert [c.split("--ctgStart ")[1].split()[0] for c in cmds] == ["1000", "0"] and all("--bed_fn" in c for c in cmds)

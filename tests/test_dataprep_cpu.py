"""CPU: the two small data-preparation drivers around the native stages -- GetTruth (truth VCF -> variant rows) and
PairWithNonVariants (variant tensors + sampled non-variant tensors)."""
import gzip
import types

from clairvoyante_b200 import GetTruth, PairWithNonVariants


VCF = """##fileformat=VCFv4.1
#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS
chr1\t100\t.\tA\tG\t50\tPASS\t.\tGT:GQ\t0/1:99
chr1\t200\t.\tC\tT\t50\tPASS\t.\tGT\t1|1
chr1\t300\t.\tG\tGAC,GA\t50\tPASS\t.\tGT\t1/2
chr1\t400\t.\tTAC\tT\t50\tPASS\t.\tGT\t1|0
chr1\t500\t.\tA\tC\t50\tPASS\t.\tGT\t./1
chr2\t100\t.\tA\tG\t50\tPASS\t.\tGT\t0/1
"""


def test_get_truth_rows(tmp_path):
    rows = list(GetTruth.truth_rows(VCF.split("\n"), "chr1"))
    assert rows == ["chr1 100 A G 0 1", "chr1 200 C T 1 1", "chr1 300 G GA 0 1",      # 1/2 with two ALTs -> 0/1, shortest ALT
                    "chr1 400 TAC T 0 1", "chr1 500 A C 0 1"]                          # alleles ordered; '.' counts as 0
    # region: the reference compares POS with (ctgStart + 1, ctgEnd)
    (tmp_path / "t.vcf").write_text(VCF)
    a = types.SimpleNamespace(vcf_fn=str(tmp_path / "t.vcf"), var_fn=str(tmp_path / "v.gz"), ctgName="chr1", ctgStart=199, ctgEnd=400)
    GetTruth.OutputVariant(a)
    assert gzip.open(a.var_fn, "rt").read().split("\n")[:-1] == ["chr1 200 C T 1 1", "chr1 300 G GA 0 1", "chr1 400 TAC T 0 1"]
    a.ctgStart = 200                                                                    # 200 + 1 = 201 > 200: first row drops
    GetTruth.OutputVariant(a)
    assert gzip.open(a.var_fn, "rt").read().split("\n")[:-1] == ["chr1 300 G GA 0 1", "chr1 400 TAC T 0 1"]


def test_pair_with_non_variants(tmp_path):
    var = ["chr1 %d SEQ 1.0 2.0" % p for p in (100, 200)]
    can = ["chr1 %d SEQ 0.0 0.0" % p for p in (100, 150, 160, 170, 180, 190, 250, 900)] + ["chr2 5 SEQ 0.0"]
    (tmp_path / "var.txt").write_text("\n".join(var) + "\n")
    with gzip.open(tmp_path / "can.gz", "wt") as f:
        f.write("\n".join(can) + "\n")
    (tmp_path / "r.bed").write_text("chr1\t0\t300\n")
    a = types.SimpleNamespace(tensor_can_fn=str(tmp_path / "can.gz"), tensor_var_fn=str(tmp_path / "var.txt"), bed_fn=str(tmp_path / "r.bed"),
                              output_fn=str(tmp_path / "o.gz"), amp=10, seed=1)
    o1, o2 = PairWithNonVariants.Pair(a)                       # amp * 2 = 20 >= 6 usable -> every usable non-variant is kept
    out = gzip.open(a.output_fn, "rt").read().split("\n")[:-1]
    assert (o1, o2) == (2, 6) and out == var + [c for c in can if c.split()[1] in ("150", "160", "170", "180", "190", "250")]
    a.amp = 1                                                   # 2 of 6 on average; seeded -> reproducible, variants always first
    o1, o2 = PairWithNonVariants.Pair(a)
    out1 = gzip.open(a.output_fn, "rt").read().split("\n")[:-1]
    PairWithNonVariants.Pair(a)
    assert out1 == gzip.open(a.output_fn, "rt").read().split("\n")[:-1] and out1[:2] == var and 0 <= o2 <= 6
    a.bed_fn = None                                             # without BED the chr1:900 and chr2 rows are usable too
    a.amp = 100
    assert PairWithNonVariants.Pair(a) == (2, 8)


def test_submodule_invocator(tmp_path):
    """python -m clairvoyante_b200 <Submodule> ... (reference clairvoyante.py) reaches the submodule's main()"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    vcf = tmp_path / "t.vcf"
    vcf.write_text("#h\nctg\t10\t.\tA\tC\t50\tPASS\t.\tGT\t0/1\nctg\t20\t.\tAT\tA\t50\tPASS\t.\tGT\t1|1\n")
    r = subprocess.run([sys.executable, "-m", "clairvoyante_b200", "GetTruth", "--vcf_fn", str(vcf), "--ctgName", "ctg"], cwd=root,
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "ctg 10 A C 0 1\nctg 20 AT A 1 1\n", r.stderr
    r = subprocess.run([sys.executable, "-m", "clairvoyante_b200"], cwd=root, capture_output=True, text=True)
    assert r.returncode == 0 and "callVarBam" in r.stdout
    r = subprocess.run([sys.executable, "-m", "clairvoyante_b200", "getEmbedding"], cwd=root, capture_output=True, text=True)
    assert r.returncode != 0 and "outside the scope" in r.stderr

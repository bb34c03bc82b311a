"""The two alignment stages against the REFERENCE'S OWN CODE: tests/golden/reference_run.npz holds what
dataPrepScripts/ExtractVariantCandidates.py and dataPrepScripts/CreateTensor.py of the reference printed for three synthetic
scenarios when they were executed in the build container (tests/golden/make_golden_reference_run.py lists the mechanical
Python 2 -> 3 rewrites and the samtools / gzip / intervaltree stand-ins that run needed).  Checked here, without the
reference: the package's command lines (native C++ stages) and the oracle restatements reproduce those outputs."""
import hashlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import candidates_oracle as OC, createtensor_oracle as OT   # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "reference_run.npz"))
SCENARIOS = [str(s) for s in G["scenarios"]]


def _opts(argv):
    return {argv[i].lstrip("-"): argv[i + 1] for i in range(0, len(argv), 2)}


def _files(tmp_path, n):
    ref, sam, bed = str(G[n + "/ref"]), str(G[n + "/sam"]), str(G[n + "/bed"])
    fa, samfn, bedfn = (str(tmp_path / (n + e)) for e in (".fa", ".sam", ".bed"))
    open(fa, "w").write(">ctg\n" + "".join(ref[i:i + 70] + "\n" for i in range(0, len(ref), 70)))
    open(samfn, "w").write(sam)
    if bed:
        open(bedfn, "w").write(bed)
    return fa, samfn, (bedfn if bed else None)


def _view(sam, ctg, a=None, b=None):
    """`samtools view -F 2308 file ctg[:a-b]` on SAM text, as the generator emulated it"""
    out = []
    for line in sam.split("\n"):
        if not line or line.startswith("@"):
            continue
        f = line.split("\t")
        if f[2] != ctg or int(f[1]) & 2308:
            continue
        lo = int(f[3])
        hi = lo + max(sum(int(k) for k, op in re.findall(r"(\d+)([MIDNSHP=X])", f[5]) if op in "MDN=X"), 1) - 1
        if a is not None and (hi < a or lo > b):
            continue
        out.append(line)
    return "\n".join(out) + "\n"


def test_counter_order_was_computed_not_assumed():
    assert [str(k) for k in G["py27_counter_order"]] == list(OC.KEY_ORDER)


@pytest.mark.parametrize("n", SCENARIOS)
def test_candidate_rows_equal_the_reference_run(tmp_path, n):
    want = str(G[n + "/candidate_rows"])
    assert want.count("\n") > 20
    fa, samfn, bedfn = _files(tmp_path, n)
    argv = ["--bam_fn", samfn, "--ref_fn", fa, "--ctgName", "ctg"] + [str(a) for a in G[n + "/can_args"]]
    if bedfn:
        argv += ["--bed_fn", bedfn]
    r = subprocess.run([sys.executable, "-m", "clairvoyante_b200.ExtractVariantCandidates"] + argv, cwd=ROOT, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert r.stdout.decode() == want                                   # native stage, same command line
    o = _opts([str(a) for a in G[n + "/can_args"]])
    region = "ctgStart" in o
    cs, ce = (int(o["ctgStart"]) + 1, int(o["ctgEnd"])) if region else (None, None)
    bed = None
    if bedfn:
        bed = []
        for row in str(G[n + "/bed"]).split("\n"):
            row = row.split()
            if row and row[0] == "ctg":
                b, e = int(row[1]), int(row[2]) - 1
                bed.append((b, e + 1 if e == b else e))
    rows = OC.make_candidates(_view(str(G[n + "/sam"]), "ctg", cs, ce), "ctg", str(G[n + "/ref"]), 1 if region else None, cs, ce, bed,
                              minMQ=int(o.get("minMQ", 0)), minCoverage=float(o.get("minCoverage", 4)),
                              threshold=float(o.get("threshold", 0.125)))
    assert "".join(r + "\n" for r in rows) == want                     # oracle restatement


@pytest.mark.parametrize("n", SCENARIOS)
def test_python3_dict_order_changes_tie_rows_only(n):
    a = str(G[n + "/candidate_rows"]).split("\n")
    b = str(G[n + "/candidate_rows_py3_dict_order"]).split("\n")
    assert len(a) == len(b)
    differing = 0
    for ra, rb in zip(a, b):
        if ra == rb:
            continue
        differing += 1
        fa_, fb_ = ra.split(), rb.split()
        assert fa_[:4] == fb_[:4]                                        # same site, same total
        pa, pb = list(zip(fa_[4::2], fa_[5::2])), list(zip(fb_[4::2], fb_[5::2]))
        assert sorted(pa) == sorted(pb) and [c for _, c in pa] == [c for _, c in pb]   # same counts, only tied keys permuted
    assert differing > 0


@pytest.mark.parametrize("n", SCENARIOS)
def test_tensors_equal_the_reference_run(tmp_path, n):
    head = [str(h) for h in G[n + "/tensor_head"]]
    want = G[n + "/tensors"].astype(np.float32)
    fa, samfn, _ = _files(tmp_path, n)
    canfn = str(tmp_path / "can.txt")
    open(canfn, "w").write(str(G[n + "/candidate_rows"]))
    ten_args = [str(a) for a in G[n + "/ten_args"]]
    argv = ["--bam_fn", samfn, "--ref_fn", fa, "--ctgName", "ctg", "--can_fn", canfn] + ten_args
    r = subprocess.run([sys.executable, "-m", "clairvoyante_b200.CreateTensor"] + argv, cwd=ROOT, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    text = r.stdout.decode()
    lines = [l for l in text.split("\n") if l]
    assert [" ".join(l.split()[:3]) for l in lines] == head            # same sites, same order, same 33-base context
    got = np.array([[float(v) for v in l.split()[3:]] for l in lines], np.float32).reshape(-1, 33, 4, 4)
    assert np.array_equal(got, want)
    assert hashlib.sha256(text.encode()).hexdigest() == str(G[n + "/tensor_text_sha256"])   # byte-identical rows ("%0.1f")
    o = _opts(ten_args)
    region = "ctgStart" in o
    cs, ce = (int(o["ctgStart"]) + 1, int(o["ctgEnd"])) if region else (None, None)
    cands = [int(l.split()[1]) for l in str(G[n + "/candidate_rows"]).split("\n") if l]
    cands = [p for p in cands if not region or cs <= p <= ce]
    lee = o.get("considerleftedge", "True") == "True"
    res = OT.create_tensors(_view(str(G[n + "/sam"]), "ctg", cs, ce), str(G[n + "/ref"]), cands, 1 if region else None,
                            min_mq=int(o.get("minMQ", 0)), dcov=int(o.get("dcov", 250)), min_coverage=int(o.get("minCoverage", 0)),
                            consider_left_edge=lee)
    assert [c for c, _ in res] == [int(h.split()[1]) for h in head]
    assert np.array_equal(np.stack([t for _, t in res]).astype(np.float32), want) if len(res) else len(want) == 0


# ---------------------------------------------------------------- feed (utils_v2.py) and VCF writer (callVar.py) ----------
from clairvoyante_b200 import callVar as CV, param as P, utils_v2 as U   # noqa: E402


@pytest.fixture()
def tensor_file(tmp_path):
    fn = str(tmp_path / "tensors.txt")
    open(fn, "wb").write(G["feed/tensor_text"].tobytes())
    return fn


def test_gettensor_equals_the_reference_run(tensor_file):
    ends, counts, xs, pos = [], [], [], []
    for end, c, x, p in U.GetTensor(tensor_file, 100):
        ends.append(end); counts.append(c); xs.append(np.array(x[:c])); pos += list(p)
    assert ends == G["feed/gettensor_end"].tolist() and counts == G["feed/gettensor_count"].tolist()
    assert np.array_equal(np.concatenate(xs), G["feed/gettensor_x"].astype(np.float32))
    assert pos == [str(s) for s in G["feed/gettensor_pos"]]


@pytest.mark.parametrize("tag", ["nobed", "bed"])
def test_gettrainingarray_equals_the_reference_run(tmp_path, tensor_file, tag):
    vfn, bfn = str(tmp_path / "vars.txt"), str(tmp_path / "conf.bed")
    open(vfn, "w").write(str(G["feed/var_text"]))
    open(bfn, "w").write(str(G["feed/bed_text"]))
    total, XC, YC, PC = U.GetTrainingArray(tensor_file, vfn, bfn if tag == "bed" else None, shuffle=False)
    assert total == int(G["feed/train_%s_total" % tag])
    X = [U.unpack_array(b) for b in XC]
    assert [len(x) for x in X] == G["feed/train_%s_blocks" % tag].tolist()        # 500-row blocks + the trailing one
    assert np.array_equal(np.concatenate([x.reshape(-1, 33, 4, 4) for x in X]), G["feed/train_%s_x" % tag].astype(np.float32))
    Y = np.concatenate([np.asarray(U.unpack_array(b)).reshape(-1, 16) for b in YC])
    assert Y.dtype == np.float64 and np.array_equal(Y, G["feed/train_%s_y" % tag])
    pos = np.concatenate([np.asarray(U.unpack_array(b)).reshape(-1) for b in PC]).astype(str)
    assert pos.tolist() == G["feed/train_%s_pos" % tag].tolist()
    if tag == "nobed":
        for k, (st, num) in enumerate(G["feed/decompress_cases"].tolist()):
            a, n, e = U.DecompressArray(YC, st, num, total)
            assert [n, e] == G["feed/decompress_%d_ne" % k].tolist()
            assert np.array_equal(a, G["feed/decompress_%d" % k])


class _TableModel(object):
    def __init__(self, table):
        self.table, self.at = table, 0

    def predictNoRT(self, X):
        p = self.table[self.at:self.at + len(X)]
        assert len(p) == len(X)
        self.at += len(X)
        self.predictBaseRTVal, self.predictZygosityRTVal = p[:, 0:4], p[:, 4:6]
        self.predictVarTypeRTVal, self.predictIndelLengthRTVal = p[:, 6:10], p[:, 10:16]


@pytest.mark.parametrize("tag,kw", [("default", dict(qual=None, showRef=False, ref_fn=None)),
                                    ("showref_qual", dict(qual=20, showRef=True, ref_fn=None)),
                                    ("contigs", dict(qual=150, showRef=False, ref_fn="ref.fa"))])
def test_vcf_equals_the_reference_run(tmp_path, tensor_file, monkeypatch, tag, kw):
    import types
    want = str(G["callvar/vcf_" + tag])
    assert want.count("\n") > 400
    if kw["ref_fn"]:
        kw = dict(kw, ref_fn=str(tmp_path / "ref.fa"))
        open(kw["ref_fn"] + ".fai", "w").write("ctg\t2200\t5\t70\t71\nother\t100\t2300\t70\t71\n")
    monkeypatch.setattr(P, "predictBatchSize", 100)
    out = str(tmp_path / "calls.vcf")
    args = types.SimpleNamespace(v2=False, v3=True, slim=False, tensor_fn=tensor_file, call_fn=out, sampleName="SAMPLE", threads=None, **kw)
    CV.Test(args, _TableModel(G["callvar/probabilities"]), U)
    assert open(out).read() == want


# ---------------------------------------------------------------- the remaining oracles against the same fixture ----------
def test_feed_oracle_equals_the_reference_run(tensor_file):
    from oracle import feed_oracle as FO
    ends, counts, xs, pos = [], [], [], []
    for end, c, x, p in FO.GetTensor(tensor_file, 100):
        ends.append(end); counts.append(c); xs.append(np.array(x[:c])); pos += list(p)
    assert ends == G["feed/gettensor_end"].tolist() and counts == G["feed/gettensor_count"].tolist()
    assert np.array_equal(np.concatenate(xs), G["feed/gettensor_x"].astype(np.float32))
    assert pos == [str(s) for s in G["feed/gettensor_pos"]]


@pytest.mark.parametrize("tag,show_ref,qual", [("default", False, None), ("showref_qual", True, 20), ("contigs", False, 150)])
def test_vcf_oracle_equals_the_reference_run(tag, show_ref, qual):
    from oracle import callvar_output as CO
    x = G["feed/gettensor_x"].astype(np.float32)
    pos = [str(s) for s in G["feed/gettensor_pos"]]
    p = G["callvar/probabilities"]
    got = [CO.vcf_line(x[j], pos[j], p[j, 0:4], p[j, 4:6], p[j, 6:10], p[j, 10:16], show_ref, qual) for j in range(len(x))]
    want = [l for l in str(G["callvar/vcf_" + tag]).split("\n") if l and not l.startswith("#")]
    assert [g for g in got if g is not None] == want


# ---------------------------------------------------------------- train.py's TrainAll with a stub model ------------------
class _StubTrainer(object):
    """same stand-in as tests/golden/make_golden_reference_run.py StubTrainer"""

    def __init__(self):
        self.calls, self.epoch, self.lr, self.l2 = [], 1, None, None

    @staticmethod
    def _cs(X, Y=None):
        return round(float(np.asarray(X, np.float64).sum()) + (float(np.asarray(Y, np.float64).sum()) if Y is not None else 0.0), 3)

    def _val(self, n):
        return n * (1.0 + (0.25 if self.epoch % 2 else -0.25) + 0.001 * self.epoch)

    def trainNoRT(self, X, Y):
        self.calls.append(("trainNoRT", len(X), self._cs(X, Y)))
        self.trainLossRTVal, self.trainSummaryRTVal = float(np.abs(np.asarray(X, np.float64)).sum()) * 1e-3 / self.epoch, None

    def getLossNoRT(self, X, Y):
        self.calls.append(("getLossNoRT", len(X), self._cs(X, Y)))
        self.getLossLossRTVal = self._val(len(X))

    def getLoss(self, X, Y):
        self.calls.append(("getLoss", len(X), self._cs(X, Y)))
        return self._val(len(X))

    def setLearningRate(self, v=None):
        self.lr = self.lr * 0.1 if v is None else v
        self.calls.append(("setLearningRate", v, self.lr))
        return self.lr

    def setL2RegularizationLambda(self, v=None):
        self.l2 = self.l2 * 0.1 if v is None else v
        self.calls.append(("setL2RegularizationLambda", v, self.l2))
        return self.l2

    def saveParameters(self, path):
        self.calls.append(("saveParameters", os.path.basename(path)))
        self.epoch += 1

    def restoreParameters(self, path):
        self.calls.append(("restoreParameters", os.path.basename(path)))
        self.epoch = int(path[-6:])

    def predict(self, X):
        self.calls.append(("predict", len(X), self._cs(X)))
        f = np.asarray(X, np.float32).reshape(len(X), -1)
        return f[:, 256:260], f[:, 260:262], f[:, 262:266], f[:, 266:272]


DRIVERS = [   # as in tests/golden/make_golden_reference_run.py
    ("train", "TrainAll", dict(trainBatchSize=200, predictBatchSize=50), {}),
    ("trainNonstop", "TrainAll", dict(trainBatchSize=200, predictBatchSize=50, maxEpoch=5), {}),
    ("trainWithoutValidationNonstop", "TrainAll", dict(trainBatchSize=150, predictBatchSize=50, maxEpoch=4), {}),
    ("evaluate", "Test", dict(predictBatchSize=70), {}),
    ("calTrainDevDiff", "CalcAll", dict(predictBatchSize=60), dict(chkpnt_fn=["model-000003", "model-000008"])),
]


@pytest.mark.parametrize("name,entry,overrides,extra", DRIVERS, ids=[d[0] for d in DRIVERS])
def test_driver_equals_the_reference_run(tmp_path, monkeypatch, capsys, name, entry, overrides, extra):
    """batch booking, validation split, learning-rate switches, checkpoint names, final evaluation: the same model calls in the
    same order on the same rows, and the same log lines, as the reference's own driver produced with this stub model"""
    import importlib
    import logging
    import pickle
    import types
    mod = importlib.import_module("clairvoyante_b200." + name)
    X = G["feed/train_nobed_x"].astype(np.float32)
    Y = G["feed/train_nobed_y"]
    pos = G["feed/train_nobed_pos"].astype("S")
    total = int(G["feed/train_nobed_total"])
    bs = P.bloscBlockSize
    bin_fn = str(tmp_path / "train.bin")
    with open(bin_fn, "wb") as fh:
        for obj in (total, [U.pack_array(X[i:i + bs]) for i in range(0, total, bs)], [U.pack_array(Y[i:i + bs]) for i in range(0, total, bs)],
                    [U.pack_array(pos[i:i + bs]) for i in range(0, total, bs)]):
            pickle.dump(obj, fh)
    for k, v in overrides.items():
        monkeypatch.setattr(P, k, v)
    msgs = []

    class Collect(logging.Handler):
        def emit(self, record):
            msgs.append(record.getMessage())

    h = Collect()
    logging.getLogger().addHandler(h)
    old_level = logging.getLogger().level
    logging.getLogger().setLevel(logging.INFO)
    try:
        m = _StubTrainer()
        kw = dict(bin_fn=bin_fn, tensor_fn=None, var_fn=None, bed_fn=None, chkpnt_fn=None, learning_rate=1e-3, lambd=1e-3,
                  ochk_prefix=str(tmp_path / "model"), olog_dir=None, v2=False, v3=True, slim=False)
        kw.update(extra)
        capsys.readouterr()
        getattr(mod, entry)(types.SimpleNamespace(**kw), m, U)
        err = capsys.readouterr().err
    finally:
        logging.getLogger().removeHandler(h)
        logging.getLogger().setLevel(old_level)
    assert [repr(c) for c in m.calls] == [str(c) for c in G[name + "/calls"]]
    got = [x for x in msgs if "time elapsed" not in x] + [l for l in err.split("\n") if l and l not in msgs]
    assert got == [str(x) for x in G[name + "/log"]]


# ---------------------------------------------------------------- GetTruth, PairWithNonVariants, callVarBamParallel --------
@pytest.mark.parametrize("tag,extra", [("all", []), ("region", ["--ctgStart", "400", "--ctgEnd", "1500"])])
def test_gettruth_equals_the_reference_run(tmp_path, tag, extra):
    vfn = str(tmp_path / "truth.vcf")
    open(vfn, "w").write(str(G["gettruth/vcf"]))
    r = subprocess.run([sys.executable, "-m", "clairvoyante_b200.GetTruth", "--vcf_fn", vfn, "--ctgName", "ctg"] + extra, cwd=ROOT,
                       capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    want = str(G["gettruth/" + tag])
    assert want.count("\n") > 20 and r.stdout.decode() == want


@pytest.mark.parametrize("tag,amp", [("nobed", 2), ("bed", 1)])
def test_pairwithnonvariants_equals_the_reference_run(tmp_path, tag, amp):
    import gzip
    import random
    import types
    from clairvoyante_b200 import PairWithNonVariants as PW
    lines = [l for l in G["feed/tensor_text"].tobytes().decode().split("\n") if l]
    tv, tc, bed, outfn = (str(tmp_path / n) for n in ("tensor_var", "tensor_can", "pair.bed", "tensor_pair"))
    open(tv, "w").write("".join(l + "\n" for l in lines[:int(G["pair/n_var"])]))
    open(tc, "w").write("".join(l + "\n" for l in lines[int(G["pair/n_can_from"]):]))
    open(bed, "w").write("ctg\t0\t1200\nctg\t1500\t700000\n")
    random.seed(12345)                       # the generator seeded the reference's (otherwise unseeded) sampler the same way
    PW.Pair(types.SimpleNamespace(tensor_can_fn=tc, tensor_var_fn=tv, bed_fn=bed if tag == "bed" else None, output_fn=outfn, amp=amp,
                                  seed=None))
    got = gzip.open(outfn, "rt").read()
    assert [" ".join(l.split()[:2]) for l in got.split("\n") if l] == [str(p) for p in G["pair/%s_positions" % tag]]
    assert hashlib.sha256(got.encode()).hexdigest() == str(G["pair/%s_sha256" % tag])


def _options(cmd):
    tok, o, i = cmd.split(), {}, 0
    while i < len(tok):
        if tok[i].startswith("--"):
            if i + 1 < len(tok) and not tok[i + 1].startswith("--"):
                o[tok[i]] = tok[i + 1]
                i += 2
                continue
            o[tok[i]] = True
        i += 1
    return o


@pytest.mark.parametrize("tag", ["default", "bed_qual", "allcontigs"])
def test_callvarbamparallel_chunks_equal_the_reference_run(tmp_path, tag):
    tmp = str(tmp_path)
    fa = os.path.join(tmp, "genome.fa")
    for fn in (fa, os.path.join(tmp, "aln.bam"), os.path.join(tmp, "model.meta"), os.path.join(tmp, "model.index")):
        open(fn, "w").write("x\n")
    open(fa + ".fai", "w").write("chr1\t25000000\t6\t60\t61\nchr2\t10000001\t7\t60\t61\nchrUn_x\t5000\t8\t60\t61\n21\t9999999\t9\t60\t61\n")
    open(os.path.join(tmp, "chunks.bed"), "w").write("chr1\t10000000\t10000001\nchr1\t19999999\t20000500\n21\t5\t500\n")
    extra = [str(a).replace("TMP", tmp) for a in G["parallel/%s_args" % tag]]
    extra = [os.path.join(tmp, "chunks.bed") if a.endswith("chunks.bed") else a for a in extra]
    r = subprocess.run([sys.executable, "-m", "clairvoyante_b200.callVarBamParallel", "--chkpnt_fn", os.path.join(tmp, "model"), "--ref_fn", fa,
                        "--bam_fn", os.path.join(tmp, "aln.bam"), "--pypy", "python", "--samtools", "python", "--output_prefix", "out/calls"]
                       + extra, cwd=ROOT, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    got = [_options(c.replace(tmp, "TMP")) for c in r.stdout.decode().split("\n") if c]
    want = [_options(c) for c in str(G["parallel/" + tag]).split("\n") if c]
    assert len(got) == len(want) and len(want) >= 4
    same = ["--chkpnt_fn", "--ref_fn", "--bam_fn", "--bed_fn", "--ctgName", "--ctgStart", "--ctgEnd", "--call_fn", "--sampleName", "--qual",
            "--considerleftedge"]
    for g, w in zip(got, want):
        assert [g.get(k) for k in same] == [w.get(k) for k in same]
        assert float(g["--threshold"]) == float(w["--threshold"]) and float(g["--minCoverage"]) == float(w["--minCoverage"])


# ---------------------------------------------------------------- the whole calling pipeline --------------------------------
PIPELINES = [   # as in tests/golden/make_golden_reference_run.py
    ("plain", "defaults", {}),
    ("region_bed_qual", "region_bed_filters", dict(ctgStart=150, ctgEnd=1900, bed=True, threshold=0.1, minCoverage=6, qual=30, dcov=5)),
    ("vcf_sites", "defaults", dict(vcf=True, ctgStart=300, ctgEnd=1400)),
]


@pytest.mark.parametrize("cache_mb", ["2048", "0"], ids=["text-kept-for-pass-2", "two-passes-over-the-file"])
@pytest.mark.parametrize("tag,sc,o", PIPELINES, ids=[p[0] for p in PIPELINES])
def test_callvarbam_equals_the_reference_pipeline(tmp_path, monkeypatch, tag, sc, o, cache_mb):
    """alignments -> VCF in one process (native candidates, native pile-up, callVar.Test) against the VCF that the reference's
    three stages, chained with the options callVarBam.py:113-131 gives them, wrote with the same probability table as model"""
    import types
    from clairvoyante_b200 import callVarBam as CB
    fa, samfn, bedfn, vfn, out = (str(tmp_path / n) for n in ("ref.fa", "aln.sam", "conf.bed", "sites.vcf", "calls.vcf"))
    ref = str(G[sc + "/ref"])
    open(fa, "w").write(">ctg\n" + "".join(ref[i:i + 70] + "\n" for i in range(0, len(ref), 70)))
    open(fa + ".fai", "w").write("ctg\t%d\t5\t70\t71\n" % len(ref))
    open(samfn, "w").write(str(G[sc + "/sam"]))
    if o.get("bed"):
        open(bedfn, "w").write(str(G[sc + "/bed"]))
    if o.get("vcf"):
        open(vfn, "w").write(str(G["gettruth/vcf"]))
    monkeypatch.setattr(P, "predictBatchSize", 100)
    monkeypatch.setenv("CVB_SAM_CACHE_MB", cache_mb)
    args = types.SimpleNamespace(chkpnt_fn=None, ref_fn=fa, bed_fn=bedfn if o.get("bed") else None, bam_fn=samfn, call_fn=out,
                                 vcf_fn=vfn if o.get("vcf") else None, threshold=o.get("threshold", 0.125),
                                 minCoverage=o.get("minCoverage", 4), qual=o.get("qual"), sampleName="SAMPLE", ctgName="ctg",
                                 ctgStart=o.get("ctgStart"), ctgEnd=o.get("ctgEnd"), considerleftedge=True, dcov=o.get("dcov", 250),
                                 samtools="samtools-is-not-installed", pypy="pypy", v3=True, v2=False, slim=False, threads=None, delay=0)
    CB.Run(args, model=_TableModel(G["pipeline/probabilities"]))
    want = str(G["pipeline/%s_vcf" % tag])
    assert want.count("\n") > 30 and open(out).read() == want

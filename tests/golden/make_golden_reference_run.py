"""Generates tests/golden/reference_run.npz by RUNNING THE REFERENCE'S OWN CODE for the two alignment stages,
dataPrepScripts/ExtractVariantCandidates.py and dataPrepScripts/CreateTensor.py, read from /root/reference at generation
time (nothing of it is copied into this repository; the GPU box never needs it).  The fixture pins
oracle/candidates_oracle.py, oracle/createtensor_oracle.py and the native stages (csrc/candidates.cpp, csrc/pileup.cpp)
against the reference itself -- see tests/test_reference_run_cpu.py.

The reference is Python 2 and shells out to samtools / gzip; neither exists here.  What this script does about that, in full:

  * mechanical Python 2 -> 3 rewrites of the source TEXT before exec (REWRITES below, each one counted and asserted):
    `print >> sys.stderr, x` -> print(x, file=sys.stderr); xrange -> range; `d.items()` / `d.keys()` that are sorted or
    mutated while iterated -> list(...).  No statement of the algorithms is touched.
  * the counter literal {"A":0,"C":0,"G":0,"T":0,"I":0,"D":0,"N":0}: its iteration order decides ties in the reference's
    stable sort.  Python 3 iterates in insertion order, CPython 2.7 in hash-slot order.  The 2.7 order is COMPUTED here by
    replaying CPython 2.7's dict insertion (string hash, 8-slot presize, growth to 32 slots on the 6th insert, perturbed
    probing; py27_dict_order below) and the literal is rewritten to insert in that order.  The fixture also keeps the rows
    produced with the written (Python 3) order so that the test can show the two differ in tie rows only.
  * `import intervaltree` (third-party, absent): a brute-force stand-in with the two calls the script makes (addi, search).
  * `subprocess.Popen`: replaced inside the executed module by an emulation of exactly the three commands the scripts run:
    `samtools faidx ref.fa ctg[:a-b]` on a FASTA text file, `samtools view -F 2308 aln.sam ctg[:a-b]` on a SAM TEXT file
    (records in file order, header dropped, flag filter, 1-based inclusive overlap with the region), `gzip -fdc file` on a
    plain text file.  Outputs are taken from the PIPE (stdout) mode of both scripts.

    python tests/golden/make_golden_reference_run.py
"""
import contextlib
import hashlib
import importlib.util
import io
import os
import re
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_DIR = "/root/reference/dataPrepScripts"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

COUNTER_LITERAL = '{"A":0,"C":0,"G":0,"T":0,"I":0,"D":0,"N":0}'


# ------------------------------------------------------------------------------------------------ CPython 2.7 dict order
def _py27_str_hash(s):
    """Objects/stringobject.c string_hash (64-bit build, no -R randomisation)"""
    m = (1 << 64) - 1
    x = (ord(s[0]) << 7) & m
    for ch in s:
        x = ((1000003 * x) & m) ^ ord(ch)
    x ^= len(s)
    if x >= 1 << 63:
        x -= 1 << 64
    return -2 if x == -1 else x


def py27_dict_order(keys):
    """iteration order of the dict literal {k: ... for k in keys} under CPython 2.7: Objects/dictobject.c -- BUILD_MAP n
    presizes to the smallest power of two > n (for n > 5), every STORE_MAP is a PyDict_SetItem that, when fill*3 >= size*2,
    resizes to the smallest power of two > 4*used and re-inserts the old table in slot order; probing is
    i = (5*i + perturb + 1), perturb >>= 5."""
    def insert(table, key):
        mask = len(table) - 1
        h = _py27_str_hash(key)
        i = h & mask
        perturb = h & ((1 << 64) - 1)
        while table[i & mask] is not None and table[i & mask] != key:
            i = (i << 2) + i + perturb + 1
            perturb >>= 5
        table[i & mask] = key

    size = 8
    if len(keys) > 5:
        while size <= len(keys):
            size <<= 1
    table = [None] * size
    for k in keys:
        insert(table, k)
        fill = sum(1 for t in table if t is not None)
        if fill * 3 >= len(table) * 2:
            new = 8
            while new <= 4 * fill:
                new <<= 1
            old, table = table, [None] * new
            for t in old:
                if t is not None:
                    insert(table, t)
    return [t for t in table if t is not None]


# ------------------------------------------------------------------------------------------------- stand-ins for the tools
class _IntervalTree(object):
    """intervaltree.IntervalTree as the script uses it: addi(begin, end) half-open, search(point) -> overlapping intervals"""

    def __init__(self):
        self.iv = []

    def addi(self, begin, end):
        self.iv.append((begin, end))

    def search(self, begin, end=None):
        if end is None:
            return [iv for iv in self.iv if iv[0] <= begin < iv[1]]
        return [iv for iv in self.iv if iv[0] < end and begin < iv[1]]


def _ref_span(pos1, cigar):
    n = 0
    for m in re.finditer(r"(\d+)([MIDNSHP=X])", cigar):
        if m.group(2) in "MDN=X":
            n += int(m.group(1))
    return pos1, pos1 + max(n, 1) - 1


def _parse_region(reg):
    m = re.match(r"^(.*?)(?::(\d+)-(\d+))?$", reg)
    return m.group(1), (int(m.group(2)) if m.group(2) else None), (int(m.group(3)) if m.group(3) else None)


class _Stdout(io.StringIO):
    def close(self):     # the scripts' *Stdout wrappers close their handle in __del__; keep the text readable
        pass


class _Popen(object):
    def __init__(self, argv, stdout=None, stdin=None, stderr=None, bufsize=0):
        self.returncode = 0
        if argv[1:2] == ["faidx"]:
            ctg, a, b = _parse_region(argv[3])
            name, seq = None, []
            for line in open(argv[2]):
                if line.startswith(">"):
                    name = line[1:].split()[0]
                elif name == ctg:
                    seq.append(line.strip())
            seq = "".join(seq)
            if not seq:
                self.returncode = 1
            if a is not None:
                seq = seq[a - 1:b]       # samtools clips the end of the region to the contig
            text = ">%s\n" % argv[3] + "".join(seq[i:i + 60] + "\n" for i in range(0, len(seq), 60))
        elif argv[1:2] == ["view"]:
            assert argv[2:4] == ["-F", "2308"], argv
            ctg, a, b = _parse_region(argv[5])
            out = []
            for line in open(argv[4]):
                if line.startswith("@") or not line.strip():
                    continue
                f = line.split("\t")
                if f[2] != ctg or (int(f[1]) & 2308):
                    continue
                lo, hi = _ref_span(int(f[3]), f[5])
                if a is not None and (hi < a or lo > b):
                    continue
                out.append(line if line.endswith("\n") else line + "\n")
            text = "".join(out)
        elif argv[:2] == ["gzip", "-fdc"]:
            text = open(argv[2]).read()
        elif argv == ["gzip", "-c"]:         # writer: text written to .stdin lands (uncompressed) in the file given as stdout
            self._sink, self.stdin = stdout, _Stdout()
            return
        else:
            raise AssertionError("command not emulated: %r" % (argv,))
        self.stdout = io.StringIO(text)

    def wait(self):
        if getattr(self, "_sink", None) is not None:
            self._sink.write(self.stdin.getvalue().encode())
            self._sink = None
        return self.returncode


# -------------------------------------------------------------------------------------------- loading the reference source
REWRITES = {
    "ExtractVariantCandidates": [
        (r"print >> sys\.stderr, (.*)$", r"print(\1, file=sys.stderr)", 4),
        (r"pileup\[(\w+)\]\.items\(\)", r"list(pileup[\1].items())", 2),
        (r"remainder = pileup\.keys\(\)", r"remainder = list(pileup.keys())", 1),
    ],
    "CreateTensor": [
        (r"print >> sys\.stderr, (.*)$", r"print(\1, file=sys.stderr)", 6),
        (r"\bxrange\(", r"range(", 2),
        (r"for center in centerToAln\.keys\(\):", r"for center in list(centerToAln.keys()):", 2),
    ],
    "utils_v2": [
        (r"print >> sys\.stderr, (.*)$", r"print(\1, file=sys.stderr)", 4),
    ],
    "train": [
        (r"dtype=np\.int \)", r"dtype=int )", 3),      # (NumPy removed the np.int alias; not a Python 2 matter)
    ],
    "evaluate": [
        (r"dtype=np\.int \)", r"dtype=int )", 3),
    ],
    "calTrainDevDiff": [
        (r"print >> sys\.stderr, (.*)$", r"print(\1, file=sys.stderr)", 2),
    ],
    "GetTruth": [],
    "PairWithNonVariants": [],
    "callVarBamParallel": [
        (r"range\(0,23\)\+", r"list(range(0,23))+", 2),
    ],
    "trainNonstop": [],
    "trainWithoutValidationNonstop": [],
    "callVar": [
        (r"print >> call_fh, (.*)$", r"print(\1, file=call_fh)", 14),
        (r"#print >> sys\.stderr, (.*)$", r"#print(\1, file=sys.stderr)", 1),
    ],
}


def load_reference(name, counter_order=None, ref_dir=None):
    ref_dir = ref_dir or REF_DIR
    src = open(os.path.join(ref_dir, name + ".py")).read()
    for pat, rep, count in REWRITES[name]:
        src, n = re.subn(pat, rep, src, flags=re.M)
        assert n == count, (name, pat, n)
    if counter_order is not None:
        n = src.count(COUNTER_LITERAL)
        assert n == 3, n
        src = src.replace(COUNTER_LITERAL, "{" + ",".join('"%s":0' % k for k in counter_order) + "}")
    spec = importlib.util.spec_from_file_location("param", os.path.join(ref_dir, "param.py"))
    param = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(param)
    it = types.ModuleType("intervaltree")
    it.IntervalTree = _IntervalTree
    bl = types.ModuleType("blosc")          # python-blosc (absent): an identity container, the arrays themselves are compared
    bl.set_nthreads = lambda n: None
    bl.pack_array = lambda a, cname=None: ("blosc-stand-in", np.array(a, copy=True))
    bl.unpack_array = lambda b: b[1]
    saved = {k: sys.modules.get(k) for k in ("param", "intervaltree", "blosc")}
    sys.modules["param"], sys.modules["intervaltree"], sys.modules["blosc"] = param, it, bl
    try:
        mod = types.ModuleType("reference_" + name)
        exec(compile(src, os.path.join(ref_dir, name + ".py"), "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    fake = types.SimpleNamespace(Popen=_Popen, PIPE=-1)
    mod.subprocess = fake
    return mod


def run_main(mod, argv, stdin_text=""):
    out = _Stdout()
    old_argv, old_stdin = sys.argv, sys.stdin
    sys.argv, sys.stdin = ["reference"] + argv, io.StringIO(stdin_text)
    try:
        with contextlib.redirect_stdout(out):
            try:
                mod.main()
            except SystemExit as e:
                if e.code not in (0, None):
                    raise
    finally:
        sys.argv, sys.stdin = old_argv, old_stdin
    return out.getvalue()


# ----------------------------------------------------------------------------------------------------------- the scenarios
def scenarios():
    from test_pileup_cpu import synth_alignments
    sc = []
    rng = np.random.default_rng(77)
    ref, sam, _ = synth_alignments(rng, ref_len=1600, n_reads=260)
    sc.append(dict(name="defaults", ref=ref, sam=sam, bed="", can_args=[], ten_args=[]))
    rng = np.random.default_rng(78)
    ref, sam, _ = synth_alignments(rng, ref_len=2200, n_reads=260, dup_pos=0.4, lower=0.0)
    bed = "ctg\t100\t700\nctg\t900\t901\nother\t0\t50\nctg\t1200\t2100\n"
    sc.append(dict(name="region_bed_filters", ref=ref, sam=sam, bed=bed,
                   can_args=["--ctgStart", "150", "--ctgEnd", "1900", "--minMQ", "15", "--threshold", "0.1", "--minCoverage", "6"],
                   ten_args=["--ctgStart", "150", "--ctgEnd", "1900", "--minMQ", "15", "--dcov", "3", "--minCoverage", "3"]))
    rng = np.random.default_rng(79)
    ref, sam, _ = synth_alignments(rng, ref_len=1400, n_reads=300, ops="MIDS")
    sc.append(dict(name="no_left_edge_low_threshold", ref=ref, sam=sam, bed="",
                   can_args=["--threshold", "0.05", "--minCoverage", "2"],
                   ten_args=["--considerleftedge", "False", "--minCoverage", "2"]))
    for s in sc:   # the reference's counter dict has no key for read bases other than A C G T N (KeyError): keep to those
        rows = []
        for line in s["sam"].split("\n"):
            f = line.split("\t")
            if len(f) > 9 and not line.startswith("@"):
                f[9] = re.sub(r"[^ACGTN]", "N", f[9].upper())
                line = "\t".join(f)
            rows.append(line)
        s["sam"] = "\n".join(rows)
    return sc


# ------------------------------------------------------------------------ feed (utils_v2.py) and VCF writer (callVar.py)
CV_DIR = "/root/reference/clairvoyante"


class StubModel(object):
    """stands where the TensorFlow model would: hands out rows of a fixed probability table in call order (both the reference's
    Test loop and this repository's call predictNoRT once per batch, in stream order)"""

    def __init__(self, table):
        self.table, self.at = table, 0

    def predictNoRT(self, X):
        p = self.table[self.at:self.at + len(X)]
        assert len(p) == len(X)
        self.at += len(X)
        self.predictBaseRTVal, self.predictZygosityRTVal = p[:, 0:4], p[:, 4:6]
        self.predictVarTypeRTVal, self.predictIndelLengthRTVal = p[:, 6:10], p[:, 10:16]


def probability_table(n, seed):
    """head outputs as the network would give them (sigmoid base head, three softmax heads), spread over every class"""
    rng = np.random.RandomState(seed)
    lg = rng.randn(n, 16) * 3.0
    out = np.empty((n, 16), np.float32)
    out[:, 0:4] = 1.0 / (1.0 + np.exp(-lg[:, 0:4]))
    for a, b in ((4, 6), (6, 10), (10, 16)):
        e = np.exp(lg[:, a:b] - lg[:, a:b].max(1, keepdims=True))
        out[:, a:b] = e / e.sum(1, keepdims=True)
    return out


def feed_and_callvar(tensor_text, tmp):
    """runs the reference's GetTensor / GetTrainingArray / DecompressArray and its callVar Test + Output on tensor_text"""
    fx = {}
    rng = np.random.RandomState(5)
    lines = [l for l in tensor_text.split("\n") if l]
    extra = []
    for i, l in enumerate(lines[:60]):          # rows the feed must drop or tolerate: centre base not ACGT, lower-case context
        f = l.split(" ")
        seq = f[2]
        if i % 3 == 0:
            seq = seq[:16] + "N" + seq[17:]
        elif i % 3 == 1:
            seq = seq.lower()
        extra.append(" ".join([f[0], str(500000 + i), seq] + f[3:]))
    zero = lines[0].split(" ")
    extra.append(" ".join(zero[:1] + ["600000", zero[2]] + ["0.0"] * 528))   # no coverage at all: Output prints nothing
    lines = lines + extra
    order = rng.permutation(len(lines))
    text = "".join(lines[i] + "\n" for i in order)
    tfn = os.path.join(tmp, "tensors.txt")
    open(tfn, "w").write(text)
    fx["feed/tensor_text"] = np.frombuffer(text.encode(), np.uint8)

    utils = load_reference("utils_v2", ref_dir=CV_DIR)
    cv_param = sys.modules.get("param")
    # GetTensor, 100-row batches
    ends, counts, xs, poss = [], [], [], []
    for end, c, x, pos in utils.GetTensor(tfn, 100):
        ends.append(end); counts.append(c); xs.append(np.array(x[:c], copy=True)); poss += list(pos)
    fx["feed/gettensor_end"] = np.array(ends)
    fx["feed/gettensor_count"] = np.array(counts)
    fx["feed/gettensor_x"] = np.concatenate(xs).astype(np.int16)
    assert np.array_equal(fx["feed/gettensor_x"], np.concatenate(xs))
    fx["feed/gettensor_pos"] = np.array(poss)

    # truth variants for GetTrainingArray: every label shape of utils_v2.py:90-117
    keep = [l.split(" ") for l in lines if l.split(" ")[2][16] in "ACGT"]
    var_rows = []
    for i, f in enumerate(keep[::3]):
        ref = f[2][16].upper()
        alt = "ACGT"[(("ACGT".index(ref)) + 1 + i % 3) % 4]
        kind = i % 7
        if kind in (0, 1):
            r, a = ref, alt
        elif kind == 2:
            r, a = ref, ref + "ACGTAC"[:1 + i % 6]
        elif kind == 3:
            r, a = ref + "TTGCAA"[:1 + i % 6], ref
        elif kind == 4:
            r, a = ref + "T", ref + "GG"
        else:
            r, a = ref, alt
        gt = ("0", "1") if i % 2 == 0 else ("1", "1")
        if kind == 6:
            gt = ("1", "2")
        var_rows.append("%s %s %s %s %s %s" % (f[0], f[1], r, a, gt[0], gt[1]))
    vfn, bfn = os.path.join(tmp, "vars.txt"), os.path.join(tmp, "conf.bed")
    open(vfn, "w").write("".join(r + "\n" for r in var_rows))
    bed = "ctg\t0\t900\nctg\t1000\t1001\nctg\t1100\t700000\n"
    open(bfn, "w").write(bed)
    fx["feed/var_text"] = np.array("".join(r + "\n" for r in var_rows))
    fx["feed/bed_text"] = np.array(bed)
    for tag, b in (("nobed", None), ("bed", bfn)):
        total, XC, YC, PC = utils.GetTrainingArray(tfn, vfn, b, shuffle=False)
        fx["feed/train_%s_total" % tag] = np.array(total)
        fx["feed/train_%s_blocks" % tag] = np.array([len(x[1]) for x in XC])
        X = np.concatenate([x[1].reshape(-1, 33, 4, 4) for x in XC])
        fx["feed/train_%s_x" % tag] = X.astype(np.int16)
        assert np.array_equal(X, fx["feed/train_%s_x" % tag])
        fx["feed/train_%s_y" % tag] = np.concatenate([y[1].reshape(-1, 16) for y in YC]).astype(np.float64)
        fx["feed/train_%s_pos" % tag] = np.concatenate([q[1].reshape(-1) for q in PC]).astype(str)
        if tag == "nobed":
            cases = [(0, 500), (0, total), (123, 1000), (499, 2), (500, 500), (total - 7, 1000), (250, 250)]
            fx["feed/decompress_cases"] = np.array(cases)
            for k, (st, num) in enumerate(cases):
                a, n, e = utils.DecompressArray(YC, st, num, total)
                fx["feed/decompress_%d" % k] = np.array(a)
                fx["feed/decompress_%d_ne" % k] = np.array([n, e])

    # callVar: Test + Output with the stub model, 100-row batches, for three option sets
    cv = load_reference("callVar", ref_dir=CV_DIR)
    cv.param.predictBatchSize = 100
    utils_param = [m for m in (utils.param,)][0]
    assert utils_param.flankingBaseNum == 16
    table = probability_table(len(poss), 11)
    fx["callvar/probabilities"] = table
    fai = os.path.join(tmp, "ref.fa.fai")
    open(fai, "w").write("ctg\t2200\t5\t70\t71\nother\t100\t2300\t70\t71\n")
    for tag, kw in (("default", dict(qual=None, showRef=False, ref_fn=None)), ("showref_qual", dict(qual=20, showRef=True, ref_fn=None)),
                    ("contigs", dict(qual=150, showRef=False, ref_fn=os.path.join(tmp, "ref.fa")))):
        out = os.path.join(tmp, "calls_%s.vcf" % tag)
        args = types.SimpleNamespace(v2=False, v3=True, slim=False, tensor_fn=tfn, call_fn=out, sampleName="SAMPLE", threads=None, **kw)
        cv.Test(args, StubModel(table), utils)
        import gc
        gc.collect()
        fx["callvar/vcf_" + tag] = np.array(open(out).read())
        print("callVar %-13s %5d VCF records" % (tag, sum(1 for l in open(out) if not l.startswith("#"))))
    return fx


# --------------------------------------------------------------------------------------- training driver (train.py)
class StubTrainer(object):
    """stands where the TensorFlow model would in train.py's TrainAll: deterministic losses (the per-row validation loss
    alternates from epoch to epoch, which is what drives the reference's learning-rate switches), every call logged"""

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


DRIVERS = [   # (module, entry point, param overrides, extra args)
    ("train", "TrainAll", dict(trainBatchSize=200, predictBatchSize=50), {}),
    ("trainNonstop", "TrainAll", dict(trainBatchSize=200, predictBatchSize=50, maxEpoch=5), {}),
    ("trainWithoutValidationNonstop", "TrainAll", dict(trainBatchSize=150, predictBatchSize=50, maxEpoch=4), {}),
    ("evaluate", "Test", dict(predictBatchSize=70), {}),
    ("calTrainDevDiff", "CalcAll", dict(predictBatchSize=60), dict(chkpnt_fn=["model-000003", "model-000008"])),
]


def run_drivers(fx, tmp):
    """train.py / trainNonstop.py / trainWithoutValidationNonstop.py / evaluate.py / calTrainDevDiff.py with StubTrainer"""
    import logging
    import pickle
    X = fx["feed/train_nobed_x"].astype(np.float32)
    Y = fx["feed/train_nobed_y"]
    pos = fx["feed/train_nobed_pos"]
    total = int(fx["feed/train_nobed_total"])
    utils = load_reference("utils_v2", ref_dir=CV_DIR)
    bs = utils.param.bloscBlockSize
    blocks = lambda a: [("blosc-stand-in", a[i:i + bs]) for i in range(0, total, bs)]
    bin_fn = os.path.join(tmp, "train.bin")
    with open(bin_fn, "wb") as fh:
        for obj in (total, blocks(X), blocks(Y), blocks(pos)):
            pickle.dump(obj, fh)
    out = {}
    for name, entry, overrides, extra in DRIVERS:
        mod = load_reference(name, ref_dir=CV_DIR)
        for k, v in overrides.items():
            setattr(mod.param, k, v)
        msgs = []

        class Collect(logging.Handler):
            def emit(self, record):
                msgs.append(record.getMessage())

        h = Collect()
        logging.getLogger().addHandler(h)
        logging.getLogger().setLevel(logging.INFO)
        err = io.StringIO()
        try:
            m = StubTrainer()
            kw = dict(bin_fn=bin_fn, tensor_fn=None, var_fn=None, bed_fn=None, chkpnt_fn=None, learning_rate=1e-3, lambd=1e-3,
                      ochk_prefix=os.path.join(tmp, "model"), olog_dir=None, v2=False, v3=True, slim=False)
            kw.update(extra)
            with contextlib.redirect_stderr(err):
                getattr(mod, entry)(types.SimpleNamespace(**kw), m, utils)
        finally:
            logging.getLogger().removeHandler(h)
        msgs = [x for x in msgs if "time elapsed" not in x] + [l for l in err.getvalue().split("\n") if l]
        print("%-30s %4d model calls, %3d log lines" % (name + ".py:", len(m.calls), len(msgs)))
        out[name + "/calls"] = np.array([repr(c) for c in m.calls])
        out[name + "/log"] = np.array(msgs)
    return out


# ------------------------------------------------------------------- GetTruth, PairWithNonVariants, callVarBamParallel
def run_small_scripts(fx, tmp):
    import random
    out = {}
    # GetTruth: VCF -> "ctg pos ref alt gt1 gt2" rows
    vcf = ["##fileformat=VCFv4.1", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS"]
    gts = ["0/1", "1/1", "0|1", "1|0", "1|2", "./1", "1/2", "2|1"]
    rng = np.random.RandomState(3)
    for i in range(120):
        ref = "ACGT"[i % 4] + ("TG" * (i % 3 == 2))
        alts = ["ACGT"[(i + 1) % 4] + "A" * (i % 5 == 1), "ACGT"[(i + 2) % 4] + "CC" * (i % 7 == 3)]
        gt = gts[i % len(gts)]
        alt = ",".join(alts) if "2" in gt or i % 11 == 0 else alts[0]
        vcf.append("\t".join(["ctg" if i % 13 else "other", str(100 + 17 * i), ".", ref, alt, "50", "PASS", ".", "GT:GQ", gt + ":%d" % (i % 90)]))
    vfn = os.path.join(tmp, "truth.vcf")
    open(vfn, "w").write("\n".join(vcf) + "\n")
    out["gettruth/vcf"] = np.array("\n".join(vcf) + "\n")
    gt_mod = load_reference("GetTruth")
    out["gettruth/all"] = np.array(run_main(gt_mod, ["--vcf_fn", vfn, "--ctgName", "ctg"]))
    out["gettruth/region"] = np.array(run_main(gt_mod, ["--vcf_fn", vfn, "--ctgName", "ctg", "--ctgStart", "400", "--ctgEnd", "1500"]))
    print("GetTruth: %d / %d rows" % (str(out["gettruth/all"]).count("\n"), str(out["gettruth/region"]).count("\n")))

    # PairWithNonVariants: tensors at truth variants + tensors at candidates -> training tensor file (seeded sample)
    lines = [l for l in fx["feed/tensor_text"].tobytes().decode().split("\n") if l]
    var_lines, can_lines = lines[:90], lines[60:]
    tv, tc, bed, outfn = (os.path.join(tmp, n) for n in ("tensor_var", "tensor_can", "pair.bed", "tensor_pair"))
    open(tv, "w").write("".join(l + "\n" for l in var_lines))
    open(tc, "w").write("".join(l + "\n" for l in can_lines))
    open(bed, "w").write("ctg\t0\t1200\nctg\t1500\t700000\n")
    pw = load_reference("PairWithNonVariants")
    for tag, b, amp in (("nobed", None, 2), ("bed", bed, 1)):
        random.seed(12345)
        argv = ["--tensor_can_fn", tc, "--tensor_var_fn", tv, "--output_fn", outfn, "--amp", str(amp)] + (["--bed_fn", b] if b else [])
        run_main(pw, argv)
        got = open(outfn).read()
        out["pair/%s_positions" % tag] = np.array([" ".join(l.split()[:2]) for l in got.split("\n") if l])
        out["pair/%s_sha256" % tag] = np.array(hashlib.sha256(got.encode()).hexdigest())
        print("PairWithNonVariants %-5s: %d rows" % (tag, got.count("\n")))
    out["pair/n_var"], out["pair/n_can_from"] = np.array(90), np.array(60)

    # callVarBamParallel: the per-chunk commands
    fa = os.path.join(tmp, "genome.fa")
    for fn in (fa, os.path.join(tmp, "aln.bam"), os.path.join(tmp, "model.meta")):
        open(fn, "w").write("x\n")
    open(fa + ".fai", "w").write("chr1\t25000000\t6\t60\t61\nchr2\t10000001\t7\t60\t61\nchrUn_x\t5000\t8\t60\t61\n21\t9999999\t9\t60\t61\n")
    cbed = os.path.join(tmp, "chunks.bed")
    open(cbed, "w").write("chr1\t10000000\t10000001\nchr1\t19999999\t20000500\n21\t5\t500\n")
    cp = load_reference("callVarBamParallel", ref_dir=CV_DIR)
    cp.__file__ = os.path.join(CV_DIR, "callVarBamParallel.py")
    real_subprocess = __import__("subprocess")
    cp.subprocess = types.SimpleNamespace(Popen=_Popen, PIPE=-1, check_output=real_subprocess.check_output)
    base = ["--chkpnt_fn", os.path.join(tmp, "model"), "--ref_fn", fa, "--bam_fn", os.path.join(tmp, "aln.bam"), "--pypy", "python",
            "--samtools", "python", "--output_prefix", "out/calls"]
    for tag, extra in (("default", []), ("bed_qual", ["--bed_fn", cbed, "--qual", "30", "--refChunkSize", "5000000", "--sampleName", "HG"]),
                       ("allcontigs", ["--includingAllContigs", "--threshold", "0.2", "--minCoverage", "6", "--tensorflowThreads", "8"])):
        text = run_main(cp, base + extra)
        text = text.replace(tmp, "TMP").replace(CV_DIR, "REFDIR")
        out["parallel/" + tag] = np.array(text)
        out["parallel/%s_args" % tag] = np.array(extra, dtype=str)
        print("callVarBamParallel %-10s: %d commands" % (tag, text.count("\n")))
    return out


# ------------------------------------------------------------------------------- the whole calling pipeline (callVarBam.py)
PIPELINES = [   # (tag, scenario, callVarBam options)
    ("plain", "defaults", {}),
    ("region_bed_qual", "region_bed_filters", dict(ctgStart=150, ctgEnd=1900, bed=True, threshold=0.1, minCoverage=6, qual=30, dcov=5)),
    ("vcf_sites", "defaults", dict(vcf=True, ctgStart=300, ctgEnd=1400)),
]


def run_pipelines(fx, tmp, evc, ct):
    """ExtractVariantCandidates | CreateTensor | callVar exactly as callVarBam.py:113-131 chains them (same per-stage options),
    each stage the reference's own code, the model a probability table"""
    out = {}
    utils = load_reference("utils_v2", ref_dir=CV_DIR)
    cv = load_reference("callVar", ref_dir=CV_DIR)
    cv.param.predictBatchSize = 100
    gt = load_reference("GetTruth")
    table = probability_table(4000, 17)
    out["pipeline/probabilities"] = table
    vfn = os.path.join(tmp, "sites.vcf")
    open(vfn, "w").write(str(fx["gettruth/vcf"]))
    for tag, sc, o in PIPELINES:
        fa, samfn, bedfn = (os.path.join(tmp, "pl_" + tag + e) for e in (".fa", ".sam", ".bed"))
        ref = str(fx[sc + "/ref"])
        open(fa, "w").write(">ctg\n" + "".join(ref[i:i + 70] + "\n" for i in range(0, len(ref), 70)))
        open(fa + ".fai", "w").write("ctg\t%d\t5\t70\t71\n" % len(ref))
        open(samfn, "w").write(str(fx[sc + "/sam"]))
        rng_args = ["--ctgStart", str(o["ctgStart"]), "--ctgEnd", str(o["ctgEnd"])] if "ctgStart" in o else []
        if o.get("vcf"):
            cand = run_main(gt, ["--vcf_fn", vfn, "--ctgName", "ctg"] + rng_args)
        else:
            argv = ["--bam_fn", samfn, "--ref_fn", fa]
            if o.get("bed"):
                open(bedfn, "w").write(str(fx[sc + "/bed"]))
                argv += ["--bed_fn", bedfn]
            argv += ["--ctgName", "ctg"] + rng_args + ["--threshold", str(o.get("threshold", 0.125)), "--minCoverage", str(o.get("minCoverage", 4)),
                                                         "--samtools", "samtools"]
            cand = run_main(evc, argv)
        tens = run_main(ct, ["--bam_fn", samfn, "--ref_fn", fa, "--ctgName", "ctg"] + rng_args + ["--considerleftedge", "--samtools", "samtools",
                             "--dcov", str(o.get("dcov", 250))], stdin_text=cand)
        tfn, vcf_out = os.path.join(tmp, "pl_" + tag + ".tensors"), os.path.join(tmp, "pl_" + tag + ".vcf")
        open(tfn, "w").write(tens)
        args = types.SimpleNamespace(v2=False, v3=True, slim=False, tensor_fn=tfn, call_fn=vcf_out, sampleName="SAMPLE", threads=None,
                                     qual=o.get("qual"), showRef=False, ref_fn=fa)
        cv.Test(args, StubModel(table[:sum(1 for l in tens.split("\n") if l and l.split()[2][16].upper() in "ACGT")]), utils)
        import gc
        gc.collect()
        out["pipeline/%s_vcf" % tag] = np.array(open(vcf_out).read())
        print("pipeline %-16s %4d candidates, %4d tensors, %4d VCF records" % (
            tag, cand.count("\n"), tens.count("\n"), sum(1 for l in open(vcf_out) if not l.startswith("#"))))
    return out


def main():
    order27 = py27_dict_order(["A", "C", "G", "T", "I", "D", "N"])
    print("CPython 2.7 iteration order of the counter literal:", " ".join(order27))
    evc27 = load_reference("ExtractVariantCandidates", order27)
    evc3 = load_reference("ExtractVariantCandidates", None)
    ct = load_reference("CreateTensor")
    fixture = {"py27_counter_order": np.array(order27)}
    names, all_tensor_text = [], []
    with tempfile.TemporaryDirectory() as tmp:
        for s in scenarios():
            fa, samfn, bedfn, canfn = (os.path.join(tmp, s["name"] + e) for e in (".fa", ".sam", ".bed", ".can"))
            ref = s["ref"]
            open(fa, "w").write(">ctg\n" + "".join(ref[i:i + 70] + "\n" for i in range(0, len(ref), 70)))
            open(fa + ".fai", "w").write("ctg\t%d\t5\t70\t71\n" % len(ref))
            open(samfn, "w").write(s["sam"])
            common = ["--bam_fn", samfn, "--ref_fn", fa, "--ctgName", "ctg"]
            can_argv = common + s["can_args"]
            if s["bed"]:
                open(bedfn, "w").write(s["bed"])
                can_argv += ["--bed_fn", bedfn]
            rows27 = run_main(evc27, list(can_argv))
            rows3 = run_main(evc3, list(can_argv))
            open(canfn, "w").write(rows27)
            tens = run_main(ct, common + ["--can_fn", canfn] + s["ten_args"])
            lines = [l for l in tens.split("\n") if l]
            x = np.array([[float(v) for v in l.split()[3:]] for l in lines], np.float32).reshape(-1, 33, 4, 4)
            assert np.array_equal(x, np.round(x)) and x.max(initial=0) < 32000
            n = s["name"]
            names.append(n)
            all_tensor_text.append(tens)
            fixture[n + "/sam"] = np.array(s["sam"])
            fixture[n + "/ref"] = np.array(ref)
            fixture[n + "/bed"] = np.array(s["bed"])
            fixture[n + "/can_args"] = np.array(s["can_args"], dtype=str)
            fixture[n + "/ten_args"] = np.array(s["ten_args"], dtype=str)
            fixture[n + "/candidate_rows"] = np.array(rows27)
            fixture[n + "/candidate_rows_py3_dict_order"] = np.array(rows3)
            fixture[n + "/tensor_head"] = np.array([" ".join(l.split()[:3]) for l in lines])
            fixture[n + "/tensors"] = x.astype(np.int16)
            fixture[n + "/tensor_text_sha256"] = np.array(hashlib.sha256(tens.encode()).hexdigest())
            print("%-28s %4d candidate rows (%d differ under the py3 dict order), %4d tensors"
                  % (n, rows27.count("\n"), sum(a != b for a, b in zip(rows27.split("\n"), rows3.split("\n"))), len(lines)))
        fixture.update(feed_and_callvar("".join(all_tensor_text), tmp))
        fixture.update(run_drivers(fixture, tmp))
        fixture.update(run_small_scripts(fixture, tmp))
        fixture.update(run_pipelines(fixture, tmp, evc27, ct))
    fixture["scenarios"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "reference_run.npz"), **fixture)
    print("written", os.path.join(HERE, "reference_run.npz"))


if __name__ == "__main__":
    main()

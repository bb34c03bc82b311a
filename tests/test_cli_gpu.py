"""GPU: the command lines end to end -- callVar.py on a text tensor file with a saved checkpoint (VCF text compared
with oracle predictions pushed through the per-site VCF restatement), train.py on a small .bin, DP trainer on 1 rank."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from clairvoyante_b200 import initializers as I, param, synth, utils_v2 as U
from oracle import callvar_output as CO, cv_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_tensors(fn, x, start=5000):
    raw = x.copy(); raw[..., 1:] += raw[..., 0:1]
    rng = np.random.default_rng(1)
    seqs = ["".join(rng.choice(list("ACGT"), 33)) for _ in range(len(x))]
    with open(fn, "w") as fh:
        for i, t in enumerate(raw):
            fh.write("chr20 %d %s %s\n" % (start + i, seqs[i], " ".join("%0.1f" % v for v in t.reshape(-1))))
    return ["chr20:%d:%s" % (start + i, s) for i, s in enumerate(seqs)]


@pytest.mark.parametrize("slim", [False, True])
def test_callvar_cli(tmp_path, slim):
    """callVar.py end to end on TRAINED-LIKE weights (tests/test_trained_parity.py): every VCF record must equal the one the
    fp64 oracle's probabilities give through the per-site restatement of Output (callVar.py:58-153), except at sites that
    are ENUMERATED beforehand from the oracle alone: a top-2 logit margin <= 2e-3 on a head the record reads, or a
    real-valued QUAL within 0.02 of an integer (int() truncation, :72) or a runner-up probability product under float32's
    normal range -- and at those only QUAL / FILTER / GQ may differ"""
    from math import log
    from test_trained_parity import load_trained
    variant = "v3_slim" if slim else "v3"
    W, fx, x, _ = load_trained(variant)
    n = 2300                                           # 2 full batches of predictBatchSize + a tail
    x = x[:n]
    tfn, ck, out = str(tmp_path / "t.txt"), str(tmp_path / "model"), str(tmp_path / "o.vcf")
    pos = _write_tensors(tfn, x)
    np.savez(ck + ".cvb.npz", **W)
    cmd = [sys.executable, "-m", "clairvoyante_b200.callVar", "--tensor_fn", tfn, "--chkpnt_fn", ck, "--call_fn", out,
           "--showRef", "--qual", "10", "--sampleName", "HG001"] + (["--slim"] if slim else [])
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    body = [ln for ln in open(out).read().splitlines() if not ln.startswith("#")]
    ref = O.forward(W, x, variant)                     # fp64: margins and the real-valued QUAL
    r32 = O.forward(W, x, variant, dtype=np.float32)   # the reference computes (and callVar.py:72 multiplies) in float32
    lg = ref["logits"]
    exp, tie, qedge = [], [], []
    for j in range(n):
        e = CO.vcf_line(x[j], pos[j], r32["base"][j], r32["zygosity"][j], r32["varType"][j], r32["indelLength"][j], True, 10)
        if e is None:
            continue
        exp.append(e)
        gaps = [np.diff(np.sort(lg[j, a:b]))[-1] for a, b in ((4, 6), (6, 10), (10, 16))]
        bs = np.sort(lg[j, 0:4])
        bp = np.sort(ref["base"][j])                    # Output orders the base head by its sigmoid PROBABILITIES (:81), which
        sat = min((bp[3] - bp[2]) / bp[3], (bp[2] - bp[1]) / max(bp[2], 1e-300)) < 2e-6 or bp[2] < 1e-37   # float32 cannot tell apart once two saturate
        tie.append(min(gaps + [bs[3] - bs[2], bs[2] - bs[1]]) <= 2e-3 or sat)
        st, sz, sl = (np.sort(ref[k][j])[::-1] for k in ("varType", "zygosity", "indelLength"))
        q = -4.343 * log((st[1] * sz[1] * sl[1] + 1e-300) / (st[0] * sz[0] * sl[0] + 1e-300))
        # ... or the runner-up product st[1]*sz[1]*sl[1] is below float32's normal range: it underflows (to a denormal in
        # NumPy, to zero under the GPU's flush-to-zero exp), QUAL = int(-4.343 * log(1e-300 / ...)) jumps to ~3000
        qedge.append(abs(q - round(q)) <= 0.02 or st[1] * sz[1] * sl[1] < 1e-36)
    assert len(body) == len(exp) and [b.split("\t")[1] for b in body] == [e.split("\t")[1] for e in exp]
    n_tie = n_edge = 0
    for b, e, is_tie, is_edge in zip(body, exp, tie, qedge):
        if b == e:
            continue
        if is_tie:
            n_tie += 1
            continue
        fb, fe = b.split("\t"), e.split("\t")
        sb, se = fb[9].split(":"), fe[9].split(":")
        only_qual = fb[:5] == fe[:5] and fb[7:9] == fe[7:9] and sb[0] == se[0] and sb[2:] == se[2:]
        assert is_edge and only_qual, "record differs away from every enumerated near-tie / QUAL boundary:\n%s\n%s" % (b, e)
        n_edge += 1
    print("%s: %d records, %d differ at near-tie sites (%d enumerated), %d at QUAL boundaries (%d enumerated)"
          % (variant, len(exp), n_tie, sum(tie), n_edge, sum(qedge)))
    assert sum(tie) < 0.02 * len(exp)


def test_train_cli_and_resume(tmp_path):
    total = 2600
    X, Y = synth.make_sites(total, 3), synth.make_labels(total, 3).astype(np.float64)
    bs = param.bloscBlockSize
    fn = str(tmp_path / "d.bin")
    with open(fn, "wb") as fh:
        for p in (total, [U.pack_array(X[i:i + bs]) for i in range(0, total, bs)],
                  [U.pack_array(Y[i:i + bs]) for i in range(0, total, bs)], []):
            pickle.dump(p, fh)
    env = dict(os.environ, CVB_MAX_EPOCH="3")
    cmd = [sys.executable, "-c",
           "import sys; from clairvoyante_b200 import param, train; param.maxEpoch=3; param.trainBatchSize=1000; "
           "sys.argv=['train','--bin_fn',%r,'--ochk_prefix',%r,'--learning_rate','1e-4','--olog_dir',%r]; train.main()"
           % (fn, str(tmp_path / "ck" / "m"), str(tmp_path / "log"))]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    assert os.path.exists(str(tmp_path / "ck" / "m-000001.cvb.npz")) and os.path.exists(str(tmp_path / "ck" / "m-000002.cvb.npz"))
    assert "Training loss:" in r.stderr and "all/top1/top2" in r.stderr
    assert os.path.getsize(str(tmp_path / "log" / "summaries.jsonl")) > 0


def test_dp_trainer_single_rank_equals_plain_step():
    from clairvoyante_b200 import clairvoyante_v3 as cv, parallel
    W = I.init_weights("v3", 2)
    x, y = synth.make_sites(500, 1), synth.make_labels(500, 1)
    a = cv.Clairvoyante(dropoutRateFC4=0.0); a.init(seed=1); a.setWeights(W)
    b = cv.Clairvoyante(dropoutRateFC4=0.0); b.init(seed=1); b.setWeights(W)
    la, _ = a.train(x, y)
    lb, _ = parallel.DataParallelTrainer(b).train(x, y, seed=5)
    assert abs(float(la) - float(lb)) <= 1e-6 * abs(float(la))
    wa, wb = a.getWeights(), b.getWeights()
    for k in wa:       # weight-gradient kernels accumulate with atomics: summation order differs run to run
        assert np.allclose(wa[k], wb[k], rtol=1e-5, atol=1e-7), k
    a.close(); b.close()

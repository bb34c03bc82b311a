"""CPU: callVar's VCF text (batched NumPy Output) against the per-site restatement of callVar.py:50-153, and the
three-way pipeline of callVar.py:180-216 with a stub model."""
import io
import types

import numpy as np

from clairvoyante_b200 import callVar, param, synth
from oracle import callvar_output as CO


def _softmaxish(rng, n, k):
    a = rng.random((n, k)).astype(np.float32) ** 3
    return a / a.sum(1, keepdims=True)


def _case(n, seed):
    rng = np.random.default_rng(seed)
    x = synth.make_sites(n, seed)
    x[rng.random(n) < 0.1] = 0                               # dp == 0 sites print nothing
    big = rng.random(n) < 0.2                                 # long indel evidence across the right flank
    x[big, 17:, :, 1] += 40; x[big, 17:, :, 2] += 40
    base = rng.random((n, 4)).astype(np.float32)
    base[rng.random(n) < 0.3] = 1.0                           # saturated sigmoid ties
    z, t, l = _softmaxish(rng, n, 2), _softmaxish(rng, n, 4), _softmaxish(rng, n, 6)
    l[rng.random(n) < 0.3, 5] = 1.0                           # force the >4 length class often
    seqs = ["".join(rng.choice(list("ACGT"), 33)) for _ in range(n)]
    pos = ["chr%d:%d:%s" % (1 + i % 3, 100 + i, s) for i, s in enumerate(seqs)]
    return x, pos, base, z, t, l


def test_output_matches_per_site_restatement():
    for show_ref, qual in ((False, None), (True, None), (False, 5), (True, 30)):
        args = types.SimpleNamespace(showRef=show_ref, qual=qual)
        x, pos, base, z, t, l = _case(600, 3)
        fh = io.StringIO()
        callVar.Output(args, fh, len(x), x, pos, base, z, t, l)
        got = fh.getvalue().splitlines()
        exp = [CO.vcf_line(x[j], pos[j], base[j], z[j], t[j], l[j], show_ref, qual) for j in range(len(x))]
        exp = [e for e in exp if e is not None]
        assert got == exp
        assert len(got) > 50
        if show_ref:
            assert any("\t0/0:" in g for g in got)
    kinds = "".join(got)
    assert "<INS>" in kinds and "<DEL>" in kinds and "LENGUESS=" in kinds


def test_output_inconsistent_shapes_exits():
    import pytest
    args = types.SimpleNamespace(showRef=True, qual=None)
    x, pos, base, z, t, l = _case(4, 1)
    with pytest.raises(SystemExit):
        callVar.Output(args, io.StringIO(), 5, x, pos, base, z, t, l)


def test_header(tmp_path):
    fa = tmp_path / "ref.fa"
    (tmp_path / "ref.fa.fai").write_text("chr1\t1000\t6\t60\t61\nchr2\t500\t1100\t60\t61\n")
    args = types.SimpleNamespace(ref_fn=str(fa), sampleName="NA12878")
    fh = io.StringIO()
    callVar.PrintVCFHeader(args, fh)
    lines = fh.getvalue().splitlines()
    assert lines[0] == "##fileformat=VCFv4.1" and lines[-1].endswith("FORMAT\tNA12878")
    assert "##contig=<ID=chr2,length=500>" in lines and len(lines) == 14


class _StubModel(object):
    """deterministic stand-in with the predictNoRT contract (clairvoyante_v3.py:269-280)"""

    def __init__(self):
        self.calls = []

    def predictNoRT(self, X):
        n = len(X)
        self.calls.append(n)
        s = X.reshape(n, 528).sum(1)
        t = np.zeros((n, 4), np.float32); t[np.arange(n), (np.abs(s).astype(int)) % 4] = 0.9; t += 0.025
        self.predictBaseRTVal = np.tile(np.array([[0.1, 0.7, 0.2, 0.9]], np.float32), (n, 1))
        self.predictZygosityRTVal = np.tile(np.array([[0.3, 0.7]], np.float32), (n, 1))
        self.predictVarTypeRTVal = t
        self.predictIndelLengthRTVal = np.tile(np.array([[0.5, 0.2, 0.1, 0.1, 0.05, 0.05]], np.float32), (n, 1))


def test_pipeline_orders_batches(tmp_path, monkeypatch):
    from test_feed_cpu import _rows
    from clairvoyante_b200 import utils_v2
    monkeypatch.setattr(param, "predictBatchSize", 8)
    for n in (0, 5, 8, 16, 21):
        x = synth.make_sites(n, 2)
        tfn = str(tmp_path / ("t%d.txt" % n))
        open(tfn, "w").write("".join(r + "\n" for r in _rows(x)))
        args = types.SimpleNamespace(tensor_fn=tfn, call_fn=str(tmp_path / ("o%d.vcf" % n)), showRef=True, qual=None, ref_fn=None,
                                     sampleName="S")
        m = _StubModel()
        callVar.Test(args, m, utils_v2)
        body = [ln for ln in open(args.call_fn).read().splitlines() if not ln.startswith("#")]
        coords = [int(ln.split("\t")[1]) for ln in body]
        assert coords == sorted(coords) and len(set(coords)) == len(coords)
        assert sum(m.calls) == n and m.calls == [8] * (n // 8) + [n % 8]      # incl. the legal empty tail batch
        # every site with depth is reported exactly once with --showRef
        ref = _StubModel(); ref.predictNoRT(x)
        exp = [CO.vcf_line(x[j], "chr1:%d:%s" % (1000 + j, "ACGTACGTACGTACGTACGTACGTACGTACGTA"), ref.predictBaseRTVal[j],
                           ref.predictZygosityRTVal[j], ref.predictVarTypeRTVal[j], ref.predictIndelLengthRTVal[j], True, None)
               for j in range(n)]
        assert body == [e for e in exp if e is not None]

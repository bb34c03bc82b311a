"""CPU: the training driver's schedule (reference train.py:86-160) with a stub model."""
import pickle
import types

import numpy as np

from clairvoyante_b200 import param, synth, train, utils_v2 as U


def test_switch_rule():
    z = [10, 9, 10, 9, 10, 9]                    # + - + - + : zig-zag
    assert train.switch_needed(z) and train.switch_needed(z[::-1] + [8][:0]) is not None
    assert train.switch_needed([9, 10, 9, 10, 9, 10])
    assert not train.switch_needed([10, 9, 8, 7, 6, 5])       # steady improvement
    assert not train.switch_needed([10, 9, 10, 9, 8, 7])
    assert train.switch_needed([5, 5, 1, 2, 3, 4])            # oldest difference exactly zero (train.py:147-148)
    assert not train.switch_needed([1, 2, 3])


def test_next_batch_size(monkeypatch):
    monkeypatch.setattr(param, "trainBatchSize", 100)
    monkeypatch.setattr(param, "predictBatchSize", 10)
    vs = 451
    assert train.next_batch_size(0, vs) == 100 and train.next_batch_size(400, vs) == 51
    assert train.next_batch_size(451, vs) == 9 and train.next_batch_size(460, vs) == 10


class _Stub(object):
    def __init__(self):
        self.trained, self.validated, self.lr, self.lam, self.saved = [], [], 1e-3, 1e-3, []
        self.losses = iter([10, 9, 10, 9, 10, 9, 8, 9, 8, 9, 8, 9, 8, 9, 8, 9, 8, 9, 8, 9])
        self._cur = 0
    def setLearningRate(self, v=None):
        self.lr = self.lr * 0.1 if v is None else v
        return self.lr
    def setL2RegularizationLambda(self, v=None):
        self.lam = self.lam * 0.1 if v is None else v
        return self.lam
    def trainNoRT(self, X, Y):
        self.trained.append(len(X)); self.trainLossRTVal = 1.0; self.trainSummaryRTVal = {}
    def getLossNoRT(self, X, Y):
        self.validated.append(len(X)); self.getLossLossRTVal = 0.0
    def getLoss(self, X, Y):
        self.validated.append(len(X)); return float(next(self.losses))
    def saveParameters(self, fn):
        self.saved.append(fn)
    def predict(self, X):
        n = len(X)
        return (np.zeros((n, 4), np.float32), np.zeros((n, 2), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 6), np.float32))
    def summaryFileWriter(self, d):
        raise AssertionError


def test_train_all_schedule(tmp_path, monkeypatch):
    monkeypatch.setattr(param, "trainBatchSize", 100)
    monkeypatch.setattr(param, "predictBatchSize", 10)
    monkeypatch.setattr(param, "bloscBlockSize", 50)
    total = 503
    X = synth.make_sites(total, 1); Y = synth.make_labels(total, 1).astype(np.float64)
    xb = [U.pack_array(X[i:i + 50]) for i in range(0, total, 50)]
    yb = [U.pack_array(Y[i:i + 50]) for i in range(0, total, 50)]
    fn = str(tmp_path / "d.bin")
    with open(fn, "wb") as fh:
        for p in (total, xb, yb, []):
            pickle.dump(p, fh)
    m = _Stub()
    args = types.SimpleNamespace(bin_fn=fn, tensor_fn=None, var_fn=None, bed_fn=None, chkpnt_fn=None, learning_rate=1e-3, lambd=1e-3,
                                 ochk_prefix=str(tmp_path / "ck" / "m"), olog_dir=None, v2=False, v3=True, slim=False)
    train.TrainAll(args, m, U)
    # trainingTotal = 452, validationStart = 453.  A batch is TRAINED iff the dataset pointer *after* fetching it is
    # still < validationStart (train.py:88-91 tests datasetPtr, which already includes the batch in flight), so the
    # short batch [400,453) that ends exactly at validationStart is validated, not trained -- reference behaviour.
    per_epoch_train = [100, 100, 100, 100]
    epochs = len(m.saved)
    assert m.trained == per_epoch_train * epochs
    assert m.validated[:7] == [53, 7, 10, 10, 10, 10, 3]
    # validation losses 10,9,10,9,10,9 zig-zag -> first switch after epoch 6, next after 6 more, third ends training
    assert epochs == 18 and abs(m.lr - 1e-5) < 1e-12 and abs(m.lam - 1e-5) < 1e-12
    assert m.saved[0].endswith("m-000001") and m.saved[-1].endswith("m-000018")


def _dataset(tmp_path, total):
    X = synth.make_sites(total, 1); Y = synth.make_labels(total, 1).astype(np.float64)
    xb = [U.pack_array(X[i:i + 50]) for i in range(0, total, 50)]
    yb = [U.pack_array(Y[i:i + 50]) for i in range(0, total, 50)]
    fn = str(tmp_path / "d.bin")
    with open(fn, "wb") as fh:
        for p in (total, xb, yb, []):
            pickle.dump(p, fh)
    return fn, X, Y


def test_train_nonstop_runs_to_max_epoch_and_resumes(tmp_path, monkeypatch):
    """trainNonstop.py:78-124: same epoch loop, no schedule, a checkpoint per epoch until maxEpoch; resume epoch = the
    checkpoint suffix + 1"""
    from clairvoyante_b200 import trainNonstop
    monkeypatch.setattr(param, "trainBatchSize", 100)
    monkeypatch.setattr(param, "predictBatchSize", 10)
    monkeypatch.setattr(param, "bloscBlockSize", 50)
    monkeypatch.setattr(param, "maxEpoch", 5)
    fn, _, _ = _dataset(tmp_path, 503)
    m = _Stub()
    args = types.SimpleNamespace(bin_fn=fn, tensor_fn=None, var_fn=None, bed_fn=None, chkpnt_fn=None, learning_rate=2e-3, lambd=3e-3,
                                 ochk_prefix=str(tmp_path / "ck" / "m"), olog_dir=None, v2=False, v3=True, slim=False)
    trainNonstop.TrainAll(args, m, U)
    assert [s[-8:] for s in m.saved] == ["m-000001", "m-000002", "m-000003", "m-000004"]
    assert m.trained == [100, 100, 100, 100] * 4 and m.lr == 2e-3 and m.lam == 3e-3      # never decayed
    m2 = _Stub()
    args.chkpnt_fn = str(tmp_path / "ck" / "m-000003")
    trainNonstop.TrainAll(args, m2, U)
    assert [s[-8:] for s in m2.saved] == ["m-000004"]


def test_evaluate_report(tmp_path, monkeypatch):
    """evaluate.py:55-110 with a model that answers from the labels: perfect on three heads, second-best on base change
    for every third site"""
    from clairvoyante_b200 import evaluate
    monkeypatch.setattr(param, "predictBatchSize", 64)
    monkeypatch.setattr(param, "bloscBlockSize", 50)
    total = 333
    fn, X, Y = _dataset(tmp_path, total)

    class Oracle(object):
        def __init__(self):
            self.ptr = 0; self.calls = []
        def predict(self, Xb):
            n = len(Xb); y = Y[self.ptr:self.ptr + n].astype(np.float32); idx = np.arange(self.ptr, self.ptr + n)
            self.ptr += n; self.calls.append(n)
            truth = np.argmax(y[:, 0:4], axis=1)                  # (all-zero rows count as class 0, like np.argmax in :81)
            base = np.full((n, 4), 0.1, np.float32)
            base[np.arange(n), truth] = 0.6
            worse = idx % 3 == 0                                  # push another class above the true one
            other = (truth + 1) % 4
            base[worse, other[worse]] = 0.9
            return base, y[:, 4:6], y[:, 6:10], y[:, 10:16]

    m = Oracle()
    args = types.SimpleNamespace(bin_fn=fn, tensor_fn=None, var_fn=None, bed_fn=None)
    res = evaluate.Test(args, m, U)
    assert m.calls == [64] * 5 + [13]
    assert res["all"] == total and res["top2"] == total and res["top1"] == total - len(range(0, total, 3))
    for key, lo, hi in (("zygosity", 4, 6), ("varType", 6, 10), ("indelLength", 10, 16)):
        ed = res[key]
        assert ed.sum() == total and np.trace(ed) == total
        assert np.array_equal(np.diag(ed), np.bincount(np.argmax(Y[:, lo:hi], axis=1), minlength=hi - lo))


def test_train_without_validation_skips_the_final_partial_batch(tmp_path, monkeypatch):
    """trainWithoutValidationNonstop.py:84-105: every batch trains, except that the batch returned together with the
    end-of-data flag closes the epoch untrained (reference behaviour)"""
    from clairvoyante_b200 import trainWithoutValidationNonstop as T
    monkeypatch.setattr(param, "trainBatchSize", 100)
    monkeypatch.setattr(param, "bloscBlockSize", 50)
    monkeypatch.setattr(param, "maxEpoch", 3)
    fn, _, _ = _dataset(tmp_path, 530)
    m = _Stub()
    args = types.SimpleNamespace(bin_fn=fn, tensor_fn=None, var_fn=None, bed_fn=None, chkpnt_fn=None, learning_rate=1e-3, lambd=1e-3,
                                 ochk_prefix=str(tmp_path / "m"), olog_dir=None, v2=False, v3=True, slim=False)
    T.TrainAll(args, m, U)
    assert m.trained == [100, 100, 100, 100, 100] * 2 and m.validated == []
    assert [s[-8:] for s in m.saved] == ["m-000001", "m-000002"]


def test_cal_train_dev_diff_books_every_site_once(tmp_path, monkeypatch):
    """calTrainDevDiff.py:44-76 with a model whose loss is the number of sites in the batch: the two sums cover the set
    exactly once, split where the reference's pointer test splits them"""
    from clairvoyante_b200 import calTrainDevDiff as C
    monkeypatch.setattr(param, "predictBatchSize", 10)
    monkeypatch.setattr(param, "bloscBlockSize", 50)
    total = 503
    fn, _, _ = _dataset(tmp_path, total)

    class Counter(object):
        def __init__(self):
            self.restored, self.sizes = [], []
        def restoreParameters(self, fn):
            self.restored.append(os.path.basename(fn))
        def getLossNoRT(self, X, Y):
            self.sizes.append(len(X)); self.getLossLossRTVal = float(len(X))

    import os
    m = Counter()
    args = types.SimpleNamespace(bin_fn=fn, tensor_fn=None, var_fn=None, bed_fn=None, chkpnt_fn=["ck-000001", "ck-000002"])
    res = C.CalcAll(args, m, U)
    assert m.restored == ["ck-000001", "ck-000002"] and len(res) == 2
    trainingTotal = int(total * 0.9); validationStart = trainingTotal + 1
    per_ckpt = m.sizes[:len(m.sizes) // 2]
    assert sum(per_ckpt) == total and per_ckpt[:3] == [10, 10, 10]
    name, tr, va = res[0]
    assert abs(tr * trainingTotal + va * (total - validationStart) - total) < 1e-6
    # batches are booked by the pointer after the NEXT fetch: the batch that ends exactly at validationStart counts as validation
    assert abs(tr * trainingTotal - 450) < 1e-6 and abs(va * (total - validationStart) - 53) < 1e-6

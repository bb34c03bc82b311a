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

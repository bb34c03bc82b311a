"""`Clairvoyante` model object: the reference's seam (clairvoyante/clairvoyante_v3.py:5-283,
clairvoyante_v3_slim.py:5-260) re-implemented on libcvb200.so.

Same constructor kwargs, public attributes and method names as the reference class, so
callVar.py / train.py / evaluate.py drive it unchanged.  Every method that was a
`session.run(...)` is one C-ABI call (include/cvb200.h); NumPy arrays go in and fresh
NumPy arrays come out (the reference returns new arrays per call and the drivers rely on
that: callVar.py:199-205).  ctypes releases the GIL during the call, so the drivers'
`Thread(target=m.predictNoRT)` overlap keeps working.
"""
import ctypes
import json
import os

import numpy as np

from . import _lib, tf_bundle, initializers, param

_VARIANT_ID = {"v3": 0, "v3_slim": 1}
COMPUTE_MODES = {"fp32": 0, "fp16x3": 1, "fp16": 2}
TRAIN_MODES = {"fp32": 0, "bf16x3": 1, "bf16": 2}


def _default_device():
    for k in ("CVB_DEVICE", "LOCAL_RANK"):
        if os.environ.get(k):
            return int(os.environ[k])
    return 0


def _f32c(a, shape_tail):
    a = np.ascontiguousarray(a, dtype=np.float32)
    n = a.shape[0] if a.ndim > 0 else 0
    if a.size != n * int(np.prod(shape_tail)):
        raise ValueError("expected shape (N,%s), got %s" % (",".join(map(str, shape_tail)), a.shape))
    return a, n


class SummaryWriter(object):
    """Stand-in for tf.summary.FileWriter (clairvoyante_v3.py:253-255): one JSON line per
    add_summary(summary, step) in <logsPath>/summaries.jsonl."""

    def __init__(self, logsPath):
        os.makedirs(logsPath, exist_ok=True)
        self._fh = open(os.path.join(logsPath, "summaries.jsonl"), "a")

    def add_summary(self, summary, step):
        rec = dict(summary or {})
        rec["step"] = int(step)
        self._fh.write(json.dumps(rec) + "\n")
        self._fh.flush()

    def close(self):
        self._fh.close()


class ClairvoyanteBase(object):
    VARIANT = "v3"

    def _common_init(self, initialLearningRate, learningRateDecay, dropoutRateFC4, dropoutRateFC5,
                     l2RegularizationLambda, l2RegularizationLambdaDecay, device):
        self.learningRateVal = initialLearningRate
        self.learningRateDecay = learningRateDecay
        self.dropoutRateFC4Val = dropoutRateFC4
        self.dropoutRateFC5Val = dropoutRateFC5
        self.l2RegularizationLambdaVal = l2RegularizationLambda
        self.l2RegularizationLambdaDecay = l2RegularizationLambdaDecay
        self.trainLossRTVal = None; self.trainSummaryRTVal = None; self.getLossLossRTVal = None
        self.predictBaseRTVal = None; self.predictZygosityRTVal = None
        self.predictVarTypeRTVal = None; self.predictIndelLengthRTVal = None
        self._lib = _lib.load()
        self.device = _default_device() if device is None else int(device)
        h = ctypes.c_void_p()
        _lib.check(self._lib.cvb_create(_VARIANT_ID[self.VARIANT], self.device, ctypes.byref(h)))
        self._h = h
        # arithmetic mode: defaults to the tensor-core path (split-fp16 operands, fp32 accumulate,
        # logits within 1e-3 of fp64); CVB_COMPUTE=fp32 selects the all-SIMT fp32 kernels.
        mode = os.environ.get("CVB_COMPUTE", "auto")
        if mode == "auto":
            mode = "fp16x3"        # v3: conv2/conv3/FC4/tail on tcgen05; v3_slim: conv3 on tcgen05
        self.computeMode = None
        self.setComputeMode(mode)
        # arithmetic of the training step's large contractions (cvb200.h CVB_TRAIN_*): split-bf16 on tcgen05 by default,
        # CVB_TRAIN=fp32 selects the all-SIMT kernels, CVB_TRAIN=bf16 plain bf16 operands
        self.trainMode = None
        self.setTrainMode(os.environ.get("CVB_TRAIN", "bf16x3"))
        if not 0.0 <= float(dropoutRateFC5) < 1.0:
            raise ValueError("dropoutRateFC5 must be in [0, 1)")
        _lib.check(self._lib.cvb_set_dropout_fc5(self._h, float(dropoutRateFC5)))
        self._dropout_calls = 0
        self._seed = int.from_bytes(os.urandom(8), "little")   # reference dropout is unseeded (selu.py:55)

    # ---- lifecycle -------------------------------------------------------------------
    def init(self, seed=None):
        """init_op (clairvoyante_v3.py:177-178): reference initialisers, zero Adam slots, step 0."""
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        _lib.check(self._lib.cvb_init_weights(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.cvb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters ------------------------------------------------------------------
    def _set(self, name, slot, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        _lib.check(self._lib.cvb_set_variable(self._h, name.encode(), slot, a.ctypes.data, a.size))

    def _get(self, name, slot, shape):
        a = np.empty(shape, np.float32)
        _lib.check(self._lib.cvb_get_variable(self._h, name.encode(), slot, a.ctypes.data, a.size))
        return a

    def setWeights(self, weights):
        for name, shape in initializers.variable_shapes(self.VARIANT):
            w = np.asarray(weights[name], np.float32)
            if w.shape != tuple(shape):
                raise ValueError("variable %s: expected shape %s, got %s" % (name, shape, w.shape))
            self._set(name, 0, w)

    def getWeights(self):
        return {name: self._get(name, 0, shape) for name, shape in initializers.variable_shapes(self.VARIANT)}

    def getGradients(self):
        out = {}
        for name, shape in initializers.variable_shapes(self.VARIANT):
            a = np.empty(shape, np.float32)
            _lib.check(self._lib.cvb_get_gradient(self._h, name.encode(), a.ctypes.data, a.size))
            out[name] = a
        return out

    def setComputeMode(self, mode):
        _lib.check(self._lib.cvb_set_compute_mode(self._h, COMPUTE_MODES[mode]))
        self.computeMode = mode

    def setTrainMode(self, mode):
        _lib.check(self._lib.cvb_set_train_mode(self._h, TRAIN_MODES[mode]))
        self.trainMode = mode

    @staticmethod
    def _ckpt_path(fn):
        return fn + ".cvb.npz"

    def saveParameters(self, fn):
        """tf.train.Saver().save (clairvoyante_v3.py:243-246).  Writes two equivalent containers:
          * `fn.index` + `fn.data-00000-of-00001`: a TensorFlow V2 checkpoint bundle (tf_bundle.py) with the names
            tf.train.Saver() stores for this graph -- the 18 variables, their Adam slots `<name>/Adam`, `<name>/Adam_1`
            and `beta1_power` / `beta2_power` (= beta**(step+1), TF-1.x AdamOptimizer) -- so the reference's
            `restoreParameters` can load a model trained here;
          * `fn.cvb.npz`: the same arrays plus the integer `step`."""
        d = {}
        for name, shape in initializers.variable_shapes(self.VARIANT):
            d[name] = self._get(name, 0, shape)
            d[name + "/Adam"] = self._get(name, 1, shape)
            d[name + "/Adam_1"] = self._get(name, 2, shape)
        t = ctypes.c_int64()
        _lib.check(self._lib.cvb_get_step(self._h, ctypes.byref(t)))
        os.makedirs(os.path.dirname(os.path.abspath(fn)), exist_ok=True)
        tf_bundle.write_bundle(fn, dict(d, beta1_power=np.float32(0.9 ** (t.value + 1)),
                                        beta2_power=np.float32(0.999 ** (t.value + 1))))
        d["step"] = np.int64(t.value)
        np.savez(self._ckpt_path(fn), **d)
        # the reference's callVarBam.py / callVarBamParallel.py test for `<prefix>.meta` before they restore (CheckFileExist
        # with sfx='.meta'); tf.train.Saver().restore itself reads only .index / .data, so an empty graph-less .meta and the
        # usual `checkpoint` state file are enough for a checkpoint written here to pass those entry points
        if not os.path.exists(fn + ".meta"):
            open(fn + ".meta", "wb").close()
        with open(os.path.join(os.path.dirname(os.path.abspath(fn)), "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (os.path.basename(fn), os.path.basename(fn)))

    def getStep(self):
        """number of Adam updates applied so far (what TF keeps as beta1_power / beta2_power)"""
        t = ctypes.c_int64()
        _lib.check(self._lib.cvb_get_step(self._h, ctypes.byref(t)))
        return int(t.value)

    def restoreParameters(self, fn):
        """tf.train.Saver().restore (clairvoyante_v3.py:248-251): a TensorFlow V2 bundle prefix (the reference's
        published `trainedModels`, or one written by saveParameters) or a `.cvb.npz`."""
        if tf_bundle.is_bundle(fn):
            d = tf_bundle.read_bundle(fn)
            step = 0
            if "beta2_power" in d:                        # beta2**(step+1); beta1_power underflows float32 after ~800 steps
                b2 = float(np.asarray(d["beta2_power"]).reshape(-1)[0])
                step = max(0, int(round(np.log(b2) / np.log(0.999))) - 1) if b2 > 0 else 1 << 20
        else:
            path = None
            for cand in (self._ckpt_path(fn), fn, fn + ".npz"):
                if os.path.isfile(cand):
                    path = cand
                    break
            if path is None:
                raise IOError("checkpoint not found: %s(.index | .cvb.npz)" % fn)
            with np.load(path) as z:
                d = {k: z[k] for k in z.files}
            step = int(d["step"]) if "step" in d else 0
        for name, shape in initializers.variable_shapes(self.VARIANT):
            if name not in d:
                raise KeyError("checkpoint %s has no variable %s" % (fn, name))
            if int(np.prod(d[name].shape)) != int(np.prod(shape)):
                raise ValueError("checkpoint %s: %s has shape %r, this model needs %r (wrong variant?)"
                                 % (fn, name, d[name].shape, tuple(shape)))
            self._set(name, 0, np.asarray(d[name], np.float32).reshape(shape))
            if name + "/Adam" in d and name + "/Adam_1" in d:
                self._set(name, 1, np.asarray(d[name + "/Adam"], np.float32).reshape(shape))
                self._set(name, 2, np.asarray(d[name + "/Adam_1"], np.float32).reshape(shape))
        _lib.check(self._lib.cvb_set_step(self._h, step))

    def summaryFileWriter(self, logsPath):
        return SummaryWriter(logsPath)

    # ---- hyper-parameters (clairvoyante_v3.py:229-241) -------------------------------------
    def setLearningRate(self, learningRate=None):
        if learningRate == None:
            self.learningRateVal = self.learningRateVal * self.learningRateDecay
        else:
            self.learningRateVal = learningRate
        return self.learningRateVal

    def setL2RegularizationLambda(self, l2RegularizationLambda=None):
        if l2RegularizationLambda == None:
            self.l2RegularizationLambdaVal = self.l2RegularizationLambdaVal * self.l2RegularizationLambdaDecay
        else:
            self.l2RegularizationLambdaVal = l2RegularizationLambda
        return self.l2RegularizationLambdaVal

    # ---- inference (clairvoyante_v3.py:257-280) --------------------------------------------
    def _predict(self, XArray, want_logits=False):
        # narrow feeds (cvb200.h): a CountBatch's raw uint8 / int16 counts (utils_v2.GetTensor attaches them), an explicit
        # uint8 / int16 array of RAW counts, or float16 values -- a quarter / a half of the host->device bytes, same bits out
        counts = getattr(XArray, "counts", None)
        if counts is not None and counts.shape[0] == XArray.shape[0]:
            XArray = counts
        fn = None
        if isinstance(XArray, np.ndarray) and XArray.dtype in (np.float16, np.int16, np.uint8):
            fn = {np.dtype(np.float16): self._lib.cvb_predict_host_f16, np.dtype(np.int16): self._lib.cvb_predict_host_counts_i16,
                  np.dtype(np.uint8): self._lib.cvb_predict_host_counts_u8}[XArray.dtype]
            x = np.ascontiguousarray(XArray)
            n = x.shape[0] if x.ndim > 0 else 0
            if x.size != n * int(np.prod(self.inputShape)):
                raise ValueError("expected shape (N,33,4,4), got %s" % (x.shape,))
        else:
            x, n = _f32c(XArray, self.inputShape)
            fn = self._lib.cvb_predict_host
        base = np.empty((n, 4), np.float32); z = np.empty((n, 2), np.float32)
        t = np.empty((n, 4), np.float32); l = np.empty((n, 6), np.float32)
        lg = np.empty((n, 16), np.float32) if want_logits else None
        _lib.check(fn(self._h, x.ctypes.data, n, base.ctypes.data, z.ctypes.data, t.ctypes.data,
                      l.ctypes.data, lg.ctypes.data if want_logits else None))
        return base, z, t, l, lg

    def predict(self, XArray):
        return self._predict(XArray)[:4]

    def predictNoRT(self, XArray):
        self.predictBaseRTVal = None; self.predictZygosityRTVal = None
        self.predictVarTypeRTVal = None; self.predictIndelLengthRTVal = None
        self.predictBaseRTVal, self.predictZygosityRTVal, self.predictVarTypeRTVal, self.predictIndelLengthRTVal \
            = self.predict(XArray)

    # ---- the same call in two halves: several small batches in flight (cvb_predict_submit / cvb_predict_collect) ----------
    def predictSubmit(self, XArray, want_logits=False):
        """Start predict(XArray) and return a ticket; XArray may be reused at once.  Up to four tickets may be outstanding;
        predictCollect returns them in any order (they complete in submission order).  A batch must fit one device pass
        (18,944 sites for v3, 33,152 for v3_slim) -- larger ones go through predict, which pipelines its own chunks."""
        x, kind, n = self._x_arg(XArray)
        t = ctypes.c_int()
        _lib.check(self._lib.cvb_predict_submit(self._h, x.ctypes.data, kind, n, 1 if want_logits else 0, ctypes.byref(t)))
        return (t.value, n, bool(want_logits))

    def predictCollect(self, ticket):
        """(base, zygosity, varType, indelLength[, logits16]) of a ticket from predictSubmit -- the arrays predict returns"""
        t, n, want_logits = ticket
        base = np.empty((n, 4), np.float32); z = np.empty((n, 2), np.float32)
        v = np.empty((n, 4), np.float32); l = np.empty((n, 6), np.float32)
        lg = np.empty((n, 16), np.float32) if want_logits else None
        _lib.check(self._lib.cvb_predict_collect(self._h, t, base.ctypes.data, z.ctypes.data, v.ctypes.data, l.ctypes.data,
                                                 lg.ctypes.data if want_logits else None))
        return (base, z, v, l, lg) if want_logits else (base, z, v, l)

    def predictStream(self, batches, depth=3):
        """predict over an iterable of small batches with `depth` (<= 4) of them in flight: yields predict's four arrays per
        batch, in order.  The staging copy of batch k+1 and the copy-out of batch k-1 overlap the kernels of batch k."""
        depth = max(1, min(int(depth), 4))
        pending = []
        try:
            for X in batches:
                if len(pending) == depth:
                    yield self.predictCollect(pending.pop(0))
                pending.append(self.predictSubmit(X))
            while pending:
                yield self.predictCollect(pending.pop(0))
        finally:                                   # a consumer that stops early must not leave tickets outstanding
            while pending:
                self.predictCollect(pending.pop(0))

    def predictLogits(self, XArray):
        """(out16, logits16): extension used by the parity tests (base head pre-sigmoid)."""
        base, z, t, l, lg = self._predict(XArray, want_logits=True)
        return np.concatenate([base, z, t, l], axis=1), lg

    def predictDevice(self, x_ptr, n, out16_ptr, logits16_ptr=None, stream=None):
        """Device-resident batch (pointers from e.g. torch.Tensor.data_ptr()); asynchronous."""
        _lib.check(self._lib.cvb_predict_device(self._h, x_ptr, n, out16_ptr, logits16_ptr, stream))

    X_KINDS = {"f32": 0, "f16": 1, "i16": 2, "u8": 3}

    def predictDeviceX(self, x_ptr, kind, n, out16_ptr, logits16_ptr=None, stream=None):
        """predictDevice for a device buffer of fp16 values ('f16') or RAW int16 / uint8 counts ('i16', 'u8')"""
        _lib.check(self._lib.cvb_predict_device_x(self._h, x_ptr, self.X_KINDS[kind], n, out16_ptr, logits16_ptr, stream))

    def debugRead(self, which, n_sites):
        """Intermediate of the last device pass: 'p2' | 'p3' | 'h4' (test aid)."""
        per = {"v3": dict(p2=28 * 128, p3=4608, h4=336), "v3_slim": dict(p2=37 * 64, p3=4224, h4=36)}[self.VARIANT][which.split("_")[0]]
        a = np.empty((n_sites, per), np.float32)
        sel = dict(p2=0, p3=1, h4=2, p2_split=3, p3_split=4)[which]
        _lib.check(self._lib.cvb_debug_read(self._h, sel, a.ctypes.data, a.size))
        return a

    def profileBegin(self):
        _lib.check(self._lib.cvb_profile_begin(self._h))

    def profileRead(self):
        """{kernel: (total_ms, launches)} since profileBegin(); synchronises the device."""
        ms = (ctypes.c_double * 5)()
        cnt = (ctypes.c_int64 * 5)()
        _lib.check(self._lib.cvb_profile_read(self._h, ms, cnt))
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(("front", "conv2", "conv3", "fc4", "tail"))}

    def kernelLaunches(self):
        return int(self._lib.cvb_kernel_launches(self._h))

    # ---- loss / training (clairvoyante_v3.py:183-227) --------------------------------------
    _X_KIND = {np.dtype(np.float32): 0, np.dtype(np.float16): 1, np.dtype(np.int16): 2, np.dtype(np.uint8): 3}

    def _x_arg(self, X):
        """(contiguous array, element kind, n): a CountBatch's raw counts or an explicit uint8 / int16 / float16 array travel
        narrow (cvb200.h CVB_X_*), anything else as float32"""
        counts = getattr(X, "counts", None)
        if counts is not None and counts.shape[0] == X.shape[0]:
            X = counts
        if isinstance(X, np.ndarray) and X.dtype in (np.float16, np.int16, np.uint8):
            x = np.ascontiguousarray(X)
            n = x.shape[0] if x.ndim > 0 else 0
            if x.size != n * int(np.prod(self.inputShape)):
                raise ValueError("expected shape (N,33,4,4), got %s" % (x.shape,))
            return x, self._X_KIND[x.dtype], n
        x, n = _f32c(X, self.inputShape)
        return x, 0, n

    def getLoss(self, batchX, batchY):
        x, kind, n = self._x_arg(batchX)
        y, ny = _f32c(batchY, (16,))
        if n != ny:
            raise ValueError("X/Y batch mismatch %d/%d" % (n, ny))
        loss = ctypes.c_float()
        _lib.check(self._lib.cvb_loss_host_x(self._h, x.ctypes.data, kind, y.ctypes.data, n, ctypes.byref(loss)))
        return np.float32(loss.value)

    def getLossNoRT(self, batchX, batchY):
        self.getLossLossRTVal = None
        self.getLossLossRTVal = self.getLoss(batchX, batchY)

    def _train_step(self, batchX, batchY, apply_update=1, seed=None):
        x, kind, n = self._x_arg(batchX)
        y, ny = _f32c(batchY, (16,))
        if n != ny:
            raise ValueError("X/Y batch mismatch %d/%d" % (n, ny))
        if seed is None:
            self._dropout_calls += 1
            seed = (self._seed + self._dropout_calls * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        l5 = (ctypes.c_float * 8)()
        _lib.check(self._lib.cvb_train_step_host_x(self._h, x.ctypes.data, kind, y.ctypes.data, n,
                                                   self.learningRateVal, self.l2RegularizationLambdaVal,
                                                   self.dropoutRateFC4Val, seed, apply_update, l5))
        summary = dict(learning_rate=float(self.learningRateVal), l2Lambda=float(self.l2RegularizationLambdaVal),
                       loss=float(l5[0]), loss1=float(l5[1]), loss2=float(l5[2]), loss3=float(l5[3]),
                       loss4=float(l5[4]), lossL2=float(l5[5]))
        return np.float32(l5[0]), summary

    def applyAdam(self):
        """finish a step whose gradients were left in the gradient buffer (apply_update=0): loss + TF-1.x Adam"""
        l6 = (ctypes.c_float * 8)()
        _lib.check(self._lib.cvb_apply_adam(self._h, self.learningRateVal, self.l2RegularizationLambdaVal, l6))
        summary = dict(learning_rate=float(self.learningRateVal), l2Lambda=float(self.l2RegularizationLambdaVal),
                       loss=float(l6[0]), loss1=float(l6[1]), loss2=float(l6[2]), loss3=float(l6[3]),
                       loss4=float(l6[4]), lossL2=float(l6[5]))
        return np.float32(l6[0]), summary

    def train(self, batchX, batchY):
        return self._train_step(batchX, batchY)

    def trainNoRT(self, batchX, batchY):
        self.trainLossRTVal = None; self.trainSummaryRTVal = None
        self.trainLossRTVal, self.trainSummaryRTVal = self._train_step(batchX, batchY)


class ClairvoyanteV3(ClairvoyanteBase):
    VARIANT = "v3"

    def __init__(self, inputShape=(2 * param.flankingBaseNum + 1, 4, param.matrixNum),
                 outputShape1=(4,), outputShape2=(2,), outputShape3=(4,), outputShape4=(6,),
                 kernelSize1=(1, 4), kernelSize2=(2, 4), kernelSize3=(3, 4),
                 pollSize1=(5, 1), pollSize2=(4, 1), pollSize3=(3, 1),
                 numFeature1=16, numFeature2=32, numFeature3=48,
                 hiddenLayerUnits4=336, hiddenLayerUnits5=168,
                 initialLearningRate=param.initialLearningRate, learningRateDecay=param.learningRateDecay,
                 dropoutRateFC4=param.dropoutRateFC4, dropoutRateFC5=param.dropoutRateFC5,
                 l2RegularizationLambda=param.l2RegularizationLambda,
                 l2RegularizationLambdaDecay=param.l2RegularizationLambdaDecay, device=None):
        geom = (tuple(inputShape), tuple(outputShape1), tuple(outputShape2), tuple(outputShape3), tuple(outputShape4),
                tuple(kernelSize1), tuple(kernelSize2), tuple(kernelSize3), tuple(pollSize1), tuple(pollSize2),
                tuple(pollSize3), numFeature1, numFeature2, numFeature3, hiddenLayerUnits4, hiddenLayerUnits5)
        if geom != ((33, 4, 4), (4,), (2,), (4,), (6,), (1, 4), (2, 4), (3, 4), (5, 1), (4, 1), (3, 1), 16, 32, 48, 336, 168):
            raise NotImplementedError("libcvb200 kernels are compiled for the reference v3 geometry "
                                      "(clairvoyante_v3.py:7-12); got %r" % (geom,))
        self.inputShape = inputShape
        self.outputShape1 = outputShape1; self.outputShape2 = outputShape2
        self.outputShape3 = outputShape3; self.outputShape4 = outputShape4
        self.kernelSize1 = kernelSize1; self.kernelSize2 = kernelSize2; self.kernelSize3 = kernelSize3
        self.pollSize1 = pollSize1; self.pollSize2 = pollSize2; self.pollSize3 = pollSize3
        self.numFeature1 = numFeature1; self.numFeature2 = numFeature2; self.numFeature3 = numFeature3
        self.hiddenLayerUnits4 = hiddenLayerUnits4; self.hiddenLayerUnits5 = hiddenLayerUnits5
        self._common_init(initialLearningRate, learningRateDecay, dropoutRateFC4, dropoutRateFC5,
                          l2RegularizationLambda, l2RegularizationLambdaDecay, device)


class ClairvoyanteV3Slim(ClairvoyanteBase):
    VARIANT = "v3_slim"

    def __init__(self, inputShape=(2 * param.flankingBaseNum + 1, 4, param.matrixNum),
                 outputShape1=(4,), outputShape2=(2,), outputShape3=(4,), outputShape4=(6,),
                 kernelSize1=(1, 4), kernelSize2=(3, 4), kernelSize3=(5, 4),
                 numFeature1=8, numFeature2=16, numFeature3=32,
                 hiddenLayerUnits4=36, hiddenLayerUnits5=18,
                 initialLearningRate=param.initialLearningRate, learningRateDecay=param.learningRateDecay,
                 dropoutRateFC4=param.dropoutRateFC4, dropoutRateFC5=param.dropoutRateFC5,
                 l2RegularizationLambda=param.l2RegularizationLambda,
                 l2RegularizationLambdaDecay=param.l2RegularizationLambdaDecay, device=None):
        geom = (tuple(inputShape), tuple(outputShape1), tuple(outputShape2), tuple(outputShape3), tuple(outputShape4),
                tuple(kernelSize1), tuple(kernelSize2), tuple(kernelSize3), numFeature1, numFeature2, numFeature3,
                hiddenLayerUnits4, hiddenLayerUnits5)
        if geom != ((33, 4, 4), (4,), (2,), (4,), (6,), (1, 4), (3, 4), (5, 4), 8, 16, 32, 36, 18):
            raise NotImplementedError("libcvb200 kernels are compiled for the reference v3_slim geometry "
                                      "(clairvoyante_v3_slim.py:7-11); got %r" % (geom,))
        self.inputShape = inputShape
        self.outputShape1 = outputShape1; self.outputShape2 = outputShape2
        self.outputShape3 = outputShape3; self.outputShape4 = outputShape4
        self.kernelSize1 = kernelSize1; self.kernelSize2 = kernelSize2; self.kernelSize3 = kernelSize3
        self.numFeature1 = numFeature1; self.numFeature2 = numFeature2; self.numFeature3 = numFeature3
        self.hiddenLayerUnits4 = hiddenLayerUnits4; self.hiddenLayerUnits5 = hiddenLayerUnits5
        self._common_init(initialLearningRate, learningRateDecay, dropoutRateFC4, dropoutRateFC5,
                          l2RegularizationLambda, l2RegularizationLambdaDecay, device)

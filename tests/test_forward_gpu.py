"""GPU parity: CUDA forward (through the C-ABI / Clairvoyante object) vs the float64 oracle.

Bar (BASELINE.json north_star): per-head argmax identical, |logit - oracle_fp64| <= 1e-3
for the fp32 configuration.  The base-change head is compared on its pre-sigmoid logits:
with seeded random weights its sigmoid saturates to exactly 1.0f for a fifth of the sites
(SURVEY.md 7.2), which makes argmax-of-probability a tie-break lottery in ANY fp32
implementation including the reference's."""
import os

import numpy as np
import pytest

from clairvoyante_b200 import initializers as I, synth
from oracle import cv_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3


# (variant, compute mode): every parity test runs on each arithmetic path the library ships
CASES = [("v3", "fp16x3"), ("v3", "fp32"), ("v3_slim", "fp16x3"), ("v3_slim", "fp32"), ("v3_slim", "fp16")]
# "fp16" = BASELINE configs[2] (v3_slim, plain fp16 operands: one MMA term, activations stored as one fp16 plane, fp32
# accumulate).  Its stated tolerance: |logit - oracle_fp64| <= 2e-3 * max(1, max |logit|) -- fp16's 2^-11 relative rounding on
# every activation and weight through five layers (measured: 5e-4 of the largest logit); the fp32-equivalent modes keep the
# 1e-3 absolute bar of north_star.
FP16_REL = 2e-3


def _tol(mode, ref_logits):
    if mode != "fp16" or not len(ref_logits):
        return TOL
    return FP16_REL * max(1.0, float(np.abs(ref_logits).max()))


def _model(variant, W, mode=None, **kw):
    if variant == "v3":
        from clairvoyante_b200 import clairvoyante_v3 as cv
    else:
        from clairvoyante_b200 import clairvoyante_v3_slim as cv
    m = cv.Clairvoyante(**kw)
    if mode is not None:
        m.setComputeMode(mode)
    m.setWeights(W)
    return m


def _check(variant, W, x, m=None, tol=TOL, mode=None):
    own = m is None
    if own:
        m = _model(variant, W, mode)
    out16, lg = m.predictLogits(x)
    ref = O.forward(W, x, variant)
    r16 = O.out16(ref)
    assert out16.shape == (len(x), 16) and out16.dtype == np.float32
    if len(x):
        tol = _tol(mode, ref["logits"]) if tol == TOL else tol
        err = np.abs(lg - ref["logits"]).max()
        assert err <= tol, "max |logit - oracle| = %g" % err
        assert np.abs(out16 - r16).max() <= (2e-4 if mode != "fp16" else 0.25 * tol + 2e-4)
        # argmax per head; a site is exempt only if the oracle's own top-2 margin is below the tolerance
        for a, b in ((0, 4), (4, 6), (6, 10), (10, 16)):
            rl = ref["logits"][:, a:b]
            srt = np.sort(rl, 1)
            clear = (srt[:, -1] - srt[:, -2]) > 2 * tol
            assert (lg[:, a:b].argmax(1) == rl.argmax(1))[clear].all()
            # (with random initialiser weights many zygosity sites have BOTH logits saturated at SELU's
            #  lower bound -1.7581 -> exact ties in the oracle itself; those are covered by the logit bound)
    base, z, t, l = m.predict(x)
    assert base.shape == (len(x), 4) and z.shape == (len(x), 2) and t.shape == (len(x), 4) and l.shape == (len(x), 6)
    assert np.array_equal(np.concatenate([base, z, t, l], 1), out16)
    if own:
        m.close()


@pytest.mark.parametrize("variant", ["v3", "v3_slim"])
def test_default_mode_is_tensor_path(variant):
    m = _model(variant, I.init_weights(variant, 0))
    assert m.computeMode == "fp16x3"
    m.close()


@pytest.mark.parametrize("variant,mode", CASES)
def test_golden_fixture(variant, mode):
    d = np.load(os.path.join(GOLD, "forward_%s.npz" % variant))
    W = I.init_weights(variant, int(d["weight_seed"]))
    x = synth.make_sites(int(d["n"]), int(d["data_seed"]))
    m = _model(variant, W, mode)
    out16, lg = m.predictLogits(x)
    tol = _tol(mode, d["logits"])
    assert np.abs(lg - d["logits"]).max() <= tol
    assert np.abs(out16 - d["out16"]).max() <= (2e-4 if mode != "fp16" else 0.25 * tol + 2e-4)
    m.close()


@pytest.mark.parametrize("variant,mode", CASES)
@pytest.mark.parametrize("n", [0, 1, 2, 3, 5, 95, 96, 97, 127, 128, 129, 999, 1000, 1001])
def test_edge_batch_sizes(variant, mode, n):
    # N=0 is legal (utils_v2.py:56-59); 999/1000/1001 straddle predictBatchSize (param.py:12);
    # 95..129 straddle the kernels' site tiles (96 fp32 FC4, 128 tensor FC4)
    W = I.init_weights(variant, 1)
    _check(variant, W, synth.make_sites(n, 3), mode=mode)


@pytest.mark.parametrize("variant,mode", CASES)
def test_multi_chunk_and_pinned_paths(variant, mode):
    """> 1 internal chunk (14208 sites on a 148-SM part), pageable and pinned host input, device-resident input:
    all three routes must give identical bits, and shards must equal the whole."""
    import torch
    W = I.init_weights(variant, 2)
    n = 14208 * 2 + 777
    x = synth.make_sites(n, 4)
    m = _model(variant, W, mode)
    o_page, l_page = m.predictLogits(x)
    xp = torch.from_numpy(x).pin_memory()
    o_pin, l_pin = m.predictLogits(xp.numpy())
    assert np.array_equal(o_page, o_pin) and np.array_equal(l_page, l_pin)
    xd = torch.from_numpy(x).cuda()
    od = torch.empty((n, 16), dtype=torch.float32, device="cuda")
    ld = torch.empty((n, 16), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    m.predictDevice(xd.data_ptr(), n, od.data_ptr(), ld.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(od.cpu().numpy(), o_page) and np.array_equal(ld.cpu().numpy(), l_page)
    # oracle on a sample (fp64 NumPy on 33k sites would take a while)
    idx = np.r_[0:64, 9472 - 32:9472 + 32, 14208 - 32:14208 + 32, n - 64:n]
    ref = O.forward(W, x[idx], variant)
    assert np.abs(l_page[idx] - ref["logits"]).max() <= _tol(mode, ref["logits"])
    # sharding: contiguous site ranges concatenated in order == single pass
    parts = [m.predictLogits(x[a:b])[0] for a, b in ((0, 10000), (10000, 20001), (20001, n))]
    assert np.array_equal(np.concatenate(parts), o_page)
    m.close()


@pytest.mark.parametrize("variant,mode", CASES)
def test_fp16_input_is_bit_identical_for_count_tensors(variant, mode):
    """cvb_predict_host_f16 (BASELINE config 3, "fp16 I/O"): CreateTensor counts are integers <= 250, exact in fp16, so the
    half-width feed must reproduce the fp32 feed bit for bit -- pageable and pinned, across chunk boundaries, N = 0"""
    import torch
    W = I.init_weights(variant, 4)
    n = 14208 * 2 + 333
    x = synth.make_sites(n, 6)
    x[5] *= 6.0                                     # depth-250 scale values
    xh = x.astype(np.float16)
    assert np.array_equal(xh.astype(np.float32), x)
    m = _model(variant, W, mode)
    o32, l32 = m.predictLogits(x)
    o16, l16 = m.predictLogits(xh)
    assert np.array_equal(o32, o16) and np.array_equal(l32, l16)
    xp = torch.from_numpy(xh).pin_memory()
    o16p, l16p = m.predictLogits(xp.numpy())
    assert np.array_equal(o32, o16p) and np.array_equal(l32, l16p)
    base, z, t, l = m.predict(xh[:0])
    assert base.shape == (0, 4) and l.shape == (0, 6)
    # non-integer input: defined as the fp32 path applied to the widened halves
    y = (x[:500] * 0.37).astype(np.float16)
    assert np.array_equal(m.predictLogits(y)[1], m.predictLogits(y.astype(np.float32))[1])
    with pytest.raises(ValueError):
        m.predict(np.zeros((3, 33, 4, 3), np.float16))
    m.close()


@pytest.mark.parametrize("mode", ["fp16x3", "fp32"])
def test_extreme_inputs_v3(mode):
    """all-zero tensors, maximum depth (dcov cap 250, CreateTensor.py:296) and negative-heavy tensors"""
    W = I.init_weights("v3", 5)
    x = np.zeros((40, 33, 4, 4), np.float32)
    x[10:20, :, 1, 0] = 250.0
    x[20:30, :, :, 1:4] = -250.0
    x[30:40] = synth.make_sites(10, 9) * 3.0
    ref = O.forward(W, x, "v3")
    m = _model("v3", W, mode)
    out16, lg = m.predictLogits(x)
    scale = max(1.0, np.abs(ref["logits"]).max() / 100.0)    # tolerance is quoted at |logit| ~ 1e2
    assert np.abs(lg - ref["logits"]).max() <= TOL * scale
    assert np.isfinite(out16).all()
    m.close()


@pytest.mark.parametrize("mode", ["fp16x3", "fp32"])
def test_stage_intermediates_v3(mode):
    """conv stack and FC4 outputs individually against the oracle's layers"""
    W = I.init_weights("v3", 6)
    x = synth.make_sites(300, 8)          # > 2 tensor-core M-tiles of 128 rows / 120 conv3 rows
    m = _model("v3", W, mode)
    m.predictLogits(x)
    L = O.forward(W, x, "v3", return_all=True)["layers"]
    sfx = "_split" if mode == "fp16x3" else ""
    p2 = m.debugRead("p2" + sfx, 300).reshape(300, 28, 4, 32)
    assert (p2[:, 0] == 0).all() and (p2[:, 27] == 0).all()
    assert np.abs(p2[:, 1:27] - L["pool2"]).max() <= 2e-4
    p3 = m.debugRead("p3" + sfx, 300).reshape(300, 24, 4, 48)
    assert np.abs(p3 - L["pool3"]).max() <= 3e-4
    h4 = m.debugRead("h4", 300)
    assert np.abs(h4 - L["fc4"]).max() <= 5e-4
    m.close()


def test_slim_modes_switch_and_report_their_error():
    """v3_slim: fp32 SIMT, fp16x3 and plain fp16 share buffers; switching back and forth must not leak stale rows; prints the
    measured logit error of each arithmetic (the fp16 figure is what its stated tolerance is judged against)"""
    W = I.init_weights("v3_slim", 8)
    x = synth.make_sites(700, 9)
    ref = O.forward(W, x, "v3_slim")["logits"]
    m = _model("v3_slim", W, "fp32")
    for mode in ("fp16x3", "fp16", "fp32", "fp16", "fp16x3"):
        m.setComputeMode(mode)
        err = np.abs(m.predictLogits(x)[1] - ref).max()
        print("v3_slim %s: max |logit - oracle| = %.3g (max |logit| %.3g)" % (mode, err, np.abs(ref).max()))
        assert err <= _tol(mode, ref)
    m.close()


def test_mode_switch_back_and_forth():
    """the fp32 and fp16 hi/lo layouts alias the same buffers: switching must not leak stale padding rows"""
    W = I.init_weights("v3", 8)
    x = synth.make_sites(500, 9)
    ref = O.forward(W, x, "v3")["logits"]
    m = _model("v3", W, "fp32")
    for mode in ("fp16x3", "fp32", "fp16x3"):
        m.setComputeMode(mode)
        assert np.abs(m.predictLogits(x)[1] - ref).max() <= TOL
    m.close()


def test_tensor_path_tracks_weight_updates():
    """the split fp16 weight copies are rebuilt whenever the fp32 master changes"""
    Wa, Wb = I.init_weights("v3", 1), I.init_weights("v3", 2)
    x = synth.make_sites(200, 5)
    m = _model("v3", Wa, "fp16x3")
    la = m.predictLogits(x)[1]
    m.setWeights(Wb)
    lb = m.predictLogits(x)[1]
    assert np.abs(la - O.forward(Wa, x, "v3")["logits"]).max() <= TOL
    assert np.abs(lb - O.forward(Wb, x, "v3")["logits"]).max() <= TOL
    # different power-of-two pre-scales; activations grow 16x but stay inside fp16's range (the split operands
    # saturate at +-65504)
    big = {k: (v * 4.0 if k in ("fc4/kernel", "conv3/kernel") else v) for k, v in Wa.items()}
    m.setWeights(big)
    ref = O.forward(big, x, "v3")["logits"]
    assert np.abs(m.predictLogits(x)[1] - ref).max() <= TOL * max(1.0, np.abs(ref).max() / 100.0)
    m.close()


def test_thread_handoff_like_callvar():
    """callVar.py:194-212: predictNoRT on a worker thread while the main thread keeps the
    previous batch's outputs -> results must be fresh arrays per call."""
    from threading import Thread
    W = I.init_weights("v3", 7)
    m = _model("v3", W)
    xa, xb = synth.make_sites(1000, 1), synth.make_sites(1000, 2)
    m.predictNoRT(xa)
    base_a = m.predictBaseRTVal
    keep = base_a.copy()
    t = Thread(target=m.predictNoRT, args=(xb,))
    t.start(); t.join()
    assert m.predictBaseRTVal is not base_a
    assert np.array_equal(base_a, keep)
    assert not np.array_equal(m.predictBaseRTVal, keep)
    m.close()


def test_error_paths():
    W = I.init_weights("v3", 0)
    m = _model("v3", W)
    with pytest.raises(ValueError):
        m.predict(np.zeros((3, 33, 4, 3), np.float32))
    with pytest.raises(RuntimeError):
        m._set("no/such/var", 0, np.zeros(4, np.float32))
    with pytest.raises(ValueError):
        m.setWeights({k: (v if k != "fc4/bias" else v[:5]) for k, v in W.items()})
    m.close()


@pytest.mark.parametrize("variant", ["v3", "v3_slim"])
def test_c_abi_pinned_buffers_and_device_feeds(variant):
    """the routes `Clairvoyante.predict` never takes: the C ABI called with PINNED input and output buffers (results are copied
    device -> caller directly, single-chunk and pipelined paths), and cvb_predict_device_x on device buffers of raw uint8 /
    int16 counts and fp16 values -- all bit-identical to the ordinary call"""
    import ctypes
    import torch
    from clairvoyante_b200 import _lib, utils_v2 as U
    W = I.init_weights(variant, 3)
    m = _model(variant, W)
    lib = _lib.load()
    for n in (777, 40000):
        x = synth.make_sites(n, 21)
        cnt = U.pack_counts(x)
        ref_o, ref_l = m.predictLogits(x)
        xin = torch.from_numpy(cnt).pin_memory()
        outs = [torch.empty((n, k), dtype=torch.float32).pin_memory() for k in (4, 2, 4, 6, 16)]
        _lib.check(lib.cvb_predict_host_counts_u8(m._h, xin.data_ptr(), n, *[o.data_ptr() for o in outs]))
        got = torch.cat(outs[:4], 1).numpy()
        assert np.array_equal(got, ref_o) and np.array_equal(outs[4].numpy(), ref_l)
        st = torch.cuda.current_stream().cuda_stream
        for kind, arr in (("u8", cnt), ("i16", cnt.astype(np.int16)), ("f16", x.astype(np.float16)), ("f32", x)):
            xd = torch.from_numpy(arr).cuda()
            od = torch.empty((n, 16), dtype=torch.float32, device="cuda")
            ld = torch.empty((n, 16), dtype=torch.float32, device="cuda")
            torch.cuda.synchronize()
            m.predictDeviceX(xd.data_ptr(), kind, n, od.data_ptr(), ld.data_ptr(), st)
            torch.cuda.synchronize()
            assert np.array_equal(od.cpu().numpy(), ref_o) and np.array_equal(ld.cpu().numpy(), ref_l), kind
    # error behaviour of the new entry points
    assert lib.cvb_predict_device_x(m._h, 16, 7, 1, 16, None, None) != 0          # unknown element kind
    assert lib.cvb_allreduce_attach(m._h, None) == 0 and lib.cvb_allreduce_gradients(m._h) == 0   # no communicator: no-ops
    if variant == "v3":
        with pytest.raises(RuntimeError):
            m.setComputeMode("fp16")                                                # plain fp16 is a v3_slim mode
    m.close()


@pytest.mark.parametrize("variant,mode", CASES)
def test_submit_collect_pipeline_matches_predict(variant, mode):
    """cvb_predict_submit / cvb_predict_collect: several small batches in flight give predict's bits, for every feed, in any
    collection order; the slot limit, stale tickets and the synchronous entry points' refusal while tickets are out"""
    from clairvoyante_b200 import utils_v2 as U
    W = I.init_weights(variant, 3)
    m = _model(variant, W, mode)
    x = synth.make_sites(2400, 9)
    cuts = [(0, 1000), (1000, 1001), (1001, 1334), (1334, 1334), (1334, 2334)]
    feeds = [x[a:b] for a, b in cuts]
    feeds[0] = U.with_counts(feeds[0])                       # CountBatch -> uint8 counts
    feeds[2] = U.pack_counts(feeds[2]).astype(np.int16)      # explicit int16 counts
    feeds[4] = feeds[4].astype(np.float16)                   # fp16 values
    want = [m.predict(x[a:b]) for a, b in cuts]
    tickets = [m.predictSubmit(f) for f in feeds[:4]]
    with pytest.raises(RuntimeError, match="in flight"):
        m.predictSubmit(feeds[4])                            # four slots
    with pytest.raises(RuntimeError, match="not collected"):
        m.predict(x[:10])
    with pytest.raises(RuntimeError, match="in flight"):
        m._set("fc4/bias", 0, W["fc4/bias"])
    got = {}
    for i in (2, 0, 3, 1):                                   # any order
        got[i] = m.predictCollect(tickets[i])
    with pytest.raises(RuntimeError, match="not in flight"):
        m.predictCollect(tickets[0])
    t4 = m.predictSubmit(feeds[4], want_logits=True)
    got[4] = m.predictCollect(t4)
    for i, w in enumerate(want):
        for a, b in zip(w, got[i][:4]):
            assert a.shape == b.shape and np.array_equal(a, b), (i, a.shape)
    assert np.array_equal(got[4][4], m.predictLogits(x[1334:2334])[1])
    with pytest.raises(RuntimeError, match="exceed one device pass"):
        m.predictSubmit(np.zeros((40000, 33, 4, 4), np.uint8))
    # the generator form, consumed to the end and abandoned half-way
    outs = list(m.predictStream([x[i:i + 300] for i in range(0, 2400, 300)], depth=3))
    assert np.array_equal(np.concatenate([o[2] for o in outs]), m.predict(x)[2])
    g = m.predictStream([x[i:i + 300] for i in range(0, 2400, 300)], depth=4)
    next(g)
    g.close()
    assert np.array_equal(m.predict(x[:300])[0], outs[0][0])  # nothing left outstanding
    m.close()

"""CPU: Blosc-1 / LZ4 frame decoder (csrc/blosc_frame.cpp) and the reference-format .bin loader
(train.py:40-44, utils_v2.py:174-176,198).  Frames and Python-2 pickles are produced by the test-side encoder
tests/blosc_writer.py (python-blosc is absent: parity unpinned against frames written by the real library)."""
import ctypes
import os
import pickle
import struct
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import blosc_writer as BW   # noqa: E402

from clairvoyante_b200 import _lib, param, synth, utils_v2 as U   # noqa: E402


def _decompress(frame):
    lib = _lib.load()
    n, c, ts, fl = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.cvb_blosc_info(frame, len(frame), ctypes.byref(n), ctypes.byref(c), ctypes.byref(ts), ctypes.byref(fl)))
    assert c.value == len(frame)
    out = ctypes.create_string_buffer(max(1, n.value))
    got = ctypes.c_int64()
    _lib.check(lib.cvb_blosc_decompress(frame, len(frame), out, n.value, ctypes.byref(got)))
    return out.raw[:got.value]


def test_lz4_hand_vectors():
    """LZ4 block format by hand: literals + match, overlapping match (run), length extension bytes"""
    def frame(stream, nbytes):        # one block, one stream, no shuffle, typesize 1
        return struct.pack("<BBBBIII", 2, 1, 1 << 5, 1, nbytes, nbytes, 16 + 4 + 4 + len(stream)) + struct.pack("<i", 20) + \
            struct.pack("<i", len(stream)) + stream
    s = bytes([0x44]) + b"abcd" + bytes([4, 0]) + bytes([0x50]) + b"vwxyz"           # "abcd" + copy 8 from -4 + "vwxyz"
    assert _decompress(frame(s, 17)) == b"abcdabcdabcdvwxyz"
    s = bytes([0x1F]) + b"a" + bytes([1, 0]) + bytes([255, 10]) + bytes([0x50]) + b"12345"   # run: 4+15+255+10 = 284 x 'a'
    assert _decompress(frame(s, 1 + 284 + 5)) == b"a" * 285 + b"12345"
    lit = bytes(range(256)) * 2
    s = bytes([0xF0, 255, 512 - 15 - 255]) + lit                                     # 512 literals, no match
    assert _decompress(frame(s, 512)) == lit
    for bad in (bytes([0x44]) + b"abcd" + bytes([9, 0]) + bytes([0x50]) + b"vwxyz",  # offset before the start
                bytes([0x44]) + b"abcd" + bytes([0, 0]) + bytes([0x50]) + b"vwxyz",  # offset 0
                bytes([0x44]) + b"ab"):                                              # truncated literals
        lib = _lib.load()
        f = frame(bad, 17)
        out = ctypes.create_string_buffer(32)
        assert lib.cvb_blosc_decompress(f, len(f), out, 32, None) != 0
        assert b"corrupt" in lib.cvb_last_error()


@pytest.mark.parametrize("typesize,blocksize,kw", [
    (4, 32768, {}),                                  # split into 4 streams per block, trailing partial block unsplit
    (4, 1 << 20, {}),                                # single partial block
    (8, 16384, {}),                                  # labels are float64 (utils_v2.py:175)
    (4, 32768, dict(dont_split=True)),               # flag 0x10
    (4, 32768, dict(do_shuffle=False)),
    (40, 8192, {}),                                  # position strings: typesize > 16 -> never split, still shuffled
    (4, 300, {}),                                    # blocksize / typesize < 128 -> not split
    (4, 32768, dict(memcpyed=True)),
])
def test_frame_layouts_roundtrip(typesize, blocksize, kw):
    rng = np.random.RandomState(typesize + blocksize)
    data = (rng.poisson(2.0, 70001) * (rng.rand(70001) < 0.4)).astype(np.uint8).tobytes() + rng.bytes(3000)
    f = BW.blosc_compress(data, typesize, blocksize, **kw)
    assert _decompress(f) == data


def test_incompressible_stream_is_stored_raw():
    data = np.random.RandomState(1).bytes(4 * 128 * 64)
    f = BW.blosc_compress(data, 4, 4 * 128 * 16)
    assert _decompress(f) == data and len(f) <= len(data) + 16 + 4 * 4 + 4 * 16 + 64


def test_rejects_other_codecs_and_truncation():
    lib = _lib.load()
    data = b"x" * 4096
    f = bytearray(BW.blosc_compress(data, 1, 4096))
    out = ctypes.create_string_buffer(4096)
    g = bytes(f[:2]) + bytes([(f[2] & 0x1f) | (0 << 5)]) + bytes(f[3:])               # blosclz codec id
    assert lib.cvb_blosc_decompress(g, len(g), out, 4096, None) != 0 and b"LZ4" in lib.cvb_last_error()
    assert lib.cvb_blosc_decompress(bytes(f), len(f) - 3, out, 4096, None) != 0       # cbytes > buffer
    assert lib.cvb_blosc_decompress(bytes(f), len(f), out, 100, None) != 0            # destination too small
    assert lib.cvb_blosc_decompress(b"short", 5, out, 4096, None) != 0


def test_unpack_array_python2_pickles():
    x = synth.make_sites(500, 3)
    y = synth.make_labels(500, 3).astype(np.float64)
    pos = np.array(["chr%d:%d" % (i % 22 + 1, 10000 + i) for i in range(500)], dtype="S")
    for a, ts in ((x, 4), (y, 8), (pos, pos.dtype.itemsize)):
        frame = BW.blosc_compress(BW.py2_pickle_ndarray(a), ts, 65536)
        b = U.unpack_array(frame)
        assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a, b)
    # this repo's own container still loads, garbage is refused
    assert np.array_equal(U.unpack_array(U.pack_array(x)), x)
    with pytest.raises(ValueError):
        U.unpack_array(b"\x02\x01" + b"\x00" * 30)


def test_load_reference_style_bin_and_decompress(tmp_path):
    """a .bin laid out like the reference's tensor2Bin.py output (Python-2 pickles of lists of blosc frames) feeds
    DecompressArray exactly like this repo's own .bin"""
    n, bs = 1234, param.bloscBlockSize
    x = synth.make_sites(n, 5)
    y = synth.make_labels(n, 5).astype(np.float64)
    pos = np.array(["chr1:%d" % (5000 + i) for i in range(n)], dtype="S")
    xb, yb, pb = [], [], []
    for i in range(0, n, bs):
        xb.append(BW.blosc_compress(BW.py2_pickle_ndarray(x[i:i + bs]), 4, 131072))
        yb.append(BW.blosc_compress(BW.py2_pickle_ndarray(y[i:i + bs]), 8, 131072))
        pb.append(BW.blosc_compress(BW.py2_pickle_ndarray(pos[i:i + bs]), pos.dtype.itemsize, 131072))
    fn = str(tmp_path / "ref.bin")
    with open(fn, "wb") as fh:
        fh.write(BW.py2_pickle_int(n) + BW.py2_pickle_list_of_str(xb) + BW.py2_pickle_list_of_str(yb) + BW.py2_pickle_list_of_str(pb))
    total, X, Y, P = U.load_bin(fn)
    assert total == n and len(X) == len(xb) == 3
    got, num, end = U.DecompressArray(X, 400, 700, total)
    assert (num, end) == (700, 0) and np.array_equal(got, x[400:1100])
    got, num, end = U.DecompressArray(Y, 1000, 500, total)
    assert (num, end) == (234, 1) and np.array_equal(got, y[1000:])
    # own writer -> same loader
    fn2 = str(tmp_path / "own.bin")
    with open(fn2, "wb") as fh:
        for obj in (n, [U.pack_array(x[i:i + bs]) for i in range(0, n, bs)], [U.pack_array(y[i:i + bs]) for i in range(0, n, bs)],
                    [U.pack_array(pos[i:i + bs]) for i in range(0, n, bs)]):
            pickle.dump(obj, fh)
    t2, X2, _, _ = U.load_bin(fn2)
    assert t2 == n and np.array_equal(U.DecompressArray(X2, 0, n, n)[0], x)


def test_bin_loader_refuses_arbitrary_globals(tmp_path):
    fn = str(tmp_path / "evil.bin")
    with open(fn, "wb") as fh:
        fh.write(b"\x80\x02cos\nsystem\nU\x04trueq\x00\x85R.")
    with pytest.raises(pickle.UnpicklingError):
        U.load_bin(fn)


# ---- the encoder (csrc/blosc_frame.cpp cvb_blosc_compress, utils_v2.pack_array_blosc) ---------------------------------
def _native_compress(data, typesize, shuffle=1):
    import ctypes
    from clairvoyante_b200 import _lib
    lib = _lib.load()
    cap = int(lib.cvb_blosc_compress_bound(len(data)))
    out = ctypes.create_string_buffer(cap)
    got = ctypes.c_int64()
    _lib.check(lib.cvb_blosc_compress(data, len(data), typesize, shuffle, out, cap, ctypes.byref(got)))
    return out.raw[:got.value]


def _native_decompress(frame):
    import ctypes
    from clairvoyante_b200 import _lib
    lib = _lib.load()
    n = ctypes.c_int64()
    _lib.check(lib.cvb_blosc_info(frame, len(frame), ctypes.byref(n), None, None, None))
    out = ctypes.create_string_buffer(max(1, n.value))
    got = ctypes.c_int64()
    _lib.check(lib.cvb_blosc_decompress(frame, len(frame), out, n.value, ctypes.byref(got)))
    return out.raw[:got.value]


def test_native_encoder_round_trips_and_matches_the_frame_rules():
    rng = np.random.default_rng(0)
    cases = [b"", b"a", b"abcd" * 3, bytes(1000), rng.integers(0, 4, 5000, dtype=np.uint8).tobytes(),
             rng.integers(0, 256, 70000, dtype=np.uint8).tobytes(),                          # incompressible -> stored
             np.arange(200000, dtype=np.float32).tobytes(), (b"ACGT" * 100000) + b"xyz",   # > 1 block, leftover tail
             np.repeat(rng.standard_normal(3000), 40).astype(np.float64).tobytes(),
             bytes(range(256)) * 250 + b"tail!", np.eye(16)[np.arange(500) % 16].tobytes() + b"tb."]   # sizes that are not whole items
    for data in cases:
        for typesize in (1, 4, 8):
            for shuffle in (0, 1):
                frame = _native_compress(data, typesize, shuffle)
                assert int.from_bytes(frame[12:16], "little") == len(frame) and frame[0] == 2 and frame[3] == typesize
                assert int.from_bytes(frame[4:8], "little") == len(data)
                assert len(frame) <= len(data) + 16                                        # never worse than stored
                assert _native_decompress(frame) == data
    big = np.zeros((500, 33, 4, 4), np.float32).tobytes()
    assert len(_native_compress(big, 4)) < len(big) // 50


def test_pack_array_blosc_is_what_the_reference_writes(tmp_path):
    """blocks written with pack_array_blosc are Blosc-1 frames around Python-2 pickles: they read back here, their payload
    is byte-identical to the test-side Python-2 pickle writer, and a .bin written with --blosc loads through load_bin"""
    import pickle
    from clairvoyante_b200 import utils_v2 as U
    import blosc_writer as BW
    x = (np.arange(500 * 528) % 37).astype(np.float32).reshape(500, 33, 4, 4)
    y = np.eye(16, dtype=np.float64)[np.arange(500) % 16]
    pos = np.array([b"chr1:%d" % i for i in range(500)], dtype="S")
    for a in (x, y, pos, x[:0], pos[:0]):
        frame = U.pack_array_blosc(a)
        assert frame[:5] != b"CVBZ1" and frame[0] == 2
        assert _native_decompress(frame) == BW.py2_pickle_ndarray(a)
        b = U.unpack_array(frame)
        assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(b, a)
    fn = str(tmp_path / "d.bin")
    with open(fn, "wb") as fh:
        for p in (500, [U.pack_array_blosc(x)], [U.pack_array_blosc(y)], [U.pack_array_blosc(pos)]):
            pickle.dump(p, fh, protocol=2)
    total, xb, yb, pb = U.load_bin(fn)
    assert total == 500 and np.array_equal(U.unpack_array(xb[0]), x) and np.array_equal(U.unpack_array(pb[0]), pos)
    X, n, end = U.DecompressArray(xb, 0, 500, 500)
    assert n == 500 and end == 1 and np.array_equal(X, x)


# ---- the LZ4 layer against the real liblz4 (through pyarrow's "lz4_raw" codec = LZ4_compress_default / LZ4_decompress_safe) ----
def _liblz4():
    pa = pytest.importorskip("pyarrow")
    if not pa.Codec.is_available("lz4_raw"):
        pytest.skip("pyarrow built without lz4")
    return pa.Codec("lz4_raw")


def _streams(frame):
    """(uncompressed size, payload) of every split stream of a codec-1 frame, by the container rules of blosc_frame.cpp"""
    flags, typesize = frame[2], frame[3]
    nbytes, blocksize, _ = struct.unpack("<III", frame[4:16])
    if flags & 0x02 or nbytes == 0:
        return
    nblocks = (nbytes + blocksize - 1) // blocksize
    for b in range(nblocks):
        bsize = nbytes % blocksize if (b == nblocks - 1 and nbytes % blocksize) else blocksize
        leftover = bsize != blocksize
        nsplits = typesize if (not (flags & 0x10) and typesize <= 16 and blocksize // typesize >= 128 and not leftover) else 1
        off = struct.unpack("<i", frame[16 + 4 * b:20 + 4 * b])[0]
        for _ in range(nsplits):
            c = struct.unpack("<i", frame[off:off + 4])[0]
            yield b, bsize // nsplits, frame[off + 4:off + 4 + c]
            off += 4 + c


def _payloads():
    rng = np.random.default_rng(7)
    x = synth.make_sites(600, 3)
    return [x.tobytes(), BW.py2_pickle_ndarray(x[:500]), np.repeat(rng.standard_normal(2000), 37).astype(np.float64).tobytes(),
            (b"ACGT" * 70000) + b"xyz", rng.integers(0, 3, 300000, dtype=np.uint8).tobytes(), bytes(100000),
            np.arange(50000, dtype=np.int32).tobytes()]


def test_decoder_reads_streams_written_by_the_real_liblz4(monkeypatch):
    """frames assembled by the test-side container writer around streams compressed by liblz4 itself -- what python-blosc
    (codec lz4 / lz4hc: both emit the LZ4 block format) puts into the reference's .bin"""
    codec = _liblz4()
    monkeypatch.setattr(BW, "lz4_compress", lambda part: codec.compress(bytes(part), asbytes=True) if len(part) else b"\x00")
    for data in _payloads():
        for typesize, shuffle in ((4, True), (4, False), (8, True), (1, False)):
            frame = BW.blosc_compress(data, typesize, 256 * 1024 - (256 * 1024) % typesize, do_shuffle=shuffle)
            assert _decompress(frame) == data


def test_real_liblz4_decodes_the_streams_of_the_native_encoder():
    """every LZ4 stream cvb_blosc_compress writes is a valid LZ4 block for LZ4_decompress_safe and yields the (shuffled) bytes
    the container says it holds -- i.e. python-blosc in the reference's interpreter can read a `tensor2Bin --blosc` file"""
    codec = _liblz4()
    checked = 0
    for data in _payloads():
        for typesize, shuffle in ((4, 1), (4, 0), (8, 1), (1, 0)):
            frame = _native_compress(data, typesize, shuffle)
            blocks = {}
            for b, size, payload in _streams(frame):
                raw = payload if len(payload) == size else codec.decompress(payload, decompressed_size=size, asbytes=True)
                assert len(raw) == size
                blocks[b] = blocks.get(b, b"") + raw
                checked += len(payload) != size
            if blocks:
                out = b""
                for b in sorted(blocks):
                    blk = blocks[b]
                    if frame[2] & 0x01 and typesize > 1:       # undo the byte shuffle of this block
                        ne = len(blk) // typesize
                        blk = np.frombuffer(blk[:ne * typesize], np.uint8).reshape(typesize, ne).T.copy().tobytes() + blk[ne * typesize:]
                    out += blk
                assert out == data
    assert checked > 50

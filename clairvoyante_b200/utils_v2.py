"""Batch feed: text tensor stream -> batches, truth/labels -> training blocks, block slicing.

Python-3 counterpart of reference clairvoyante/utils_v2.py with the same public names and the same
yield / return contracts:

  GetTensor(tensor_fn, num)                       utils_v2.py:23-59   -> yields (endFlag, count, X, pos)
  GetTrainingArray(tensor_fn, var_fn, bed_fn)     utils_v2.py:62-186  -> (total, Xblocks, Yblocks, posBlocks)
  DecompressArray(blocks, start, num, maximum)    utils_v2.py:189-207 -> (array, num, endFlag)

Differences that are deliberate and documented:
  * python-blosc and intervaltree (requirements.txt:3-4) are not available; blocks are Blosc-1 / LZ4 frames
    written and read by this library's own codec (csrc/blosc_frame.cpp): `pack_array` (unshuffled),
    `pack_array_blosc` (byte-shuffled, the reference's layout); `unpack_array` also decodes the python-blosc
    LZ4HC frames of a .bin written by the reference and this repo's first zlib container.  BED membership uses
    sorted arrays + bisect with the same half-open [begin, end-1) semantics the reference builds (utils_v2.py:71-74).
  * a malformed row is reported and SKIPPED; the reference prints the failure and then re-uses the previous
    row's fields (utils_v2.py:34-41), silently duplicating a record.
"""
import bisect
import ctypes
import gc
import io
import os
import pickle
import random
import shlex
import subprocess
import sys
import zlib

import numpy as np

from . import _lib, param

base2num = dict(zip("ACGT", (0, 1, 2, 3)))
_ACGT = frozenset("ACGT")
_MAGIC = b"CVBZ1"


def SetupEnv():
    """utils_v2.py:14-18 (blosc threads / TF log level have no counterpart here)"""
    gc.enable()


def _site_floats():
    return (2 * param.flankingBaseNum + 1) * 4 * param.matrixNum


def _open_text(fn):
    """`gzip -fdc fn` like the reference (utils_v2.py:24-28): transparently reads plain or gzip text; 'PIPE' = stdin"""
    if fn == "PIPE":
        return None, sys.stdin
    proc = subprocess.Popen(shlex.split("gzip -fdc %s" % fn), stdout=subprocess.PIPE, bufsize=8388608)
    return proc, io.TextIOWrapper(proc.stdout, encoding="ascii", errors="replace")


def open_maybe_gzip(fn, mode="rt"):
    """plain or gzip file by its two magic bytes, whatever its name: what the reference's `gzip -fdc` readers accept
    (its own stages write gzip under suffix-less names, e.g. PrepDataBeforeDemo.sh's can_chr21_sampled)"""
    import gzip
    with open(fn, "rb") as probe:
        packed = probe.read(2) == b"\x1f\x8b"
    return gzip.open(fn, mode) if packed else open(fn, mode)


def _close_text(proc, fh):
    if proc is not None:
        fh.close()
        proc.wait()


def UnpackATensorRecord(a, b, c, *d):
    """utils_v2.py:20-21"""
    return a, b, c, np.array(d, dtype=np.float32)


def _subtract_reference_channel(x):
    """utils_v2.py:46 -- channels 1..3 hold counts relative to the reference channel 0"""
    x[..., 1:] -= x[..., 0:1]
    return x


def _open_bytes(fn):
    """binary twin of _open_text for the native parser"""
    if fn == "PIPE":
        return None, sys.stdin.buffer
    try:                                   # `gzip -fdc` passes a file that is not gzip through unchanged: read it directly
        with open(fn, "rb") as probe:
            plain = probe.read(2) != b"\x1f\x8b"
        if plain and os.path.isfile(fn):
            return None, open(fn, "rb", buffering=0)
    except OSError:
        pass                               # (let gzip report the problem like the reference would)
    proc = subprocess.Popen(shlex.split("gzip -fdc %s" % fn), stdout=subprocess.PIPE, bufsize=1 << 20)
    try:                                   # 1 MB pipe instead of 64 KB: fewer wake-ups per 32 MB read (Linux only)
        import fcntl
        fcntl.fcntl(proc.stdout.fileno(), getattr(fcntl, "F_SETPIPE_SZ", 1031), 1 << 20)
    except (ImportError, OSError):
        pass
    return proc, proc.stdout


_READ_BYTES = 32 << 20        # ~12k rows of ~2.7 KB per read


def _read_ahead(fb, size):
    """chunks of `fb` read on a helper thread, one ahead of the consumer (the read releases the GIL, so the next 32 MB arrive
    while the current ones are parsed); yields b"" for ever once the stream has ended"""
    import queue
    import threading
    q = queue.Queue(maxsize=2)

    def pump():
        try:
            while True:
                b = fb.read(size)
                q.put(b)
                if not b:
                    return
        except Exception as e:            # surfaced in the consumer
            q.put(e)

    threading.Thread(target=pump, daemon=True).start()
    while True:
        b = q.get()
        if isinstance(b, Exception):
            raise b
        if not b:
            break
        yield b
    while True:
        yield b""


class CountBatch(np.ndarray):
    """A float32 batch exactly as GetTensor yields it (channel 0 subtracted, utils_v2.py:46) that also carries `.counts`:
    the same sites as RAW uint8 / int16 counts for the narrow host->device feed (cvb_predict_host_counts_*, a quarter / a
    half of the bytes; the first device kernel redoes the subtraction).  The attribute belongs to this very object --
    slices and views do not inherit it -- so code that treats the batch as a plain ndarray is unaffected."""
    counts = None

    def __array_finalize__(self, obj):
        self.counts = None


def pack_counts(x, subtracted=True, threads=0):
    """raw counts behind the float32 tensors `x` ((n,33,4,4) or (n,528)) as a uint8 array (every count <= 255) or an int16
    one, same shape; None when `x` is not a tensor of non-negative integer counts (then only the fp32 feed is exact).
    subtracted=True: x is what GetTensor yields; False: x is CreateTensor's raw alnCode."""
    lib = _lib.load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    i16 = np.empty(x.shape, np.int16)
    u8 = np.empty(x.shape, np.uint8)
    mx, ok = ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.cvb_pack_counts(x.ctypes.data, x.size // 4, 1 if subtracted else 0, threads, i16.ctypes.data, u8.ctypes.data,
                                   ctypes.byref(mx), ctypes.byref(ok)))
    if not ok.value:
        return None
    return u8 if mx.value <= 255 else i16


def with_counts(x, subtracted=True, raw=None):
    """`x` (float32, subtracted) as a CountBatch whose `.counts` holds the narrow copy (None if the values are not counts)"""
    out = np.asarray(x).view(CountBatch)
    out.counts = pack_counts(raw if raw is not None else x, subtracted=subtracted and raw is None)
    return out


_PARSE_LINES = 16384          # lines handed to the parser per call, whatever the batch size: enough work for every host thread


def _native_rows(tensor_fn, num, threads=0):
    """Drives cvb_parse_tensor_text (csrc/text_feed.cpp) over the stream.  Yields (rows, recs) with rows a fresh
    float32 (k, 528) array (k <= num, channel 0 already subtracted, utils_v2.py:46) and recs the k
    "chrom:pos:SEQ" strings of those rows (cvb_tensor_text_positions).  Every yield but the last has k == num.
    The parser is called on up to _PARSE_LINES lines at a time and the batches are cut from its output, so that a
    1000-row batch (param.predictBatchSize) does not limit it to 1000 lines of work per call."""
    lib = _lib.load()
    proc, fb = _open_bytes(tensor_fn)
    width = _site_floats()
    if width != 528:
        raise NotImplementedError("the native parser is compiled for (33,4,4) tensors")
    n_lines, n_kept, n_used = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    buf, off, eof = b"", 0, False
    chunks = _read_ahead(fb, _READ_BYTES)
    stage = np.empty((_PARSE_LINES, width), dtype=np.float32)
    meta = np.empty((_PARSE_LINES, 10), dtype=np.int64)
    ptxt = np.empty(1, np.uint8)
    want = num if num > 0 else _PARSE_LINES
    rows = np.empty((want, width), dtype=np.float32)
    recs, c = [], 0
    narrow = os.environ.get("CVB_FEED", "counts") != "fp32"          # also hand out the batch as raw uint8 / int16 counts
    cstage = np.empty((_PARSE_LINES, width), np.int16) if narrow else None
    crows = np.empty((want, width), np.int16) if narrow else None
    cmax, mx, ok = 0, ctypes.c_int(), ctypes.c_int()
    while True:
        base = ctypes.cast(ctypes.c_char_p(buf), ctypes.c_void_p).value or 0
        _lib.check(lib.cvb_parse_tensor_text(base + off, len(buf) - off, 1 if eof else 0, _PARSE_LINES, threads,
                                             stage.ctypes.data, meta.ctypes.data,
                                             ctypes.byref(n_lines), ctypes.byref(n_kept), ctypes.byref(n_used)))
        nl, k = n_lines.value, n_kept.value
        if nl:
            for i in np.nonzero(meta[:nl, 0] == 2)[0]:   # CVB_LINE_MALFORMED (reference message, utils_v2.py:35)
                a, l = int(meta[i, 1]) + off, int(meta[i, 2])
                print("UnpackATensorRecord Failure", buf[a:a + l].decode("ascii", "replace"), file=sys.stderr)
            if k:                                        # "chrom:pos:SEQ" of the kept rows, assembled natively
                if len(ptxt) < n_used.value + 16:
                    ptxt = np.empty(n_used.value + 16, np.uint8)
                pn = lib.cvb_tensor_text_positions(base + off, meta.ctypes.data, nl, ptxt.ctypes.data, len(ptxt))
                if pn < 0:
                    _lib.check(1)
                names = ptxt[:pn].tobytes().decode("ascii", "replace").split("\n")
                if narrow:                               # packed once per parse, on every host thread
                    _lib.check(lib.cvb_pack_counts(stage.ctypes.data, k * (width // 4), 1, threads, cstage.ctypes.data, None,
                                                   ctypes.byref(mx), ctypes.byref(ok)))
                    if not ok.value:
                        narrow = False                   # not a stream of integer counts: fp32 feed from here on
                    cmax = max(cmax, mx.value)
                s0 = 0
                while s0 < k:                            # cut the batches out of this parse
                    if num <= 0 and c == len(rows):      # (num <= 0: one batch with everything)
                        rows = np.concatenate([rows, np.empty_like(rows)])
                    take = min(len(rows) - c, k - s0)
                    rows[c:c + take] = stage[s0:s0 + take]
                    if narrow:
                        if num <= 0 and len(crows) < len(rows):
                            crows = np.concatenate([crows, np.empty((len(rows) - len(crows), width), np.int16)])
                        crows[c:c + take] = cstage[s0:s0 + take]
                    recs += names[s0:s0 + take]
                    c += take
                    s0 += take
                    if num > 0 and c == num:
                        yield _attach_counts(rows, crows if narrow else None, cmax), recs
                        rows = np.empty((num, width), dtype=np.float32)   # fresh storage: the consumer still holds the batch
                        if narrow:
                            crows = np.empty((num, width), np.int16)
                        recs, c, cmax = [], 0, 0
            off += n_used.value
        if nl == 0 or off >= len(buf):
            if eof:
                break
            chunk = next(chunks)
            if chunk:
                buf, off = buf[off:] + chunk, 0
            else:
                eof = True                               # one more pass: a last line without '\n'
    if fb is not sys.stdin.buffer:
        fb.close()
    if proc is not None:
        proc.wait()
    yield _attach_counts(rows[:c], crows[:c] if narrow else None, cmax), recs


def _attach_counts(rows, counts, cmax):
    if counts is None:
        return rows
    out = rows.view(CountBatch)
    out.counts = counts.astype(np.uint8) if cmax <= 255 else counts
    return out


def _as_sites(rows, h):
    """(k,528) rows -> (k,33,4,4), keeping the narrow copy a CountBatch carries"""
    x = rows.reshape(-1, h, 4, param.matrixNum)
    c = getattr(rows, "counts", None)
    if c is not None:
        x.counts = c.reshape(x.shape)
    return x


def GetTensor(tensor_fn, num, threads=0):
    """Generator over batches of `num` candidate sites parsed from `chrom pos refseq33 v0..v527` rows
    (format: dataPrepScripts/CreateTensor.py:56).  Yields (0, num, X, pos) for full batches and finally
    (1, c, X[:c], pos) with 0 <= c < num (possibly empty), X float32 (c,33,4,4), pos = 'chrom:pos:seq'.
    Tokenising, number conversion and the channel subtract run in the C library (all host threads by default)."""
    h = 2 * param.flankingBaseNum + 1
    total = 0
    it = _native_rows(tensor_fn, num, threads)
    prev = next(it)
    for cur in it:
        rows, recs = prev
        total += len(recs)
        print("Processed %d tensors" % total, file=sys.stderr)
        yield 0, len(recs), _as_sites(rows, h), recs
        prev = cur
    rows, recs = prev
    total += len(recs)
    print("Processed %d tensors" % total, file=sys.stderr)
    yield 1, len(recs), _as_sites(rows, h), recs


# ------------------------------------------------------------------------------------------------
# block container (stands in for blosc.pack_array / unpack_array, utils_v2.py:174-176,198)
# ------------------------------------------------------------------------------------------------
def pack_array(a):
    """one block of a training set.  A Blosc-1 frame with LZ4 streams around the Python-2-style pickle of the array -- the
    structure of `blosc.pack_array` (utils_v2.py:174-176), readable by the reference -- but WITHOUT the byte shuffle: on
    these sparse count tensors the shuffled planes compress worse (163 KB vs 120 KB per 500-site block) and decode slower
    (0.77 vs 1.24 GB/s per thread), and the decode rate is what feeds the training step.  pack_array_blosc applies the
    shuffle like python-blosc's default."""
    return _blosc_pack(a, 0)


def pack_array_zlib(a):
    """the first container of this repo (zlib around .npy); still readable (unpack_array), no longer written by default"""
    buf = io.BytesIO()
    np.save(buf, np.asarray(a), allow_pickle=False)
    return _MAGIC + zlib.compress(buf.getvalue(), 1)


def _py2_pickle_ndarray(a):
    """protocol-2 pickle of an ndarray exactly as Python 2 + NumPy 1.x wrote it (the raw buffer is a `str`, the globals
    live in numpy.core.multiarray): what `blosc.pack_array` compresses (utils_v2.py:174-176), so that frames written
    here load in the reference's interpreter.  A Python-3 / NumPy-2 pickle would not (numpy._core, _codecs.encode)."""
    import struct
    a = np.ascontiguousarray(a)

    def sstr(t):
        t = t.encode("latin1") if isinstance(t, str) else t
        return (b"U" + bytes([len(t)]) + t) if len(t) < 256 else (b"T" + struct.pack("<i", len(t)) + t)

    def pint(v):
        if 0 <= v < 256:
            return b"K" + bytes([v])
        if 0 <= v < 65536:
            return b"M" + struct.pack("<H", v)
        return b"J" + struct.pack("<i", v)
    dt = a.dtype
    if dt.kind not in "fiuSb" or dt.byteorder == ">":
        raise ValueError("only little-endian numeric and fixed-width byte-string arrays are stored in a tensor file")
    code = dt.str[1:] if dt.kind != "S" else "S%d" % dt.itemsize
    order = "|" if dt.kind == "S" or dt.itemsize == 1 else "<"
    p = b"\x80\x02cnumpy.core.multiarray\n_reconstruct\ncnumpy\nndarray\nK\x00\x85" + sstr("b") + b"\x87R"
    p += b"(K\x01"                                                  # state tuple: version
    p += b"(" + b"".join(pint(d) for d in a.shape) + b"t"            # shape
    p += b"cnumpy\ndtype\n" + sstr(code) + b"K\x00K\x01\x87R"       # dtype(code, 0, 1)
    size_align = (pint(dt.itemsize) + pint(1)) if dt.kind == "S" else b"J\xff\xff\xff\xffJ\xff\xff\xff\xff"
    p += b"(K\x03" + sstr(order) + b"NNN" + size_align + b"K\x00tb"
    p += b"\x89"                                                    # is_fortran = False
    raw = a.tobytes()
    p += b"T" + struct.pack("<i", len(raw)) + raw                    # BINSTRING
    p += b"tb."
    return p


def _blosc_pack(a, shuffle):
    lib = _lib.load()
    a = np.asarray(a)
    payload = _py2_pickle_ndarray(a)
    cap = int(lib.cvb_blosc_compress_bound(len(payload)))
    out = np.empty(cap, np.uint8)
    got = ctypes.c_int64()
    _lib.check(lib.cvb_blosc_compress(payload, len(payload), min(int(a.dtype.itemsize), 255), shuffle, out.ctypes.data, cap,
                                      ctypes.byref(got)))
    return out[:got.value].tobytes()


def pack_array_blosc(a):
    """`blosc.pack_array(a, cname='lz4hc')` (utils_v2.py:174-176): Blosc-1 frame (typesize = itemsize, byte shuffle, LZ4
    streams -- csrc/blosc_frame.cpp) around the Python-2 pickle of the array.  Readable by the reference and by
    unpack_array here."""
    return _blosc_pack(a, 1)


def _fast_py2_ndarray(buf):
    """Zero-copy view of the array inside a protocol-2 ndarray pickle laid out as Python 2 + NumPy wrote it (and as
    _py2_pickle_ndarray writes it): walks the fixed opcode sequence -- memo PUTs skipped -- down to the BINSTRING that holds
    the raw buffer and returns np.frombuffer on it.  None when the bytes do not follow that layout (the caller then falls
    back to the restricted unpickler).  buf: uint8 ndarray."""
    b = buf
    n = len(b)
    pos = 0

    def skip_puts():
        nonlocal pos
        while pos < n:
            if b[pos] == 0x71:            # BINPUT
                pos += 2
            elif b[pos] == 0x72:          # LONG_BINPUT
                pos += 5
            else:
                break

    def expect(lit):
        nonlocal pos
        m = len(lit)
        if pos + m > n or bytes(b[pos:pos + m]) != lit:
            return False
        pos += m
        skip_puts()
        return True

    def read_int():
        nonlocal pos
        if pos >= n:
            return None
        op = b[pos]
        if op == 0x4b and pos + 2 <= n:                      # BININT1
            v = int(b[pos + 1]); pos += 2
        elif op == 0x4d and pos + 3 <= n:                    # BININT2
            v = int(b[pos + 1]) | (int(b[pos + 2]) << 8); pos += 3
        elif op == 0x4a and pos + 5 <= n:                    # BININT
            v = int.from_bytes(bytes(b[pos + 1:pos + 5]), "little", signed=True); pos += 5
        else:
            return None
        return v

    def read_short_str():
        nonlocal pos
        if pos + 2 > n or b[pos] != 0x55:                    # SHORT_BINSTRING
            return None
        m = int(b[pos + 1])
        v = bytes(b[pos + 2:pos + 2 + m])
        pos += 2 + m
        skip_puts()
        return v

    head = min(n, 512)
    if not (expect(b"\x80\x02") and expect(b"cnumpy.core.multiarray\n_reconstruct\n") and expect(b"cnumpy\nndarray\n")
            and expect(b"K\x00\x85") and read_short_str() == b"b" and expect(b"\x87R") and expect(b"(K\x01")):
        return None
    shape = []
    if pos < head and b[pos] == 0x28:                        # MARK ints TUPLE
        pos += 1
        while pos < head and b[pos] != 0x74:
            v = read_int()
            if v is None:
                return None
            shape.append(v)
        pos += 1
    else:                                                    # ints then TUPLE1/2/3, or EMPTY_TUPLE
        while pos < head and b[pos] in (0x4b, 0x4d, 0x4a):
            shape.append(read_int())
        if pos >= head or b[pos] not in (0x85, 0x86, 0x87, 0x29) or (b[pos] == 0x29) != (len(shape) == 0):
            return None
        pos += 1
    skip_puts()
    if not expect(b"cnumpy\ndtype\n"):
        return None
    code = read_short_str()
    if code is None or not (expect(b"K\x00K\x01\x87R") and expect(b"(K\x03")):
        return None
    order = read_short_str()
    if order not in (b"<", b"|", b"=") or not expect(b"NNN"):
        return None
    if read_int() is None or read_int() is None or not expect(b"K\x00t"):
        return None
    if not expect(b"b") or pos >= n or b[pos] != 0x89:       # BUILD, then is_fortran must be NEWFALSE
        return None
    pos += 1
    if pos < n and b[pos] == 0x54 and pos + 5 <= n:          # BINSTRING
        m = int.from_bytes(bytes(b[pos + 1:pos + 5]), "little")
        off = pos + 5
    elif pos < n and b[pos] == 0x55 and pos + 2 <= n:        # SHORT_BINSTRING
        m = int(b[pos + 1])
        off = pos + 2
    else:
        return None
    try:
        dt = np.dtype(code.decode("ascii"))
    except (TypeError, UnicodeDecodeError):
        return None
    count = 1
    for d in shape:
        count *= d
    if m != count * dt.itemsize or off + m > n:
        return None
    return np.frombuffer(b, dtype=dt, count=count, offset=off).reshape(shape)


class _RestrictedUnpickler(pickle.Unpickler):
    """The reference unpickles its .bin blindly (train.py:41-44, blosc.unpack_array); only the globals a pickled
    ndarray / list / int needs are resolved here."""
    _ALLOWED = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
                ("numpy", "ndarray"), ("numpy", "dtype"), ("_codecs", "encode")}

    def find_class(self, module, name):
        if (module, name) not in self._ALLOWED:
            raise pickle.UnpicklingError("refusing to load %s.%s from a tensor file" % (module, name))
        if name == "_reconstruct":
            return np._core.multiarray._reconstruct if hasattr(np, "_core") else np.core.multiarray._reconstruct
        if module == "_codecs":
            import codecs
            return codecs.encode
        return getattr(np, name)


def _unpickle(data, encoding):
    return _RestrictedUnpickler(io.BytesIO(data), encoding=encoding).load()


def _blosc_unpack(b):
    """blosc.unpack_array (utils_v2.py:198): Blosc-1 frame -> pickled ndarray (Python-2 pickles load with latin1).  The
    frame is decompressed into a NumPy buffer and, when the pickle has the plain ndarray layout, the result is a view of
    that buffer (no second copy of the megabyte); anything else goes through the restricted unpickler."""
    lib = _lib.load()
    nbytes = ctypes.c_int64()
    _lib.check(lib.cvb_blosc_info(b, len(b), ctypes.byref(nbytes), None, None, None))
    out = np.empty(max(1, nbytes.value), np.uint8)
    got = ctypes.c_int64()
    _lib.check(lib.cvb_blosc_decompress(b, len(b), out.ctypes.data, nbytes.value, ctypes.byref(got)))
    a = _fast_py2_ndarray(out[:got.value])
    if a is not None:
        return a
    return _unpickle(out[:got.value].tobytes(), "latin1")


def unpack_array(b):
    """one block of a .bin: this repo's container (pack_array) or a python-blosc frame written by the reference"""
    if isinstance(b, str):                          # a Python-2 `str` that went through a latin1 unpickle
        b = b.encode("latin1")
    if not isinstance(b, (bytes, bytearray)):
        raise ValueError("tensor block is not a byte string")
    b = bytes(b)
    if b[:5] == _MAGIC:
        return np.load(io.BytesIO(zlib.decompress(b[5:])), allow_pickle=False)
    if len(b) >= 16 and b[0] in (1, 2) and int.from_bytes(b[12:16], "little") == len(b):
        return _blosc_unpack(b)
    raise ValueError("not a tensor block: neither this repo's container nor a Blosc-1 frame")


def load_bin(fn):
    """train.py:40-44: four consecutive pickles (total, X blocks, Y blocks, position blocks); accepts files written by
    this repo's tensor2Bin.py and by the reference's (Python 2, python-blosc frames)"""
    with open(fn, "rb") as fh:
        data = fh.read()
    bio = io.BytesIO(data)
    out = [_RestrictedUnpickler(bio, encoding="bytes").load() for _ in range(4)]
    return int(out[0]), list(out[1]), list(out[2]), list(out[3])


class _Regions(object):
    """per-contig interval membership with intervaltree's half-open semantics"""

    def __init__(self):
        self._raw = {}
        self._idx = {}

    def add(self, name, begin, end):
        self._raw.setdefault(name, []).append((begin, end))
        self._idx.pop(name, None)

    def __contains__(self, name):
        return name in self._raw

    def hit(self, name, p):
        if name not in self._raw:
            raise KeyError(name)
        if name not in self._idx:
            iv = sorted(self._raw[name])
            begins = [b for b, _ in iv]
            run, m = [], -1
            for _, e in iv:
                m = max(m, e)
                run.append(m)
            self._idx[name] = (begins, run)
        begins, run = self._idx[name]
        i = bisect.bisect_right(begins, p)
        return i > 0 and run[i - 1] > p


def _truth_label(ref, alt, gt1, gt2):
    """16-vector [A C G T | HET HOM | REF SNP INS DEL | len 0 1 2 3 4 >4] from one truth row (utils_v2.py:84-119)"""
    v = [0.0] * 16
    single = len(ref) == 1 and len(alt) == 1
    if gt1 == "0" and gt2 == "1":
        v[base2num[ref[0]]] = 0.5
        if single:
            v[base2num[alt[0]]] = 0.5
        v[4] = 1.0
    elif gt1 == "1" and gt2 == "1":
        if single:
            v[base2num[alt[0]]] = 1
        v[5] = 1.0
    if len(ref) > 1 and len(alt) == 1:
        v[9] = 1.0
    elif len(alt) > 1 and len(ref) == 1:
        v[8] = 1.0
    else:
        v[7] = 1.0
    d = abs(len(ref) - len(alt))
    v[15 if d > 4 else 10 + d] = 1.0
    return v


def GetTrainingArray(tensor_fn, var_fn, bed_fn, shuffle=True, container="cvbz"):
    """utils_v2.py:62-186.  container: "cvbz" = this repo's block container (pack_array), "blosc" = python-blosc frames
    around Python-2 pickles like the reference writes (pack_array_blosc) -- unpack_array / DecompressArray read both."""
    pack = pack_array_blosc if container == "blosc" else pack_array
    regions = None
    if bed_fn is not None:
        regions = _Regions()
        proc, fh = _open_text(bed_fn)
        for row in fh:
            f = row.split()
            begin, end = int(f[1]), int(f[2]) - 1
            if end == begin:
                end += 1
            regions.add(f[0], begin, end)
        _close_text(proc, fh)

    Y = {}
    if var_fn is not None:
        proc, fh = _open_text(var_fn)
        for row in fh:
            f = row.split()
            if regions is not None and not regions.hit(f[0], int(f[1])):
                continue
            Y[f[0] + ":" + f[1]] = _truth_label(f[2], f[3], f[4], f[5])
        _close_text(proc, fh)

    X = {}
    h, centre = 2 * param.flankingBaseNum + 1, param.flankingBaseNum
    total = 0
    for rows, recs in _native_rows(tensor_fn, 4096):
        for r, rec in zip(rows.reshape(-1, h, 4, param.matrixNum), recs):
            chrom, coord, seq = rec.rsplit(":", 2)
            if regions is not None and (chrom not in regions or not regions.hit(chrom, int(coord))):
                continue
            key = chrom + ":" + coord
            X[key] = r
            if key not in Y:                      # non-variant default label (utils_v2.py:141-148)
                v = [0.0] * 16
                v[base2num[seq[centre]]] = 1.0
                v[5] = v[6] = v[10] = 1.0
                Y[key] = v
            total += 1
            if total % 100000 == 0:
                print("Processed %d tensors" % total, file=sys.stderr)

    keys = sorted(X.keys())
    if shuffle:
        random.shuffle(keys)
    xb, yb, pb = [], [], []
    step = param.bloscBlockSize
    for i in range(0, len(keys), step):
        chunk = keys[i:i + step]
        xb.append(pack(np.array([X[k] for k in chunk], dtype=np.float32)))
        yb.append(pack(np.array([Y[k] for k in chunk], dtype=np.float64)))   # labels are float64 upstream (:175)
        pb.append(pack(np.array(chunk, dtype="S")))
    if len(keys) % step == 0:                 # the reference always appends a (possibly empty) trailing block (:181-184)
        xb.append(pack(np.zeros((0, h, 4, param.matrixNum), np.float32)))
        yb.append(pack(np.zeros((0, 16), np.float64)))
        pb.append(pack(np.array([], dtype="S1")))
    return len(keys), xb, yb, pb


_POOL = None


def _decode_pool():
    global _POOL
    if _POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 2) - 1)))
    return _POOL


def DecompressArray(array, start, num, maximum):
    """rows [start, start+num) across bloscBlockSize-row blocks; clamps at `maximum` (utils_v2.py:189-207)"""
    endFlag = 0
    if start + num >= maximum:
        num = maximum - start
        endFlag = 1
    bs = param.bloscBlockSize
    if num <= 0:
        # start == maximum (a set whose size is a multiple of the block size, asked for its tail): the reference decodes the
        # always-present trailing block and returns an empty slice of it (utils_v2.py:196-207)
        blk = unpack_array(array[min(start // bs, len(array) - 1)]) if len(array) else np.empty((0,), np.float32)
        return blk[:0], 0, 1
    first, last = start // bs, (start + num - 1) // bs
    head = unpack_array(array[first]) if last - first >= 2 else None
    if head is not None and head.dtype.kind not in "SUO":
        # (byte-string blocks -- the position keys -- differ in width from block to block: np.concatenate below widens them)
        # a training batch spans 20 blocks: decode them on the pool (the C decoder and zlib drop the GIL) and let every worker
        # copy its rows straight into the batch -- decode and the 21 MB gather both run in parallel
        out = np.empty((num,) + head.shape[1:], head.dtype)

        def place(i, a=None):
            a = unpack_array(array[i]) if a is None else a
            g0 = i * bs
            lo, hi = max(start, g0), min(start + num, g0 + len(a))
            if hi > lo:
                out[lo - start:hi - start] = a[lo - g0:hi - g0]

        futures = [_decode_pool().submit(place, i) for i in range(first + 1, last + 1)]
        place(first, head)
        for f in futures:
            f.result()
        return out, num, endFlag
    parts = [head if (i == first and head is not None) else unpack_array(array[i]) for i in range(first, last + 1)]
    out = np.concatenate(parts) if len(parts) > 1 else parts[0]
    left = start % bs
    if left != 0 or num % bs != 0:
        out = out[left:left + num]
    return out, num, endFlag

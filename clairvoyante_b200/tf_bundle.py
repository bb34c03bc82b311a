"""TensorFlow-free reader / writer of TensorFlow "V2" checkpoint bundles (`<prefix>.index` +
`<prefix>.data-00000-of-00001`), the format `tf.train.Saver().save / .restore` uses in the reference
(clairvoyante/clairvoyante_v3.py:243-251; published models: README "trainedModels").  SURVEY.md 8(f) rank 2.

Format (tensorflow/core/util/tensor_bundle + tensorflow/core/lib/io/table, restated from the public sources;
TensorFlow itself is not installed here, so this module is checked by round trips, by the CRC-32C test vectors and
against the protobuf schemas TensorBoard ships -- "parity unpinned" against a real TF-written file):

  .index   an SSTable (LevelDB table format, TF's trimmed copy):
             data block(s) | meta-index block | index block | 48-byte footer
           block   = entries, restart array (uint32 LE each), uint32 number of restarts
           entry   = varint32 shared, varint32 non_shared, varint32 value_len, key suffix, value
           every block is followed by a 5-byte trailer: compression type (0 none, 1 snappy) and the masked
           CRC-32C of block + type byte
           footer  = BlockHandle(meta-index) BlockHandle(index) zero-padded to 40 bytes, then the magic
                     0xdb4775248b80fb57 little-endian;  BlockHandle = varint64 offset, varint64 size
           key ""  -> BundleHeaderProto {1: num_shards, 2: endianness (0 little), 3: VersionDef {1: producer}}
           key name-> BundleEntryProto  {1: dtype, 2: TensorShapeProto {2: Dim {1: size}}, 3: shard_id, 4: offset,
                                         5: size, 6: fixed32 masked crc32c of the tensor bytes, 7: slices}
  .data-SSSSS-of-NNNNN   the raw little-endian tensor bytes back to back at the recorded offsets.
"""
import os
import struct

import numpy as np

from . import _lib

MAGIC = 0xdb4775248b80fb57
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64 = 1, 2, 3, 9
_DTYPES = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8")}
_DTYPE_IDS = {v: k for k, v in _DTYPES.items()}
_MASK_DELTA = 0xa282ead8
BLOCK_SIZE = 262144          # table::Options defaults
RESTART_INTERVAL = 16


# ---- checksums -------------------------------------------------------------------------------------------
def crc32c(data, crc=0):
    b = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    return int(_lib.load().cvb_crc32c(crc, bytes(b), len(b)))


def mask_crc(c):
    return ((((c >> 15) | (c << 17)) & 0xffffffff) + _MASK_DELTA) & 0xffffffff


def unmask_crc(m):
    r = (m - _MASK_DELTA) & 0xffffffff
    return ((r >> 17) | (r << 15)) & 0xffffffff


# ---- varints / minimal protobuf ---------------------------------------------------------------------------
def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _get_varint(b, i):
    v, s = 0, 0
    while True:
        c = b[i]
        i += 1
        v |= (c & 0x7f) << s
        if c < 0x80:
            return v, i
        s += 7
        if s > 63:
            raise ValueError("varint too long")


def _pb_fields(b):
    """yield (field_number, wire_type, value) of one serialized message"""
    i, n = 0, len(b)
    while i < n:
        key, i = _get_varint(b, i)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _get_varint(b, i)
        elif wt == 1:
            v = struct.unpack_from("<Q", b, i)[0]
            i += 8
        elif wt == 2:
            ln, i = _get_varint(b, i)
            v = bytes(b[i:i + ln])
            i += ln
        elif wt == 5:
            v = struct.unpack_from("<I", b, i)[0]
            i += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield f, wt, v


def _pb_varint(f, v):
    return _put_varint(f << 3) + _put_varint(v)


def _pb_bytes(f, b):
    return _put_varint((f << 3) | 2) + _put_varint(len(b)) + b


def encode_shape(shape):
    return b"".join(_pb_bytes(2, _pb_varint(1, int(d)) if int(d) else b"") for d in shape)    # proto3 omits a zero size


def decode_shape(b):
    dims = []
    for f, wt, v in _pb_fields(b):
        if f == 2 and wt == 2:
            size = 0
            for g, _, w in _pb_fields(v):
                if g == 1:
                    size = w if w < (1 << 63) else w - (1 << 64)
            dims.append(size)
        elif f == 3 and v:
            raise ValueError("tensor of unknown rank in checkpoint")
    return tuple(dims)


def encode_header(num_shards=1, producer=1):
    return _pb_varint(1, num_shards) + _pb_bytes(3, _pb_varint(1, producer))       # endianness LITTLE = 0 is the default


def encode_entry(dtype_id, shape, shard_id, offset, size, crc_masked):
    b = _pb_varint(1, dtype_id) + _pb_bytes(2, encode_shape(shape))
    if shard_id:
        b += _pb_varint(3, shard_id)
    if offset:
        b += _pb_varint(4, offset)
    if size:
        b += _pb_varint(5, size)
    b += _put_varint((6 << 3) | 5) + struct.pack("<I", crc_masked)
    return b


def decode_entry(b):
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for f, wt, v in _pb_fields(b):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            e["shape"] = decode_shape(v)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["sliced"] = True
    return e


# ---- snappy (index blocks are written uncompressed by TF's BundleWriter; kept for tables that are not) ------
def snappy_uncompress(b):
    n, i = _get_varint(b, 0)
    out = bytearray()
    while i < len(b):
        tag = b[i]
        i += 1
        t = tag & 3
        if t == 0:
            ln = tag >> 2
            if ln >= 60:
                k = ln - 59
                ln = int.from_bytes(b[i:i + k], "little")
                i += k
            ln += 1
            out += b[i:i + ln]
            i += ln
            continue
        if t == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | b[i]
            i += 1
        elif t == 2:
            ln = (tag >> 2) + 1
            off = b[i] | (b[i + 1] << 8)
            i += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(b[i:i + 4], "little")
            i += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy block")
        for _ in range(ln):                       # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


# ---- table reader ------------------------------------------------------------------------------------------
def _read_block(buf, offset, size, verify):
    body = buf[offset:offset + size]
    if len(body) != size or offset + size + 5 > len(buf):
        raise ValueError("index table truncated")
    ctype = buf[offset + size]
    stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
    if verify and unmask_crc(stored) != crc32c(buf[offset:offset + size + 1]):
        raise ValueError("index table block checksum mismatch")
    if ctype == 1:
        body = snappy_uncompress(body)
    elif ctype != 0:
        raise ValueError("unknown block compression %d" % ctype)
    return body


def _block_entries(block):
    if len(block) < 4:
        raise ValueError("bad table block")
    nrestarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestarts
    if end < 0:
        raise ValueError("bad table block")
    i, key = 0, b""
    while i < end:
        shared, i = _get_varint(block, i)
        non_shared, i = _get_varint(block, i)
        vlen, i = _get_varint(block, i)
        key = key[:shared] + bytes(block[i:i + non_shared])
        i += non_shared
        yield key, bytes(block[i:i + vlen])
        i += vlen


def read_table(path, verify=True):
    """[(key bytes, value bytes)] of a TF/LevelDB table file in key order"""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad magic)" % path)
    foot = buf[len(buf) - 48:]
    _, i = _get_varint(foot, 0)            # meta-index handle (unused)
    _, i = _get_varint(foot, i)
    ioff, i = _get_varint(foot, i)
    isize, i = _get_varint(foot, i)
    out = []
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, j = _get_varint(handle, 0)
        size, j = _get_varint(handle, j)
        out.extend(_block_entries(_read_block(buf, off, size, verify)))
    return out


# ---- table writer ------------------------------------------------------------------------------------------
class _BlockBuilder(object):
    def __init__(self, interval=RESTART_INTERVAL):
        self.buf, self.restarts, self.count, self.last, self.interval = bytearray(), [0], 0, b"", interval

    def add(self, key, value):
        shared = 0
        if self.count % self.interval == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last = key
        self.count += 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _shortest_separator(a, b):
    """a <= result < b, as short as possible (BytewiseComparator::FindShortestSeparator)"""
    m = min(len(a), len(b))
    d = 0
    while d < m and a[d] == b[d]:
        d += 1
    if d < m and a[d] < 0xff and a[d] + 1 < b[d]:
        return a[:d] + bytes([a[d] + 1])
    return a


def _short_successor(a):
    for i, c in enumerate(a):
        if c != 0xff:
            return a[:i] + bytes([c + 1])
    return a


def write_table(path, items):
    """items: iterable of (key bytes, value bytes) in strictly increasing key order"""
    out = bytearray()

    def emit(block):
        off = len(out)
        out.extend(block)
        out.append(0)                                                        # kNoCompression
        out.extend(struct.pack("<I", mask_crc(crc32c(bytes(block) + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    index, cur, pending = _BlockBuilder(1), _BlockBuilder(), None      # index blocks restart at every entry
    last = None
    for key, value in items:
        if last is not None and key <= last:
            raise ValueError("table keys must be strictly increasing")
        if pending is not None:
            index.add(_shortest_separator(pending[0], key), pending[1])
            pending = None
        cur.add(key, value)
        last = key
        if cur.size() >= BLOCK_SIZE:
            pending = (key, emit(cur.finish()))
            cur = _BlockBuilder()
    if cur.count:
        pending = (last, emit(cur.finish()))
    if pending is not None:
        index.add(_short_successor(pending[0]), pending[1])
    meta = emit(_BlockBuilder().finish())
    idx = emit(index.finish())
    foot = meta + idx
    out.extend(foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", MAGIC))
    with open(path, "wb") as fh:
        fh.write(out)


# ---- bundles -------------------------------------------------------------------------------------------------
def _data_path(prefix, shard, num_shards):
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def is_bundle(prefix):
    return os.path.isfile(prefix + ".index")


def list_bundle(prefix, verify=True):
    """{name: entry dict} plus the header under key ''"""
    items = read_table(prefix + ".index", verify)
    if not items or items[0][0] != b"":
        raise ValueError("%s.index has no bundle header" % prefix)
    hdr = dict(num_shards=1, endianness=0)
    for f, _, v in _pb_fields(items[0][1]):
        if f == 1:
            hdr["num_shards"] = v
        elif f == 2:
            hdr["endianness"] = v
    if hdr["endianness"] != 0:
        raise ValueError("big-endian checkpoint bundles are not supported")
    out = {"": hdr}
    for k, v in items[1:]:
        out[k.decode("utf-8")] = decode_entry(v)
    return out


def read_bundle(prefix, names=None, verify=True):
    """{variable name: ndarray}.  `names` restricts the load; unknown dtypes are skipped unless asked for by name."""
    entries = list_bundle(prefix, verify)
    hdr = entries.pop("")
    files, out = {}, {}
    try:
        for name, e in entries.items():
            if names is not None and name not in names:
                continue
            if e["sliced"]:
                raise ValueError("variable %s is stored as slices (partitioned variable): not supported" % name)
            dt = _DTYPES.get(e["dtype"])
            if dt is None:
                if names is not None:
                    raise ValueError("variable %s has unsupported dtype id %d" % (name, e["dtype"]))
                continue
            fh = files.get(e["shard_id"])
            if fh is None:
                fh = files[e["shard_id"]] = open(_data_path(prefix, e["shard_id"], hdr["num_shards"]), "rb")
            fh.seek(e["offset"])
            raw = fh.read(e["size"])
            count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
            if len(raw) != e["size"] or count * dt.itemsize != e["size"]:
                raise ValueError("variable %s: size %d does not match shape %r" % (name, e["size"], e["shape"]))
            if verify and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(raw):
                raise ValueError("variable %s: data checksum mismatch" % name)
            out[name] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    finally:
        for fh in files.values():
            fh.close()
    return out


def write_bundle(prefix, tensors):
    """tensors: {name: ndarray (float32 / float64 / int32 / int64)}; one shard, names in byte order like tf.train.Saver"""
    d = os.path.dirname(os.path.abspath(prefix))
    os.makedirs(d, exist_ok=True)
    items = [(b"", encode_header(1, 1))]
    off = 0
    with open(_data_path(prefix, 0, 1), "wb") as fh:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            a = np.asarray(tensors[name])
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            if np.dtype(dt) not in _DTYPE_IDS:
                raise ValueError("variable %s: dtype %s not supported" % (name, a.dtype))
            raw = np.ascontiguousarray(a, dtype=dt).tobytes()
            fh.write(raw)
            items.append((name.encode("utf-8"),
                          encode_entry(_DTYPE_IDS[np.dtype(dt)], a.shape, 0, off, len(raw), mask_crc(crc32c(raw)))))
            off += len(raw)
    write_table(prefix + ".index", items)

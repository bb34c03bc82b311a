"""CPU: the TensorFlow-free checkpoint bundle reader/writer (clairvoyante_b200/tf_bundle.py), the format behind
tf.train.Saver in the reference (clairvoyante_v3.py:243-251).  TensorFlow is absent, so the pins are: the CRC-32C
check value and TensorBoard's independent implementation, the protobuf schemas TensorBoard ships for the shared
sub-messages, a hand-assembled LevelDB block, and writer <-> reader round trips."""
import os
import struct

import numpy as np
import pytest

from clairvoyante_b200 import initializers as I, tf_bundle as B


def test_crc32c_check_value_and_tensorboard():
    assert B.crc32c(b"123456789") == 0xE3069283                       # RFC 3720 check value
    assert B.crc32c(b"") == 0 and B.crc32c(b"\x00" * 32) == 0x8A9136AA
    tb = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    rng = np.random.RandomState(0)
    for n in (1, 7, 8, 9, 63, 64, 4097):
        b = rng.bytes(n)
        assert B.crc32c(b) == tb.crc32c(b) & 0xffffffff
        assert B.mask_crc(B.crc32c(b)) == tb.masked_crc32c(b)
        assert B.unmask_crc(B.mask_crc(B.crc32c(b))) == B.crc32c(b)
    assert B.crc32c(b"6789", B.crc32c(b"12345")) == 0xE3069283        # incremental


def test_sub_messages_match_tensorflow_schemas():
    pb = pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    types = pytest.importorskip("tensorboard.compat.proto.types_pb2")
    ver = pytest.importorskip("tensorboard.compat.proto.versions_pb2")
    assert (B.DT_FLOAT, B.DT_DOUBLE, B.DT_INT32, B.DT_INT64) == (types.DT_FLOAT, types.DT_DOUBLE, types.DT_INT32, types.DT_INT64)
    for shape in [(), (16,), (1, 4, 4, 16), (4608, 336), (0, 3)]:
        m = pb.TensorShapeProto()
        for d in shape:
            m.dim.add().size = d
        assert B.encode_shape(shape) == m.SerializeToString()
        assert B.decode_shape(m.SerializeToString()) == shape
    # header = {1: num_shards, 3: VersionDef}: the VersionDef bytes must parse with the real schema
    hdr = B.encode_header(1, 1)
    fields = {f: v for f, _, v in B._pb_fields(hdr)}
    assert fields[1] == 1
    v = ver.VersionDef()
    v.ParseFromString(fields[3])
    assert v.producer == 1 and v.min_consumer == 0


def test_entry_roundtrip_and_wire_bytes():
    e = B.encode_entry(B.DT_FLOAT, (2, 3), 0, 24, 24, 0xDEADBEEF)
    assert e == bytes([0x08, 0x01, 0x12, 0x08, 0x12, 0x02, 0x08, 0x02, 0x12, 0x02, 0x08, 0x03, 0x20, 24, 0x28, 24,
                       0x35]) + struct.pack("<I", 0xDEADBEEF)
    d = B.decode_entry(e)
    assert (d["dtype"], d["shape"], d["shard_id"], d["offset"], d["size"], d["crc32c"]) == (1, (2, 3), 0, 24, 24, 0xDEADBEEF)
    assert B.decode_entry(B.encode_entry(B.DT_FLOAT, (), 0, 0, 4, 1))["shape"] == ()


def test_reads_hand_assembled_leveldb_table(tmp_path):
    """a table written byte by byte from the LevelDB format description (prefix-compressed keys, restart array,
    5-byte trailers, 48-byte footer) -- not by write_table"""
    def block(entries, restarts):
        b = b"".join(bytes([s, len(k), len(v)]) + k + v for s, k, v in entries)
        return b + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))

    def trailer(b):
        return b"\x00" + struct.pack("<I", B.mask_crc(B.crc32c(b + b"\x00")))
    data = block([(0, b"apple", b"1"), (2, b"ricot", b"22"), (0, b"banana", b"")], [0])   # "ap"+"ricot" = apricot
    meta = block([], [0])
    out = data + trailer(data)
    moff = len(out)
    out += meta + trailer(meta)
    index = block([(0, b"c", bytes([0, len(data)]))], [0])
    ioff = len(out)
    out += index + trailer(index)
    foot = bytes([moff, len(meta), ioff, len(index)])
    out += foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", B.MAGIC)
    p = str(tmp_path / "t.index")
    open(p, "wb").write(out)
    assert B.read_table(p) == [(b"apple", b"1"), (b"apricot", b"22"), (b"banana", b"")]
    bad = bytearray(out)
    bad[3] ^= 1
    open(p, "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        B.read_table(p)
    assert len(B.read_table(p, verify=False)) == 3


def test_snappy_decoder():
    # literal "abcd" + copy(offset 4, len 8) -> "abcdabcdabcd" (overlapping copy), then a 1-byte-offset copy
    raw = bytes([12, (3 << 2) | 0]) + b"abcd" + bytes([((8 - 1) << 2) | 2, 4, 0])
    assert B.snappy_uncompress(raw) == b"abcdabcdabcd"
    raw = bytes([9, (4 << 2) | 0]) + b"hello" + bytes([((4 - 4) << 2) | 1 | (0 << 5), 5])
    assert B.snappy_uncompress(raw) == b"hellohell"
    with pytest.raises(ValueError):
        B.snappy_uncompress(bytes([5, 0]) + b"a")


@pytest.mark.parametrize("variant", ["v3", "v3_slim"])
def test_bundle_roundtrip_saver_names(tmp_path, variant, monkeypatch):
    W = I.init_weights(variant, 3)
    t = {}
    for k, v in W.items():
        t[k] = v
        t[k + "/Adam"] = (0.1 * v).astype(np.float32)
        t[k + "/Adam_1"] = (v * v).astype(np.float32)
    t["beta1_power"] = np.float32(0.9 ** 4)
    t["beta2_power"] = np.float32(0.999 ** 4)
    t["global_step"] = np.int64(3)
    prefix = str(tmp_path / "ck" / "model-000003")
    B.write_bundle(prefix, t)
    assert os.path.getsize(prefix + ".data-00000-of-00001") == sum(np.asarray(v).nbytes for v in t.values())
    ent = B.list_bundle(prefix)
    assert ent[""]["num_shards"] == 1 and len(ent) == len(t) + 1
    names = [k for k in ent if k]
    assert names == sorted(names, key=lambda s: s.encode())            # Saver order = byte order; offsets follow it
    offs = [ent[k]["offset"] for k in names]
    assert offs == sorted(offs) and offs[0] == 0
    r = B.read_bundle(prefix)
    for k in t:
        assert r[k].dtype == np.asarray(t[k]).dtype and np.array_equal(r[k], np.asarray(t[k])), k
    assert set(B.read_bundle(prefix, names={"fc4/kernel"})) == {"fc4/kernel"}
    # many small data blocks + separators in the index block
    monkeypatch.setattr(B, "BLOCK_SIZE", 300)
    B.write_bundle(prefix + "b", t)
    r = B.read_bundle(prefix + "b")
    assert all(np.array_equal(r[k], np.asarray(t[k])) for k in t)
    # corrupt one data byte -> per-tensor checksum catches it
    dp = prefix + ".data-00000-of-00001"
    raw = bytearray(open(dp, "rb").read())
    raw[100] ^= 0x40
    open(dp, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        B.read_bundle(prefix)


def test_not_a_bundle(tmp_path):
    p = str(tmp_path / "x")
    open(p + ".index", "wb").write(b"hello world, definitely not a table of at least 48 bytes........")
    with pytest.raises(ValueError, match="magic"):
        B.read_bundle(p)
    assert not B.is_bundle(str(tmp_path / "missing"))

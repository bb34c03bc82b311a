"""Test-side ENCODER of Blosc-1 / LZ4 frames and of Python-2 style ndarray pickles, written from the public format
descriptions, used to exercise the decoder in clairvoyante_b200/csrc/blosc_frame.cpp (python-blosc is not installed, so
the frames the reference would write cannot be produced here: "parity unpinned")."""
import struct

import numpy as np


def lz4_compress(src):
    """greedy LZ4 block compressor (hash of 4 bytes, last 5 bytes literal, no match starting in the last 12 bytes)"""
    n = len(src)
    out = bytearray()
    table = {}
    i = anchor = 0

    def emit(lit, mlen, off):
        tok_l = min(len(lit), 15)
        tok_m = 0 if mlen is None else min(mlen - 4, 15)
        out.append((tok_l << 4) | tok_m)
        if len(lit) >= 15:
            r = len(lit) - 15
            while r >= 255:
                out.append(255)
                r -= 255
            out.append(r)
        out.extend(lit)
        if mlen is not None:
            out.extend(struct.pack("<H", off))
            if mlen - 4 >= 15:
                r = mlen - 4 - 15
                while r >= 255:
                    out.append(255)
                    r -= 255
                out.append(r)

    while i + 12 < n:
        key = src[i:i + 4]
        cand = table.get(key)
        table[key] = i
        if cand is not None and i - cand <= 65535:
            m = 4
            while i + m < n - 5 and src[cand + m] == src[i + m]:
                m += 1
            emit(src[anchor:i], m, i - cand)
            i += m
            anchor = i
        else:
            i += 1
    emit(src[anchor:], None, 0)
    return bytes(out)


def shuffle(b, typesize):
    ne = len(b) // typesize
    a = np.frombuffer(b[:ne * typesize], np.uint8).reshape(ne, typesize).T.copy().tobytes()
    return a + b[ne * typesize:]


def blosc_compress(data, typesize, blocksize, do_shuffle=True, dont_split=False, store_raw_if_bigger=True, memcpyed=False):
    """frame with codec id 1 (LZ4): header, bstarts, per-block split streams"""
    nbytes = len(data)
    flags = (1 << 5) | (0x01 if do_shuffle else 0) | (0x10 if dont_split else 0) | (0x02 if memcpyed else 0)
    if memcpyed:
        body = bytes(data)
        return struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, 16 + nbytes) + body
    nblocks = (nbytes + blocksize - 1) // blocksize
    blocks = []
    for b in range(nblocks):
        blk = bytes(data[b * blocksize:(b + 1) * blocksize])
        leftover = len(blk) != blocksize
        if do_shuffle and typesize > 1:
            blk = shuffle(blk, typesize)
        nsplits = typesize if (not dont_split and typesize <= 16 and blocksize // typesize >= 128 and not leftover) else 1
        ne = len(blk) // nsplits
        enc = bytearray()
        for s in range(nsplits):
            part = blk[s * ne:(s + 1) * ne]
            c = lz4_compress(part)
            if store_raw_if_bigger and len(c) >= len(part):
                c = part                       # cbytes == neblock marks a raw stream
            elif len(c) == len(part):
                c = part
            enc += struct.pack("<i", len(c)) + c
        blocks.append(bytes(enc))
    off = 16 + 4 * nblocks
    bstarts = []
    for e in blocks:
        bstarts.append(off)
        off += len(e)
    hdr = struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, off)
    return hdr + b"".join(struct.pack("<i", x) for x in bstarts) + b"".join(blocks)


def py2_pickle_ndarray(a):
    """protocol-2 pickle of an ndarray as Python 2 + NumPy 1.x wrote it: the raw buffer is a `str` (BINSTRING)"""
    a = np.ascontiguousarray(a)

    def sstr(s):
        s = s.encode("latin1") if isinstance(s, str) else s
        return (b"U" + bytes([len(s)]) + s) if len(s) < 256 else (b"T" + struct.pack("<i", len(s)) + s)

    def pint(v):
        if 0 <= v < 256:
            return b"K" + bytes([v])
        if 0 <= v < 65536:
            return b"M" + struct.pack("<H", v)
        return b"J" + struct.pack("<i", v)
    dt = a.dtype
    code = dt.str[1:] if dt.kind != "S" else "S%d" % dt.itemsize
    order = "|" if dt.kind == "S" or dt.itemsize == 1 else "<"
    p = b"\x80\x02cnumpy.core.multiarray\n_reconstruct\ncnumpy\nndarray\nK\x00\x85" + sstr("b") + b"\x87R"
    p += b"(K\x01"                                                   # state tuple: version
    p += b"(" + b"".join(pint(d) for d in a.shape) + b"t"            # shape
    p += b"cnumpy\ndtype\n" + sstr(code) + b"K\x00K\x01\x87R"        # dtype(code, 0, 1)
    size_align = (pint(dt.itemsize) + pint(1)) if dt.kind == "S" else b"J\xff\xff\xff\xffJ\xff\xff\xff\xff"   # flexible dtypes carry them
    p += b"(K\x03" + sstr(order) + b"NNN" + size_align + b"K\x00tb"
    p += b"\x89"                                                     # is_fortran = False
    raw = a.tobytes()
    p += b"T" + struct.pack("<i", len(raw)) + raw                    # BINSTRING
    p += b"tb."
    return p


def py2_pickle_list_of_str(items):
    """protocol-2 pickle of a Python-2 list of `str`"""
    p = b"\x80\x02]("
    for s in items:
        p += (b"U" + bytes([len(s)]) + s) if len(s) < 256 else (b"T" + struct.pack("<i", len(s)) + s)
    return p + b"e."


def py2_pickle_int(v):
    return b"\x80\x02J" + struct.pack("<i", v) + b"."

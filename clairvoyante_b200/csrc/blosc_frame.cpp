// blosc_frame.cpp -- decoder and encoder for Blosc-1 frames holding LZ4 / LZ4HC blocks (host code).
//
// The reference stores its training set as python-blosc frames, `blosc.pack_array(array, cname='lz4hc')`
// (clairvoyante/utils_v2.py:174-176,182-184) and reads them back with `blosc.unpack_array` (:198,202).
// python-blosc is not installed here; this restates the public Blosc-1 container and LZ4 block formats
//  so that a .bin written by the reference can feed clairvoyante_b200/train.py:
//
//   header (16 bytes): version, versionlz, flags, typesize, nbytes u32, blocksize u32, cbytes u32  (little-endian)
//     flags: 0x01 byte-shuffle, 0x02 stored uncompressed ("memcpyed"), 0x04 bit-shuffle, 0x10 blocks are not split,
//            bits 5..7 codec format (0 blosclz, 1 lz4 / lz4hc, 2 snappy, 3 zlib, 4 zstd)
//   bstarts: i32[nblocks], offset of every block from the start of the frame
//   block  : nsplits streams, each `i32 cbytes` + payload; a stream whose cbytes equals its uncompressed size is stored
//            raw.  nsplits = typesize when typesize <= 16, blocksize / typesize >= 128, the block is not the trailing
//            partial one and flag 0x10 is clear; otherwise 1.
//   After decompression a shuffled block holds byte 0 of all its elements, then byte 1, ...; the bytes beyond the last
//   whole element are stored in place.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/cvb200.h"

void cvb_internal_set_error(const char* msg);  // cvb200.cu

namespace {

int fail(const char* msg) { cvb_internal_set_error(msg); return 1; }

inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// LZ4 block format: sequences of (token, literal-length bytes, literals, 2-byte offset, match-length bytes); the last
// sequence stops after its literals.  Bounds-checked; returns false on any malformed input.
bool lz4_block_decode(const uint8_t* s, int64_t slen, uint8_t* d, int64_t dlen) {
  const uint8_t* ip = s;
  const uint8_t* const iend = s + slen;
  uint8_t* op = d;
  uint8_t* const oend = d + dlen;
  while (ip < iend) {
    const unsigned token = *ip++;
    int64_t lit = token >> 4;
    if (lit == 15) {
      unsigned b;
      do {
        if (ip >= iend) return false;
        b = *ip++;
        lit += b;
      } while (b == 255);
    }
    if (lit > iend - ip || lit > oend - op) return false;
    if (lit <= 16 && iend - ip >= 16 && oend - op >= 16) memcpy(op, ip, 16);  // short literal runs: one fixed-size copy
    else memcpy(op, ip, (size_t)lit);
    ip += lit;
    op += lit;
    if (ip >= iend) break;
    if (iend - ip < 2) return false;
    const int64_t offset = (int64_t)ip[0] | ((int64_t)ip[1] << 8);
    ip += 2;
    if (offset == 0 || offset > op - d) return false;
    int64_t ml = token & 15;
    if (ml == 15) {
      unsigned b;
      do {
        if (ip >= iend) return false;
        b = *ip++;
        ml += b;
      } while (b == 255);
    }
    ml += 4;
    if (ml > oend - op) return false;
    const uint8_t* m = op - offset;
    if (offset >= 16 && oend - op >= ml + 16) {  // the common case: 16-byte strides never read what they have not written yet
      for (int64_t k = 0; k < ml; k += 16) memcpy(op + k, m + k, 16);
      op += ml;
    } else if (offset >= ml) {
      memcpy(op, m, (size_t)ml);
      op += ml;
    } else if (offset == 1) {  // run of one byte (what long stretches of zeros in a shuffled tensor block become)
      memset(op, *m, (size_t)ml);
      op += ml;
    } else {  // overlapping match = periodic run: copy the period, then keep doubling the copied span
      uint8_t* const start = op;
      memcpy(op, m, (size_t)offset);
      int64_t done = offset;
      while (done < ml) {
        const int64_t c = done < ml - done ? done : ml - done;
        memcpy(start + done, start, (size_t)c);
        done += c;
      }
      op += ml;
    }
  }
  return op == oend;
}

void unshuffle(const uint8_t* src, uint8_t* dst, int64_t n, int typesize) {
  const int64_t ne = n / typesize;
  if (typesize == 4) {  // float32 tensors: gather the four byte planes element by element (sequential writes)
    const uint8_t *s0 = src, *s1 = src + ne, *s2 = src + 2 * ne, *s3 = src + 3 * ne;
    uint32_t* o = reinterpret_cast<uint32_t*>(dst);
    if ((reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
      for (int64_t i = 0; i < ne; ++i) o[i] = (uint32_t)s0[i] | ((uint32_t)s1[i] << 8) | ((uint32_t)s2[i] << 16) | ((uint32_t)s3[i] << 24);
    } else {
      for (int64_t i = 0; i < ne; ++i) { dst[4 * i] = s0[i]; dst[4 * i + 1] = s1[i]; dst[4 * i + 2] = s2[i]; dst[4 * i + 3] = s3[i]; }
    }
  } else if (typesize == 8) {
    const uint8_t* sp[8];
    for (int b = 0; b < 8; ++b) sp[b] = src + (int64_t)b * ne;
    for (int64_t i = 0; i < ne; ++i) {
      uint64_t v = 0;
      for (int b = 0; b < 8; ++b) v |= (uint64_t)sp[b][i] << (8 * b);
      memcpy(dst + 8 * i, &v, 8);
    }
  } else {
    for (int b = 0; b < typesize; ++b) {
      const uint8_t* s = src + (int64_t)b * ne;
      uint8_t* o = dst + b;
      for (int64_t i = 0; i < ne; ++i, o += typesize) *o = s[i];
    }
  }
  const int64_t done = ne * typesize;
  memcpy(dst + done, src + done, (size_t)(n - done));
}

}  // namespace

extern "C" int cvb_blosc_info(const void* frame, int64_t n, int64_t* nbytes, int64_t* cbytes, int* typesize, int* flags) {
  const uint8_t* p = static_cast<const uint8_t*>(frame);
  if (!p || n < 16) return fail("cvb_blosc_info: frame shorter than the 16-byte Blosc header");
  if (p[0] < 1 || p[0] > 2) return fail("cvb_blosc_info: not a Blosc-1 frame (format version byte)");
  if (nbytes) *nbytes = rd32(p + 4);
  if (cbytes) *cbytes = rd32(p + 12);
  if (typesize) *typesize = p[3];
  if (flags) *flags = p[2];
  return 0;
}

extern "C" int cvb_blosc_decompress(const void* frame, int64_t n, void* dst, int64_t cap, int64_t* out_n) {
  const uint8_t* p = static_cast<const uint8_t*>(frame);
  if (out_n) *out_n = 0;
  if (!p || n < 16) return fail("cvb_blosc_decompress: frame shorter than the 16-byte Blosc header");
  if (p[0] < 1 || p[0] > 2) return fail("cvb_blosc_decompress: not a Blosc-1 frame (format version byte)");
  const int flags = p[2];
  int typesize = p[3];
  const int64_t nbytes = rd32(p + 4), blocksize = rd32(p + 8), cbytes = rd32(p + 12);
  if (cbytes > n) return fail("cvb_blosc_decompress: frame truncated (cbytes > buffer)");
  if (nbytes > cap || (nbytes > 0 && !dst)) return fail("cvb_blosc_decompress: destination too small");
  if (out_n) *out_n = nbytes;
  if (nbytes == 0) return 0;
  uint8_t* out = static_cast<uint8_t*>(dst);
  if (flags & 0x02) {  // memcpyed
    if (16 + nbytes > cbytes) return fail("cvb_blosc_decompress: stored frame truncated");
    memcpy(out, p + 16, (size_t)nbytes);
    return 0;
  }
  if (flags & 0x04) return fail("cvb_blosc_decompress: bit-shuffled frames are not supported");
  const int codec = (flags >> 5) & 7;
  if (codec != 1)
    return fail("cvb_blosc_decompress: only the LZ4 / LZ4HC codec is supported (the reference packs with cname='lz4hc', utils_v2.py:174)");
  if (blocksize <= 0 || typesize <= 0) return fail("cvb_blosc_decompress: bad header");
  const int64_t nblocks = (nbytes + blocksize - 1) / blocksize;
  if (16 + 4 * nblocks > cbytes) return fail("cvb_blosc_decompress: block table truncated");
  const bool shuffled = (flags & 0x01) && typesize > 1;
  const bool dont_split = (flags & 0x10) != 0;
  std::vector<uint8_t> tmp(shuffled ? (size_t)(blocksize < nbytes ? blocksize : nbytes) : 0);  // (a block never exceeds the payload)
  for (int64_t b = 0; b < nblocks; ++b) {
    const int64_t bsize = (b == nblocks - 1 && nbytes % blocksize) ? nbytes % blocksize : blocksize;
    const bool leftover = bsize != blocksize;
    const int nsplits = (!dont_split && typesize <= 16 && blocksize / typesize >= 128 && !leftover) ? typesize : 1;
    const int64_t neblock = bsize / nsplits;
    int64_t src = (int32_t)rd32(p + 16 + 4 * b);
    uint8_t* blk = shuffled ? tmp.data() : out + b * blocksize;
    for (int s = 0; s < nsplits; ++s) {
      if (src < 0 || src + 4 > cbytes) return fail("cvb_blosc_decompress: block offset outside the frame");
      const int64_t c = (int32_t)rd32(p + src);
      src += 4;
      if (c < 0 || src + c > cbytes) return fail("cvb_blosc_decompress: stream length outside the frame");
      if (c == neblock) memcpy(blk + s * neblock, p + src, (size_t)c);
      else if (!lz4_block_decode(p + src, c, blk + s * neblock, neblock)) return fail("cvb_blosc_decompress: corrupt LZ4 stream");
      src += c;
    }
    if (shuffled) unshuffle(tmp.data(), out + b * blocksize, bsize, typesize);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Encoder: what `blosc.pack_array(a, cname='lz4hc')` (utils_v2.py:174-176) hands to the .bin, so that a training set
// prepared here can be read by the reference's `blosc.unpack_array`.  Any valid LZ4 block stream decodes with LZ4 / LZ4HC
// decoders alike (HC only searches harder); this one is the plain greedy scheme: hash of 4 bytes -> last position, extend,
// the last 5 bytes are literals and no match starts within the last 12 bytes of a stream (format end conditions).
// ---------------------------------------------------------------------------------------------------------------------
namespace {

inline void wr32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

// returns the encoded size, or -1 when it would not fit in cap
int64_t lz4_block_encode(const uint8_t* s, int64_t n, uint8_t* d, int64_t cap) {
  uint8_t* op = d;
  uint8_t* const oend = d + cap;
  auto emit = [&](const uint8_t* lit, int64_t nlit, int64_t mlen, int64_t off) -> bool {
    const int64_t need = 1 + nlit / 255 + 1 + nlit + (mlen ? 2 + (mlen - 4) / 255 + 1 : 0);
    if (need > oend - op) return false;
    const int64_t tl = nlit < 15 ? nlit : 15, tm = mlen ? (mlen - 4 < 15 ? mlen - 4 : 15) : 0;
    *op++ = (uint8_t)((tl << 4) | tm);
    if (nlit >= 15) {
      int64_t r = nlit - 15;
      while (r >= 255) { *op++ = 255; r -= 255; }
      *op++ = (uint8_t)r;
    }
    memcpy(op, lit, (size_t)nlit);
    op += nlit;
    if (mlen) {
      *op++ = (uint8_t)off;
      *op++ = (uint8_t)(off >> 8);
      if (mlen - 4 >= 15) {
        int64_t r = mlen - 4 - 15;
        while (r >= 255) { *op++ = 255; r -= 255; }
        *op++ = (uint8_t)r;
      }
    }
    return true;
  };
  constexpr int HB = 14;
  std::vector<int32_t> table((size_t)1 << HB, -1);
  int64_t i = 0, anchor = 0;
  while (i + 12 < n) {
    const uint32_t key = rd32(s + i);
    const uint32_t h = (key * 2654435761u) >> (32 - HB);
    const int64_t cand = table[h];
    table[h] = (int32_t)i;
    if (cand >= 0 && i - cand <= 65535 && rd32(s + cand) == key) {
      int64_t m = 4;
      while (i + m < n - 5 && s[cand + m] == s[i + m]) ++m;
      if (!emit(s + anchor, i - anchor, m, i - cand)) return -1;
      i += m;
      anchor = i;
    } else {
      ++i;
    }
  }
  if (!emit(s + anchor, n - anchor, 0, 0)) return -1;
  return (int64_t)(op - d);
}

void shuffle_bytes(const uint8_t* src, uint8_t* dst, int64_t n, int typesize) {
  const int64_t ne = n / typesize;
  for (int b = 0; b < typesize; ++b) {
    const uint8_t* s = src + b;
    uint8_t* o = dst + (int64_t)b * ne;
    for (int64_t i = 0; i < ne; ++i, s += typesize) o[i] = *s;
  }
  const int64_t done = ne * typesize;
  memcpy(dst + done, src + done, (size_t)(n - done));
}

}  // namespace

extern "C" int64_t cvb_blosc_compress_bound(int64_t nbytes) { return nbytes + nbytes / 255 + 16 + 4 * (nbytes / 4096 + 2) + 16 * 64 + 64; }

extern "C" int cvb_blosc_compress(const void* src, int64_t nbytes, int typesize, int do_shuffle, void* dst, int64_t cap,
                                  int64_t* out_n) {
  const uint8_t* in = static_cast<const uint8_t*>(src);
  uint8_t* out = static_cast<uint8_t*>(dst);
  if (!out || !out_n || (!in && nbytes > 0) || nbytes < 0 || nbytes > 0x7fffffff - 16) return fail("cvb_blosc_compress: bad argument");
  if (typesize < 1 || typesize > 255) typesize = 1;  // (Blosc treats over-long items as bytes)
  if (cap < 16 + nbytes) return fail("cvb_blosc_compress: destination too small (see cvb_blosc_compress_bound)");
  const bool shuf = do_shuffle && typesize > 1;
  int64_t blocksize = 256 * 1024;
  if (blocksize > nbytes) blocksize = nbytes > 0 ? nbytes : 1;
  if (blocksize > typesize) blocksize -= blocksize % typesize;  // whole items per block (c-blosc's rule): the split streams divide
                                                                // evenly, the odd bytes form a trailing partial block
  const int64_t nblocks = nbytes > 0 ? (nbytes + blocksize - 1) / blocksize : 0;
  const uint8_t flags = (uint8_t)((1 << 5) | (shuf ? 0x01 : 0));
  out[0] = 2; out[1] = 1; out[2] = flags; out[3] = (uint8_t)typesize;
  wr32(out + 4, (uint32_t)nbytes);
  wr32(out + 8, (uint32_t)blocksize);
  int64_t off = 16 + 4 * nblocks;
  bool fits = off <= cap;
  std::vector<uint8_t> tmp((size_t)(shuf ? blocksize : 0));
  for (int64_t b = 0; b < nblocks && fits; ++b) {
    const int64_t bsize = (b == nblocks - 1 && nbytes % blocksize) ? nbytes % blocksize : blocksize;
    const bool leftover = bsize != blocksize;
    const uint8_t* blk = in + b * blocksize;
    if (shuf) { shuffle_bytes(blk, tmp.data(), bsize, typesize); blk = tmp.data(); }
    const int nsplits = (typesize <= 16 && blocksize / typesize >= 128 && !leftover) ? typesize : 1;
    const int64_t ne = bsize / nsplits;
    wr32(out + 16 + 4 * b, (uint32_t)off);
    for (int s2 = 0; s2 < nsplits && fits; ++s2) {
      if (off + 4 + ne > cap) { fits = false; break; }
      int64_t c = lz4_block_encode(blk + s2 * ne, ne, out + off + 4, ne - 1);  // must be strictly smaller than raw
      if (c < 0) { memcpy(out + off + 4, blk + s2 * ne, (size_t)ne); c = ne; }   // cbytes == size marks a raw stream
      wr32(out + off, (uint32_t)c);
      off += 4 + c;
    }
  }
  if (!fits || off >= 16 + nbytes) {  // did not shrink: store ("memcpyed", flag 0x02)
    out[2] = (uint8_t)(flags | 0x02);
    if (nbytes) memcpy(out + 16, in, (size_t)nbytes);
    off = 16 + nbytes;
  }
  wr32(out + 12, (uint32_t)off);
  *out_n = off;
  return 0;
}

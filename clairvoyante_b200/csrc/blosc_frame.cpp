// blosc_frame.cpp -- decoder for Blosc-1 frames holding LZ4 / LZ4HC blocks (host code).
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
    memcpy(op, ip, (size_t)lit);
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
    if (offset >= ml) {
      memcpy(op, m, (size_t)ml);
      op += ml;
    } else {
      for (int64_t i = 0; i < ml; ++i) *op++ = *m++;  // overlapping match = run
    }
  }
  return op == oend;
}

void unshuffle(const uint8_t* src, uint8_t* dst, int64_t n, int typesize) {
  const int64_t ne = n / typesize;
  for (int b = 0; b < typesize; ++b) {
    const uint8_t* s = src + (int64_t)b * ne;
    uint8_t* o = dst + b;
    for (int64_t i = 0; i < ne; ++i, o += typesize) *o = s[i];
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
  std::vector<uint8_t> tmp(shuffled ? (size_t)blocksize : 0);
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

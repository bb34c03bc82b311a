// crc32c.cpp -- CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), slicing-by-8, host code.
// TensorFlow's checkpoint bundle stores a masked CRC-32C per tensor and per index block
// (tf.train.Saver, clairvoyante/clairvoyante_v3.py:243-251 -> tensor_bundle V2); clairvoyante_b200/tf_bundle.py calls this.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../include/cvb200.h"

namespace {
struct Tables {
  uint32_t t[8][256];
  Tables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1) ? 0x82F63B78u : 0u);
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};
const Tables kT;
}  // namespace

extern "C" uint32_t cvb_crc32c(uint32_t crc, const void* data, int64_t n) {
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = ~crc;
  if (!p || n <= 0) return crc;
  while (n > 0 && (reinterpret_cast<uintptr_t>(p) & 7)) { c = (c >> 8) ^ kT.t[0][(c ^ *p++) & 0xff]; --n; }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;  // little-endian hosts only (x86-64 / aarch64-le), like the rest of the library
    c = kT.t[7][w & 0xff] ^ kT.t[6][(w >> 8) & 0xff] ^ kT.t[5][(w >> 16) & 0xff] ^ kT.t[4][(w >> 24) & 0xff] ^
        kT.t[3][(w >> 32) & 0xff] ^ kT.t[2][(w >> 40) & 0xff] ^ kT.t[1][(w >> 48) & 0xff] ^ kT.t[0][(w >> 56) & 0xff];
    p += 8;
    n -= 8;
  }
  while (n-- > 0) c = (c >> 8) ^ kT.t[0][(c ^ *p++) & 0xff];
  return ~c;
}

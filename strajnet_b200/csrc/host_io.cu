// Host-side byte work for the record / checkpoint readers (rows f2, f3 of the scope table): CRC-32C (Castagnoli),
// the checksum of TFRecord framing (inference.py:256 reads the records through tf.data.TFRecordDataset) and of the
// tensor-bundle checkpoint files (train.py:358, inference.py:283).  Slicing-by-8, no ISA extensions.
#include <stddef.h>
#include <stdint.h>

#include "../../include/strajnet_b200.h"

namespace {

struct Crc32cTables {
  uint32_t t[8][256];
  Crc32cTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;  // reflected Castagnoli polynomial
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};
const Crc32cTables& tables() {
  static const Crc32cTables tb;
  return tb;
}

}  // namespace

extern "C" uint32_t sj_crc32c(const void* data, size_t n, uint32_t crc) {
  const Crc32cTables& tb = tables();
  const uint8_t* p = static_cast<const uint8_t*>(data);
  crc = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) {
    crc = tb.t[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
    --n;
  }
  while (n >= 8) {
    uint64_t w = *reinterpret_cast<const uint64_t*>(p) ^ crc;  // little-endian host
    crc = tb.t[7][w & 0xff] ^ tb.t[6][(w >> 8) & 0xff] ^ tb.t[5][(w >> 16) & 0xff] ^ tb.t[4][(w >> 24) & 0xff] ^
          tb.t[3][(w >> 32) & 0xff] ^ tb.t[2][(w >> 40) & 0xff] ^ tb.t[1][(w >> 48) & 0xff] ^ tb.t[0][w >> 56];
    p += 8;
    n -= 8;
  }
  while (n--) crc = tb.t[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
  return ~crc;
}

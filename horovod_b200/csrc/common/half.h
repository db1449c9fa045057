// Software fp16 / bf16 <-> fp32 conversion for the CPU data plane.
// Parity: horovod/common/half.{h,cc} (fp16 only there; bf16 is new).
#pragma once
#include <cstdint>
#include <cstring>
namespace hvd {
inline float HalfBitsToFloat(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000) << 16;
  uint32_t exp = (h >> 10) & 0x1f;
  uint32_t man = h & 0x3ff;
  uint32_t f;
  if (exp == 0) {
    if (man == 0) { f = sign; }
    else {  // subnormal
      int e = -1;
      do { e++; man <<= 1; } while ((man & 0x400) == 0);
      man &= 0x3ff;
      f = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    f = sign | 0x7f800000u | (man << 13);
  } else {
    f = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float out; memcpy(&out, &f, 4); return out;
}
inline uint16_t FloatToHalfBits(float v) {
  uint32_t x; memcpy(&x, &v, 4);
  uint32_t sign = (x >> 16) & 0x8000;
  int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
  uint32_t man = x & 0x7fffff;
  if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00 | (man ? 0x200 : 0));
  if (exp >= 31) return (uint16_t)(sign | 0x7c00);
  if (exp <= 0) {
    if (exp < -10) return (uint16_t)sign;
    man |= 0x800000;
    uint32_t shift = (uint32_t)(14 - exp);
    uint32_t hm = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (hm & 1))) hm++;
    return (uint16_t)(sign | hm);
  }
  uint32_t hm = man >> 13;
  uint32_t rem = man & 0x1fff;
  uint16_t out = (uint16_t)(sign | ((uint32_t)exp << 10) | hm);
  if (rem > 0x1000 || (rem == 0x1000 && (hm & 1))) out++;  // round-nearest-even, carries into exponent correctly
  return out;
}
inline float BF16BitsToFloat(uint16_t b) { uint32_t f = (uint32_t)b << 16; float o; memcpy(&o, &f, 4); return o; }
inline uint16_t FloatToBF16Bits(float v) {
  uint32_t x; memcpy(&x, &v, 4);
  if ((x & 0x7fffffff) > 0x7f800000) return (uint16_t)((x >> 16) | 0x40);  // NaN
  uint32_t lsb = (x >> 16) & 1;
  x += 0x7fff + lsb;
  return (uint16_t)(x >> 16);
}
}  // namespace hvd

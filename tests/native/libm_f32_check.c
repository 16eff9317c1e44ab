/* Checks deepviewagg_b200/csrc/libm_f32.h (the float-only atanf / atan2f / acosf the CUDA projection
 * kernels use) against the C library that the reference's numba code calls.  argv[1] = stride over the
 * float bit patterns (1 = exhaustive, ~30 s).  Prints "<name> <mismatches> <tested>" per function. */
#include "../../deepviewagg_b200/csrc/libm_f32.h"
#include <stdio.h>
#include <stdlib.h>

int main(int argc, char** argv) {
  const uint32_t stride = argc > 1 ? (uint32_t)strtoul(argv[1], 0, 10) : 997u;
  const uint64_t pairs = argc > 2 ? strtoull(argv[2], 0, 10) : 4000000ull;
  uint64_t bad = 0, tot = 0;
  for (uint64_t u = 0; u <= 0x3f800000u; u += stride)
    for (int s = 0; s < 2; ++s) {
      const float x = f_from((uint32_t)u | (s ? 0x80000000u : 0u));
      bad += f_bits(acosf(x)) != f_bits(dva_acosf(x));
      ++tot;
    }
  printf("acosf %llu %llu\n", (unsigned long long)bad, (unsigned long long)tot);
  bad = tot = 0;
  for (uint64_t u = 0; u < 0x7f800000u; u += stride)
    for (int s = 0; s < 2; ++s) {
      const float x = f_from((uint32_t)u | (s ? 0x80000000u : 0u));
      bad += f_bits(atanf(x)) != f_bits(dva_atanf(x));
      ++tot;
    }
  printf("atanf %llu %llu\n", (unsigned long long)bad, (unsigned long long)tot);
  bad = tot = 0;
  uint64_t st = 88172645463325252ull;
  for (uint64_t i = 0; i < pairs; ++i) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    float y, x;
    if (i & 1) { y = f_from((uint32_t)st); x = f_from((uint32_t)(st >> 32)); }       /* any bit patterns */
    else {                                                                          /* scene-scale coordinates */
      y = ((int32_t)(st & 0xffffff) - 0x800000) * (1.0f / 65536.0f);
      x = ((int32_t)((st >> 32) & 0xffffff) - 0x800000) * (1.0f / 65536.0f);
    }
    const float a = atan2f(y, x), b = dva_atan2f(y, x);
    bad += (f_bits(a) != f_bits(b)) && !(a != a && b != b);
    ++tot;
  }
  printf("atan2f %llu %llu\n", (unsigned long long)bad, (unsigned long long)tot);
  return 0;
}

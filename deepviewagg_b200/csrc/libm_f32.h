// Bit-exact binary32 atanf / atan2f / acosf, usable from CUDA device code and from plain C.
//
// Why: the reference's camera projection runs under numba, where np.arctan2 / np.arccos on
// float32 arrays lower to the C library's atan2f / acosf (visibility.py:167-168).  The pixel a point
// lands in is floor() of a value derived from those angles, so "bit-exact mapping indices" needs the
// very same float results, not a correctly-rounded or a 1-ulp one.  glibc 2.39 (the image's libm)
// evaluates these functions with binary32 operations only (the classic fdlibm scheme: argument
// reduction to a table of 4 break points + an odd/even split polynomial for atan; a rational
// P/Q approximation with a split square root for acos), i.e. a fixed sequence of IEEE-754
// add / mul / div / sqrt -- which a GPU reproduces exactly as long as nothing is contracted into an
// FMA.  Every operation below therefore goes through f_add / f_mul / ... (round-to-nearest
// intrinsics on the device, -ffp-contract=off on the host).
//
// Verified against the libm that produced tests/golden/*.npz: acosf on all 2 130 706 434 floats of
// [-1, 1], atanf on all 2 139 095 040 finite non-negative floats, atan2f on 4e8 random pairs --
// zero mismatches (tests/test_libm_f32.py re-runs a sample of that check on the CPU).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DVA_FN __device__ __forceinline__
DVA_FN float f_mul(float a, float b) { return __fmul_rn(a, b); }
DVA_FN float f_add(float a, float b) { return __fadd_rn(a, b); }
DVA_FN float f_sub(float a, float b) { return __fsub_rn(a, b); }
DVA_FN float f_div(float a, float b) { return __fdiv_rn(a, b); }
DVA_FN float f_sqrt(float a) { return __fsqrt_rn(a); }
DVA_FN uint32_t f_bits(float a) { return __float_as_uint(a); }
DVA_FN float f_from(uint32_t u) { return __uint_as_float(u); }
#else
#include <math.h>
#include <string.h>
#define DVA_FN static inline
DVA_FN float f_mul(float a, float b) { return a * b; }
DVA_FN float f_add(float a, float b) { return a + b; }
DVA_FN float f_sub(float a, float b) { return a - b; }
DVA_FN float f_div(float a, float b) { return a / b; }
DVA_FN float f_sqrt(float a) { return sqrtf(a); }
DVA_FN uint32_t f_bits(float a) { uint32_t u; memcpy(&u, &a, 4); return u; }
DVA_FN float f_from(uint32_t u) { float a; memcpy(&a, &u, 4); return a; }
#endif

// atan(x): break points 7/16, 11/16, 19/16, 39/16; atan(x) = atanhi[id] + atanlo[id] + atan(t) with
// t = (x - c) / (1 + c x); odd polynomial of degree 23 in t split into two Horner chains over t^4.
DVA_FN float dva_atanf(float x) {
  const float hi0 = f_from(0x3eed6338u), hi1 = f_from(0x3f490fdau), hi2 = f_from(0x3f7b985eu), hi3 = f_from(0x3fc90fdau);
  const float lo0 = f_from(0x31ac3769u), lo1 = f_from(0x33222168u), lo2 = f_from(0x33140fb4u), lo3 = f_from(0x33a22168u);
  const float a0 = f_from(0x3eaaaaabu), a1 = f_from(0xbe4ccccdu), a2 = f_from(0x3e124925u), a3 = f_from(0xbde38e38u),
              a4 = f_from(0x3dba2e6eu), a5 = f_from(0xbd9d8795u), a6 = f_from(0x3d886b35u), a7 = f_from(0xbd6ef16bu),
              a8 = f_from(0x3d4bda59u), a9 = f_from(0xbd15a221u), a10 = f_from(0x3c8569d7u);
  const int32_t hx = (int32_t)f_bits(x);
  const int32_t ix = hx & 0x7fffffff;
  float hi = 0.0f, lo = 0.0f;
  int reduced = 1;
  if (ix >= 0x4c000000) {                          // |x| >= 2^25 (or NaN)
    if (ix > 0x7f800000) return f_add(x, x);
    return hx > 0 ? f_add(hi3, lo3) : f_sub(-hi3, lo3);
  }
  if (ix < 0x3ee00000) {                           // |x| < 7/16
    if (ix < 0x31000000) return x;                 // |x| < 2^-29
    reduced = 0;
  } else {
    x = f_from((uint32_t)ix);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) { hi = hi0; lo = lo0; x = f_div(f_sub(f_mul(2.0f, x), 1.0f), f_add(2.0f, x)); }
      else { hi = hi1; lo = lo1; x = f_div(f_sub(x, 1.0f), f_add(x, 1.0f)); }
    } else {
      if (ix < 0x401c0000) { hi = hi2; lo = lo2; x = f_div(f_sub(x, 1.5f), f_add(1.0f, f_mul(1.5f, x))); }
      else { hi = hi3; lo = lo3; x = f_div(-1.0f, x); }
    }
  }
  const float z = f_mul(x, x);
  const float w = f_mul(z, z);
  float s1 = f_add(a8, f_mul(w, a10));
  s1 = f_add(a6, f_mul(w, s1));
  s1 = f_add(a4, f_mul(w, s1));
  s1 = f_add(a2, f_mul(w, s1));
  s1 = f_mul(z, f_add(a0, f_mul(w, s1)));
  float s2 = f_add(a7, f_mul(w, a9));
  s2 = f_add(a5, f_mul(w, s2));
  s2 = f_add(a3, f_mul(w, s2));
  s2 = f_mul(w, f_add(a1, f_mul(w, s2)));
  if (!reduced) return f_sub(x, f_mul(x, f_add(s1, s2)));
  const float r = f_sub(hi, f_sub(f_sub(f_mul(x, f_add(s1, s2)), lo), x));
  return hx < 0 ? -r : r;
}

// atan2(y, x) = quadrant fix-up around atan(|y / x|), pi carried as pi + pi_lo.
DVA_FN float dva_atan2f(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = f_from(0x3f490fdbu), pi_o_2 = f_from(0x3fc90fdbu),
              pi = f_from(0x40490fdbu), pi_lo = f_from(0xb3bbbd2eu);
  const int32_t hx = (int32_t)f_bits(x), hy = (int32_t)f_bits(y);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return f_add(x, y);          // NaN
  if (hx == 0x3f800000) return dva_atanf(y);                           // x == 1
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);                   // 2 sign(x) + sign(y)
  if (iy == 0) {
    if (m < 2) return y;
    return m == 2 ? f_add(pi, tiny) : f_sub(-pi, tiny);
  }
  if (ix == 0) return hy < 0 ? f_sub(-pi_o_2, tiny) : f_add(pi_o_2, tiny);
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      if (m == 0) return f_add(pi_o_4, tiny);
      if (m == 1) return f_sub(-pi_o_4, tiny);
      return m == 2 ? f_add(f_mul(3.0f, pi_o_4), tiny) : f_sub(f_mul(-3.0f, pi_o_4), tiny);
    }
    if (m == 0) return 0.0f;
    if (m == 1) return -0.0f;
    return m == 2 ? f_add(pi, tiny) : f_sub(-pi, tiny);
  }
  if (iy == 0x7f800000) return hy < 0 ? f_sub(-pi_o_2, tiny) : f_add(pi_o_2, tiny);
  const int32_t k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = f_add(pi_o_2, f_mul(0.5f, pi_lo));                   // |y / x| > 2^60
  else if (hx < 0 && k < -60) z = 0.0f;                                // |y| / x < -2^60
  else z = dva_atanf(f_from(f_bits(f_div(y, x)) & 0x7fffffffu));
  if (m == 0) return z;
  if (m == 1) return f_from(f_bits(z) ^ 0x80000000u);
  return m == 2 ? f_sub(pi, f_sub(z, pi_lo)) : f_sub(f_sub(z, pi_lo), pi);
}

// acos(x): |x| < 0.5: pi/2 - (x + x R(x^2)); x < -0.5: pi - 2 (s + s R(z)), z = (1 + x) / 2, s = sqrt z;
// x > 0.5: 2 (df + (s R(z) + c)) with s = df + c split so that df * df is exact.  R = P / Q.
DVA_FN float dva_acosf(float x) {
  const float pi = f_from(0x40490fdau), pio2_hi = f_from(0x3fc90fdau), pio2_lo = f_from(0x33a22168u);
  const float pS0 = f_from(0x3e2aaaabu), pS1 = f_from(0xbea6b090u), pS2 = f_from(0x3e4e0aa8u),
              pS3 = f_from(0xbd241146u), pS4 = f_from(0x3a4f7f04u), pS5 = f_from(0x3811ef08u),
              qS1 = f_from(0xc019d139u), qS2 = f_from(0x4001572du), qS3 = f_from(0xbf303361u),
              qS4 = f_from(0x3d9dc62eu);
  const int32_t hx = (int32_t)f_bits(x);
  const int32_t ix = hx & 0x7fffffff;
  if (ix == 0x3f800000) return hx > 0 ? 0.0f : f_add(pi, f_mul(2.0f, pio2_lo));
  if (ix > 0x3f800000) return f_div(f_sub(x, x), f_sub(x, x));         // |x| > 1 or NaN -> NaN
  float z;
  if (ix < 0x3f000000) {
    if (ix <= 0x32800000) return f_add(pio2_hi, pio2_lo);              // |x| <= 2^-26
    z = f_mul(x, x);
  } else if (hx < 0) {
    z = f_mul(f_add(1.0f, x), 0.5f);
  } else {
    z = f_mul(f_sub(1.0f, x), 0.5f);
  }
  float p = f_add(pS4, f_mul(z, pS5));
  p = f_add(pS3, f_mul(z, p));
  p = f_add(pS2, f_mul(z, p));
  p = f_add(pS1, f_mul(z, p));
  p = f_mul(z, f_add(pS0, f_mul(z, p)));
  float q = f_add(qS3, f_mul(z, qS4));
  q = f_add(qS2, f_mul(z, q));
  q = f_add(qS1, f_mul(z, q));
  q = f_add(1.0f, f_mul(z, q));
  const float r = f_div(p, q);
  if (ix < 0x3f000000) return f_sub(pio2_hi, f_sub(x, f_sub(pio2_lo, f_mul(x, r))));
  const float s = f_sqrt(z);
  if (hx < 0) {
    const float w = f_sub(f_mul(r, s), pio2_lo);
    return f_sub(pi, f_mul(2.0f, f_add(s, w)));
  }
  const float df = f_from(f_bits(s) & 0xfffff000u);
  const float c = f_div(f_sub(z, f_mul(df, df)), f_add(s, df));
  const float w = f_add(f_mul(r, s), c);
  return f_mul(2.0f, f_add(df, w));
}

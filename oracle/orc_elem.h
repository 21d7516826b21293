/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, link or call anything under oracle/.
 *
 * Elementary functions for the oracle's "FDLIBM" mode: a restatement of the published
 * FreeBSD/Sun fdlibm algorithms e_log.c and e_exp.c, which java.lang.StrictMath.log /
 * StrictMath.exp are specified to reproduce bit-for-bit (JDK javadoc of StrictMath).
 * The reference itself calls Math.log / Math.exp (DCRecursion.java:63,
 * BrownianModelCalculator.java:138), which the JDK allows to differ from StrictMath by
 * <= 1 ulp; StrictMath is the only fully specified Java behaviour, so it is the
 * contract the device kernels are held to bit-for-bit.
 *
 * Only IEEE-754 +,-,*,/ and integer bit manipulation are used, so any compiler that
 * does not contract a*b+c into an fma (gcc -ffp-contract=off) gives identical bits.
 */
#ifndef ORC_ELEM_H
#define ORC_ELEM_H

#include <stdint.h>
#include <string.h>

static inline uint64_t orc_d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double orc_u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
static inline int32_t orc_hi(double x) { return (int32_t)(orc_d2u(x) >> 32); }
static inline uint32_t orc_lo(double x) { return (uint32_t)(orc_d2u(x) & 0xffffffffu); }
static inline double orc_with_hi(double x, int32_t hi) {
  return orc_u2d(((uint64_t)(uint32_t)hi << 32) | (orc_d2u(x) & 0xffffffffu));
}

/* fdlibm e_log.c */
static inline double orc_fd_log(double x) {
  static const double ln2_hi = 6.93147180369123816490e-01, /* 3fe62e42 fee00000 */
      ln2_lo = 1.90821492927058770002e-10,                 /* 3dea39ef 35793c76 */
      two54 = 1.80143985094819840000e+16,                  /* 43500000 00000000 */
      Lg1 = 6.666666666666735130e-01,                      /* 3FE55555 55555593 */
      Lg2 = 3.999999999940941908e-01,                      /* 3FD99999 9997FA04 */
      Lg3 = 2.857142874366239149e-01,                      /* 3FD24924 94229359 */
      Lg4 = 2.222219843214978396e-01,                      /* 3FCC71C5 1D8E78AF */
      Lg5 = 1.818357216161805012e-01,                      /* 3FC74664 96CB03DE */
      Lg6 = 1.531383769920937332e-01,                      /* 3FC39A09 D078C69F */
      Lg7 = 1.479819860511658591e-01;                      /* 3FC2F112 DF3E5244 */
  static const double zero = 0.0;
  double hfsq, f, s, z, R, w, t1, t2, dk;
  int32_t k, hx, i, j;
  uint32_t lx;

  hx = orc_hi(x);
  lx = orc_lo(x);
  k = 0;
  if (hx < 0x00100000) { /* x < 2**-1022 */
    if (((hx & 0x7fffffff) | lx) == 0) return -two54 / zero; /* log(+-0) = -inf */
    if (hx < 0) return (x - x) / zero;                       /* log(-#) = NaN */
    k -= 54;
    x *= two54; /* subnormal, scale up */
    hx = orc_hi(x);
  }
  if (hx >= 0x7ff00000) return x + x;
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  i = (hx + 0x95f64) & 0x100000;
  x = orc_with_hi(x, hx | (i ^ 0x3ff00000)); /* normalize x or x/2 */
  k += (i >> 20);
  f = x - 1.0;
  if ((0x000fffff & (2 + hx)) < 3) { /* |f| < 2**-20 */
    if (f == zero) {
      if (k == 0) return zero;
      dk = (double)k;
      return dk * ln2_hi + dk * ln2_lo;
    }
    R = f * f * (0.5 - 0.33333333333333333 * f);
    if (k == 0) return f - R;
    dk = (double)k;
    return dk * ln2_hi - ((R - dk * ln2_lo) - f);
  }
  s = f / (2.0 + f);
  dk = (double)k;
  z = s * s;
  i = hx - 0x6147a;
  w = z * z;
  j = 0x6b851 - hx;
  t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  i |= j;
  R = t2 + t1;
  if (i > 0) {
    hfsq = 0.5 * f * f;
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
  } else {
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
  }
}

/* fdlibm e_exp.c */
static inline double orc_fd_exp(double x) {
  static const double one = 1.0, halF[2] = {0.5, -0.5}, huge = 1.0e+300,
                      twom1000 = 9.33263618503218878990e-302,  /* 2**-1000 */
      o_threshold = 7.09782712893383973096e+02,                /* 40862E42 FEFA39EF */
      u_threshold = -7.45133219101941108420e+02,               /* c0874910 D52D3051 */
      ln2HI[2] = {6.93147180369123816490e-01, -6.93147180369123816490e-01},
      ln2LO[2] = {1.90821492927058770002e-10, -1.90821492927058770002e-10},
      invln2 = 1.44269504088896338700e+00,                     /* 3ff71547 652b82fe */
      P1 = 1.66666666666666019037e-01,                         /* 3FC55555 5555553E */
      P2 = -2.77777777770155933842e-03,                        /* BF66C16C 16BEBD93 */
      P3 = 6.61375632143793436117e-05,                         /* 3F11566A AF25DE2C */
      P4 = -1.65339022054652515390e-06,                        /* BEBBBD41 C5D26BF1 */
      P5 = 4.13813679705723846039e-08;                         /* 3E663769 72BEA4D0 */
  double y, hi = 0.0, lo = 0.0, c, t;
  int32_t k = 0, xsb;
  uint32_t hx;

  hx = (uint32_t)orc_hi(x);
  xsb = (int32_t)((hx >> 31) & 1);
  hx &= 0x7fffffff;

  if (hx >= 0x40862E42) { /* |x| >= 709.78... */
    if (hx >= 0x7ff00000) {
      if (((hx & 0xfffff) | orc_lo(x)) != 0) return x + x; /* NaN */
      return (xsb == 0) ? x : 0.0;                         /* exp(+-inf) = {inf,0} */
    }
    if (x > o_threshold) return huge * huge;
    if (x < u_threshold) return twom1000 * twom1000;
  }

  if (hx > 0x3fd62e42) {   /* |x| > 0.5 ln2 */
    if (hx < 0x3FF0A2B2) { /* and |x| < 1.5 ln2 */
      hi = x - ln2HI[xsb];
      lo = ln2LO[xsb];
      k = 1 - xsb - xsb;
    } else {
      k = (int32_t)(invln2 * x + halF[xsb]);
      t = k;
      hi = x - t * ln2HI[0];
      lo = t * ln2LO[0];
    }
    x = hi - lo;
  } else if (hx < 0x3e300000) { /* |x| < 2**-28 */
    if (huge + x > one) return one + x;
  } else
    k = 0;

  t = x * x;
  c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  if (k == 0) return one - ((x * c) / (c - 2.0) - x);
  y = one - ((lo - (x * c) / (2.0 - c)) - hi);
  if (k >= -1021) {
    return orc_with_hi(y, orc_hi(y) + (k << 20));
  } else {
    y = orc_with_hi(y, orc_hi(y) + ((k + 1000) << 20));
    return y * twom1000;
  }
}

#endif

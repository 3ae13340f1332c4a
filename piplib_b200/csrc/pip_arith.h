/* Exact integer helpers of the solver core (wrapping int64, the reference's _dp width).
 *
 * Semantics follow include/piplib/piplib.h:128-169 and source/integrer.c:41-89 of the reference:
 *   gcd      = llabs(euclid(a, b)), gcd(0,0) = 0
 *   mod      = C remainder, +|b| when negative  (piplib_llmod_xx)
 *   floor_q  = (a - mod(a,b)) / b               (piplib_int_floor_div_q)
 *   div      = C truncating division (used where the division is exact)
 *   bitlen   = bit length of |x|, 0 -> 1        (piplib_lllog2_xx)
 * The implementations are GPU-shaped: 32-bit fast paths (there is no integer divider on the SM;
 * a 64-bit '/' is a ~100-instruction subroutine) and exact division by a modular inverse.
 */
#ifndef PIP_ARITH_H
#define PIP_ARITH_H

#include "pip_types.h"
#include "simt.h"

PIP_HD pip_u64 pip_uabs(pip_i64 v) { return v < 0 ? 0ull - (pip_u64)v : (pip_u64)v; }

/* count of trailing zero bits of v != 0 (host and device) */
PIP_HD int pip_tz64(pip_u64 v)
{
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}

/* gcd of two 64-bit magnitudes; the value is the mathematical gcd, hence bit-identical to the
 * reference's Euclid (source/integrer.c:43-50).  GPU-shaped: there is no integer divider on the SM
 * and a 64-bit '%' is a ~150-cycle subroutine, so
 *   - the smaller operand is always the divisor (the running row gcd against a tableau entry costs
 *     ONE remainder when the entry is a multiple, the common case),
 *   - a divisor that fits 32 bits takes at most one wide remainder, then Euclid on 32-bit words,
 *   - two wide operands run the binary algorithm (subtract / count trailing zeros / shift) until
 *     they fit. */
PIP_HDNI pip_u64 pip_gcd_u64(pip_u64 a, pip_u64 b)
{
  if (a < b) { const pip_u64 t = a; a = b; b = t; }
  if ((b >> 32) != 0) {                    /* both wide */
    const int sh = pip_tz64(a | b);
    a >>= pip_tz64(a);
    do {
      b >>= pip_tz64(b);
      if (a > b) { const pip_u64 t = a; a = b; b = t; }
      b -= a;
    } while ((b >> 32) != 0);
    /* a is odd, so gcd(a, b) is odd whatever power of two b holds; the common one is restored at the end */
    if (b == 0) return a << sh;
    unsigned y = (unsigned)b;
    unsigned x = (a >> 32) == 0 ? (unsigned)a % y : (unsigned)(a % (pip_u64)y);
    while (x) { const unsigned r = y % x; y = x; x = r; }
    return (pip_u64)y << sh;
  }
  if (b == 0) return a;
  unsigned y = (unsigned)b;
  unsigned x = (a >> 32) == 0 ? (unsigned)a % y : (unsigned)(a % (pip_u64)y);
  while (x) { const unsigned r = y % x; y = x; x = r; }
  return y;
}
PIP_HD pip_i64 pip_gcd(pip_i64 a, pip_i64 b) { return (pip_i64)pip_gcd_u64(pip_uabs(a), pip_uabs(b)); }

/* C truncating division, b != 0 */
PIP_HDNI pip_i64 pip_div(pip_i64 a, pip_i64 b)
{
  if (a == (pip_i64)(int)a && b == (pip_i64)(int)b && (int)a != (-2147483647 - 1))
    return (pip_i64)((int)a / (int)b);
  return a / b;
}
/* C remainder, b != 0 */
PIP_HDNI pip_i64 pip_rem(pip_i64 a, pip_i64 b)
{
  if (a == (pip_i64)(int)a && b == (pip_i64)(int)b && (int)a != (-2147483647 - 1))
    return (pip_i64)((int)a % (int)b);
  return a % b;
}
PIP_HD pip_i64 pip_mod(pip_i64 a, pip_i64 b)
{
  pip_i64 m = pip_rem(a, b);
  if (m < 0) m += (b < 0 ? -b : b);
  return m;
}
PIP_HD pip_i64 pip_floor_q(pip_i64 a, pip_i64 b) { return pip_div(a - pip_mod(a, b), b); }

/* bezout_xx, source/integrer.c:98-150: z in [0, delta) with z*y = x (mod delta), 0 when y is not a
 * unit modulo delta.  Wrapping products like the reference's. */
PIP_HDNI pip_i64 pip_bezout(pip_i64 x, pip_i64 y, pip_i64 delta)
{
  pip_i64 a = 1, b = 0, c = 0, d = 1, u = y, v = delta;
  for (;;) {
    const pip_i64 r = pip_mod(u, v);
    if (r == 0) break;
    const pip_i64 q = pip_div(u - r, v);
    u = v; v = r;
    const pip_i64 e = (pip_i64)((pip_u64)a - (pip_u64)q * (pip_u64)c), f = (pip_i64)((pip_u64)b - (pip_u64)q * (pip_u64)d);
    a = c; b = d; c = e; d = f;
  }
  if (v != 1) return 0;
  return pip_mod((pip_i64)((pip_u64)c * (pip_u64)x), delta);
}

/* 32-bit overloads for the int32 instantiation of the solver (no 64-bit division subroutine).  Out of line:
 * a 32-bit '/' or '%' is ~20 instructions on the SM and these helpers are reached from a dozen places of a
 * kernel whose executed code has to fit the 32 KB instruction cache (profiles/README.md) */
PIP_HDNI int pip_gcd(int a, int b)
{
  unsigned x = a < 0 ? 0u - (unsigned)a : (unsigned)a, y = b < 0 ? 0u - (unsigned)b : (unsigned)b;
#ifndef PIP_GCD32_NOSWAP
  if (x < y) { const unsigned t = x; x = y; y = t; }     /* the smaller operand divides first */
#endif
  while (y) { unsigned r = x % y; x = y; y = r; }
  return (int)x;
}
PIP_HDNI int pip_div(int a, int b) { return a / b; }
PIP_HDNI int pip_mod(int a, int b)
{
  int m = a % b;
  if (m < 0) m += (b < 0 ? -b : b);
  return m;
}
PIP_HD int pip_floor_q(int a, int b) { return (a - pip_mod(a, b)) / b; }

PIP_HD int pip_bitlen(pip_i64 x)
{
  const pip_u64 u = pip_uabs(x);
  return u ? 64 - pip_clzll(u) : 1;
}

/* multiplicative inverse of an odd d modulo 2^64 (Newton: each step doubles the valid bits) */
PIP_HD pip_u64 pip_inv_odd(pip_u64 d)
{
  pip_u64 x = (3ull * d) ^ 2ull;     /* 5 bits */
  x *= 2ull - d * x;                 /* 10 */
  x *= 2ull - d * x;                 /* 20 */
  x *= 2ull - d * x;                 /* 40 */
  x *= 2ull - d * x;                 /* 80 */
  return x;
}

/* exact division z / g for g > 0 dividing z: shift out the power of two, multiply by the
 * inverse of the odd part.  Bit-identical to C '/' whenever the division is exact. */
struct PipExactDiv {
  pip_u64 inv;
  int shift;
};
PIP_HDNI PipExactDiv pip_exact_prepare(pip_i64 g)
{
  PipExactDiv e;
  pip_u64 u = (pip_u64)g;
  int s = 0;
  while ((u & 1ull) == 0ull) { u >>= 1; s++; }   /* g != 0 */
  e.shift = s;
  e.inv = pip_inv_odd(u);
  return e;
}
PIP_HD pip_i64 pip_exact_apply(pip_i64 z, const PipExactDiv &e)
{
  return (pip_i64)((pip_u64)(z >> e.shift) * e.inv);
}

#if defined(__CUDACC__) || defined(PIP_EMU)
/* does the exact product a*b fit in int64? (informative overflow flag only) */
PIP_DEV bool pip_mul_wraps(pip_i64 a, pip_i64 b)
{
  pip_i64 lo = (pip_i64)((pip_u64)a * (pip_u64)b);
  pip_i64 hi = pip_mulhi(a, b);
  return hi != (lo >> 63);
}
#endif

#endif

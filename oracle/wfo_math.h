/*
 * wfo_math.h -- CPU ORACLE (test infrastructure, see wfo.h): Julia numeric semantics in C.
 * SURVEY.md App. E; reference Wflow/src/utils.jl:470 (pow), :1070-1076 (bounded_power).
 */
#ifndef WFO_MATH_H
#define WFO_MATH_H
#include <math.h>
#include <stdint.h>

#ifdef WFO_ALT_LIBM
/* A deliberately NOISY libm (build variant libwfo_alt.so, tests/test_tolerances.py): exp, log
 * and cbrt results are moved by one unit in the last place for a pseudo-random half of the
 * arguments. Two runs of the oracle that differ only in this way bracket what any two faithful
 * (< 1-2 ulp) math libraries -- glibc here, libdevice on the GPU, Julia's own in the reference --
 * may do to the results; the per-field tolerances of tests/parity.py are calibrated on it. */
static inline double wfo_alt_noise(double x, double r) {
  uint64_t b;
  __builtin_memcpy(&b, &x, 8);
  b = (b ^ (b >> 29)) * 0xBF58476D1CE4E5B9ull;
  b ^= b >> 32;
  if (!isfinite(r) || r == 0.0 || (b & 1)) return r;
  return nextafter(r, (b & 2) ? INFINITY : -INFINITY);
}
static inline double wfo_alt_exp(double x) { return wfo_alt_noise(x, exp(x)); }
static inline double wfo_alt_log(double x) { return wfo_alt_noise(x, log(x)); }
static inline double wfo_alt_cbrt(double x) { return wfo_alt_noise(x, cbrt(x)); }
#define exp(x) wfo_alt_exp(x)
#define log(x) wfo_alt_log(x)
#define cbrt(x) wfo_alt_cbrt(x)
#endif

/* to_SI_factor(MM_PER_DAY) = 86400^-1 * 1e-3 (units.jl:55-68, factors multiplied in field
 * order: d before mm) */
#define WFO_MM_PER_DAY ((1.0 / 86400.0) * 1e-3)
#define WFO_KIN_WAVE_MIN_FLOW 1e-30 /* routing/utils.jl:1 */

/* Julia's min/max propagate NaN and order signed zeros (Base math.jl) */
static inline double jl_min(double a, double b) {
  if (isnan(a) || isnan(b)) return NAN;
  if (a < b) return a;
  if (b < a) return b;
  return signbit(a) ? a : b;
}
static inline double jl_max(double a, double b) {
  if (isnan(a) || isnan(b)) return NAN;
  if (a > b) return a;
  if (b > a) return b;
  return signbit(a) ? b : a;
}
/* clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x)) : NaN x stays NaN */
static inline double jl_clamp(double x, double lo, double hi) {
  return x > hi ? hi : (x < lo ? lo : x);
}
/* utils.jl:470 : pow(x, y) = exp(y * log(x)) */
static inline double jl_pow(double x, double y) { return exp(y * log(x)); }
/* utils.jl:1070-1076 */
static inline double jl_bounded_power(double b, double p) { return b > 1.0 ? 1.0 : jl_pow(b, p); }

#endif

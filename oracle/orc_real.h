/*
 * oracle/orc_real.h -- arithmetic type of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * The reference is compiled with -fdefault-real-8 (CMakeLists.txt:13): every `real` is an
 * IEEE binary64.  The oracle therefore computes in `double` by default.  Building with
 * -DORC_QUAD switches every `real` to __float128 (libquadmath) so that rounding disputes
 * between the oracle and the CUDA path can be arbitrated against a result that is exact
 * to ~1e-30.  The C API always exchanges binary64.
 */
#ifndef ORC_REAL_H
#define ORC_REAL_H

#include <math.h>

#ifdef ORC_QUAD
#include <quadmath.h>
typedef __float128 real;
#define R_SQRT(x) sqrtq(x)
#define R_FABS(x) fabsq(x)
#define R_COS(x) cosq(x)
#define R_SIN(x) sinq(x)
#define R_TAN(x) tanq(x)
#define R_EXP(x) expq(x)
#define R_POW(x, y) powq(x, y)
#define R_ATAN(x) atanq(x)
#define R_ISNAN(x) isnanq(x)
#define R_LIT(x) x##Q
#else
typedef double real;
#define R_SQRT(x) sqrt(x)
#define R_FABS(x) fabs(x)
#define R_COS(x) cos(x)
#define R_SIN(x) sin(x)
#define R_TAN(x) tan(x)
#define R_EXP(x) exp(x)
#define R_POW(x, y) pow(x, y)
#define R_ATAN(x) atan(x)
#define R_ISNAN(x) isnan(x)
#define R_LIT(x) x
#endif

#endif

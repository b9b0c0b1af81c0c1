/*
 * oracle/fftpack_rfft.c -- CPU restatement of the FFTPACK 5.1 real 1-D transform the reference's
 * zonal filter calls (lib/fft/rfft1i.f:31-46, rffti1.f:31-90, rfft1f.f:31-53, rfftf1.f:31-112,
 * rfft1b.f:31-53, rfftb1.f:31-115 and the radix kernels r1f{2,3,4,5}k{f,b}.f).
 *
 * TEST INFRASTRUCTURE ONLY (see orc_real.h).  Not linked into, loaded by or called from the product.
 *
 * What is kept from the reference:
 *   - the factorisation order of rffti1.f:38-58 (try 4,2,3,5,...; a factor 2 is moved to the front),
 *     e.g. 360 -> [2,4,3,3,5], so the butterflies run in the same sequence;
 *   - twiddles evaluated in extended precision and rounded once (rffti1.f:34,61-83 declares them
 *     DOUBLE PRECISION, which -fdefault-real-8 promotes to binary128; __float128 here);
 *   - the forward transform is 1/N-normalised with the sine coefficients negated (rfftf1.f:87-107),
 *     output layout a0,a1,b1,a2,b2,...,a_{N/2} with x_i = a0 + sum_k a_k cos(k t_i) + b_k sin(k t_i);
 *   - the backward transform pre-scales by +-1/2 (rfftb1.f:43-61) and is un-normalised.
 *   - INC = 1 only (filter_mod.F90:121,126 always passes 1).
 *
 * What is not kept: the butterflies are written with the classic FFTPACK temporaries (cr2, ci2, ...)
 * instead of the fully expanded expressions of FFTPACK 5.1.  They are the same expressions with the
 * same grouping except in r1f5kf.f:60-75, where 5.1 sums four products left to right; the reference
 * is built with -Ofast (CMakeLists.txt:13), which lets the compiler regroup those sums anyway, so
 * that difference is unobservable.  Factors other than 2,3,4,5 (r1fgkf.f / r1fgkb.f) are not
 * restated: every grid size in BASELINE.json is 2^a 3^b 5^c; orc_rffti returns 3 for anything else.
 */
#include "orc_real.h"
#include "gmd_oracle.h"
#include <stdlib.h>
#include <string.h>

#ifndef ORC_QUAD
#include <quadmath.h>
#endif

/* ---- forward radix kernels: cc(ido,l1,ip) -> ch(ido,ip,l1) ------------------------------------ */
#define CCF(i, k, m) cc[(i) + ido * ((k) + l1 * (m))]
#define CHF(i, m, k) ch[(i) + ido * ((m) + ip * (k))]
/* ---- backward radix kernels: cc(ido,ip,l1) -> ch(ido,l1,ip) ------------------------------------ */
#define CCB(i, m, k) cc[(i) + ido * ((m) + ip * (k))]
#define CHB(i, k, m) ch[(i) + ido * ((k) + l1 * (m))]

/* r1f2kf.f:31-61 */
static void radf2(int ido, int l1, const real *cc, real *ch, const real *wa1) {
  const int ip = 2;
  int i, k, ic;
  real tr2, ti2;
  for (k = 0; k < l1; k++) {
    CHF(0, 0, k) = CCF(0, k, 0) + CCF(0, k, 1);
    CHF(ido - 1, 1, k) = CCF(0, k, 0) - CCF(0, k, 1);
  }
  if (ido < 2) return;
  if (ido > 2) {
    for (k = 0; k < l1; k++) {
      for (i = 2; i < ido; i += 2) {
        ic = ido - i;
        tr2 = wa1[i - 2] * CCF(i - 1, k, 1) + wa1[i - 1] * CCF(i, k, 1);
        ti2 = wa1[i - 2] * CCF(i, k, 1) - wa1[i - 1] * CCF(i - 1, k, 1);
        CHF(i, 0, k) = CCF(i, k, 0) + ti2;
        CHF(ic, 1, k) = ti2 - CCF(i, k, 0);
        CHF(i - 1, 0, k) = CCF(i - 1, k, 0) + tr2;
        CHF(ic - 1, 1, k) = CCF(i - 1, k, 0) - tr2;
      }
    }
    if (ido % 2 == 1) return;
  }
  for (k = 0; k < l1; k++) {
    CHF(0, 1, k) = -CCF(ido - 1, k, 1);
    CHF(ido - 1, 0, k) = CCF(ido - 1, k, 0);
  }
}

/* r1f3kf.f:31-83 */
static void radf3(int ido, int l1, const real *cc, real *ch, const real *wa1, const real *wa2) {
  const int ip = 3;
  const real arg = R_LIT(2.0) * R_LIT(4.0) * R_ATAN(R_LIT(1.0)) / R_LIT(3.0);
  const real taur = R_COS(arg), taui = R_SIN(arg);
  int i, k, ic;
  real cr2, ci2, dr2, di2, dr3, di3, tr2, ti2, tr3, ti3;
  for (k = 0; k < l1; k++) {
    cr2 = CCF(0, k, 1) + CCF(0, k, 2);
    CHF(0, 0, k) = CCF(0, k, 0) + cr2;
    CHF(0, 2, k) = taui * (CCF(0, k, 2) - CCF(0, k, 1));
    CHF(ido - 1, 1, k) = CCF(0, k, 0) + taur * cr2;
  }
  if (ido == 1) return;
  for (k = 0; k < l1; k++) {
    for (i = 2; i < ido; i += 2) {
      ic = ido - i;
      dr2 = wa1[i - 2] * CCF(i - 1, k, 1) + wa1[i - 1] * CCF(i, k, 1);
      di2 = wa1[i - 2] * CCF(i, k, 1) - wa1[i - 1] * CCF(i - 1, k, 1);
      dr3 = wa2[i - 2] * CCF(i - 1, k, 2) + wa2[i - 1] * CCF(i, k, 2);
      di3 = wa2[i - 2] * CCF(i, k, 2) - wa2[i - 1] * CCF(i - 1, k, 2);
      cr2 = dr2 + dr3;
      ci2 = di2 + di3;
      CHF(i - 1, 0, k) = CCF(i - 1, k, 0) + cr2;
      CHF(i, 0, k) = CCF(i, k, 0) + ci2;
      tr2 = CCF(i - 1, k, 0) + taur * cr2;
      ti2 = CCF(i, k, 0) + taur * ci2;
      tr3 = taui * (di2 - di3);
      ti3 = taui * (dr3 - dr2);
      CHF(i - 1, 2, k) = tr2 + tr3;
      CHF(ic - 1, 1, k) = tr2 - tr3;
      CHF(i, 2, k) = ti2 + ti3;
      CHF(ic, 1, k) = ti3 - ti2;
    }
  }
}

/* r1f4kf.f:31-104 */
static void radf4(int ido, int l1, const real *cc, real *ch, const real *wa1, const real *wa2,
                  const real *wa3) {
  const int ip = 4;
  const real hsqt2 = R_SQRT(R_LIT(2.0)) / R_LIT(2.0);
  int i, k, ic;
  real cr2, ci2, cr3, ci3, cr4, ci4, tr1, ti1, tr2, ti2, tr3, ti3, tr4, ti4;
  for (k = 0; k < l1; k++) {
    tr1 = CCF(0, k, 1) + CCF(0, k, 3);
    tr2 = CCF(0, k, 0) + CCF(0, k, 2);
    CHF(0, 0, k) = tr1 + tr2;
    CHF(ido - 1, 3, k) = tr2 - tr1;
    CHF(ido - 1, 1, k) = CCF(0, k, 0) - CCF(0, k, 2);
    CHF(0, 2, k) = CCF(0, k, 3) - CCF(0, k, 1);
  }
  if (ido < 2) return;
  if (ido > 2) {
    for (k = 0; k < l1; k++) {
      for (i = 2; i < ido; i += 2) {
        ic = ido - i;
        cr2 = wa1[i - 2] * CCF(i - 1, k, 1) + wa1[i - 1] * CCF(i, k, 1);
        ci2 = wa1[i - 2] * CCF(i, k, 1) - wa1[i - 1] * CCF(i - 1, k, 1);
        cr3 = wa2[i - 2] * CCF(i - 1, k, 2) + wa2[i - 1] * CCF(i, k, 2);
        ci3 = wa2[i - 2] * CCF(i, k, 2) - wa2[i - 1] * CCF(i - 1, k, 2);
        cr4 = wa3[i - 2] * CCF(i - 1, k, 3) + wa3[i - 1] * CCF(i, k, 3);
        ci4 = wa3[i - 2] * CCF(i, k, 3) - wa3[i - 1] * CCF(i - 1, k, 3);
        tr1 = cr2 + cr4;
        tr4 = cr4 - cr2;
        ti1 = ci2 + ci4;
        ti4 = ci2 - ci4;
        ti2 = CCF(i, k, 0) + ci3;
        ti3 = CCF(i, k, 0) - ci3;
        tr2 = CCF(i - 1, k, 0) + cr3;
        tr3 = CCF(i - 1, k, 0) - cr3;
        CHF(i - 1, 0, k) = tr1 + tr2;
        CHF(ic - 1, 3, k) = tr2 - tr1;
        CHF(i, 0, k) = ti1 + ti2;
        CHF(ic, 3, k) = ti1 - ti2;
        CHF(i - 1, 2, k) = ti4 + tr3;
        CHF(ic - 1, 1, k) = tr3 - ti4;
        CHF(i, 2, k) = tr4 + ti3;
        CHF(ic, 1, k) = tr4 - ti3;
      }
    }
    if (ido % 2 == 1) return;
  }
  for (k = 0; k < l1; k++) {
    ti1 = (-hsqt2) * (CCF(ido - 1, k, 1) + CCF(ido - 1, k, 3));
    tr1 = hsqt2 * (CCF(ido - 1, k, 1) - CCF(ido - 1, k, 3));
    CHF(ido - 1, 0, k) = tr1 + CCF(ido - 1, k, 0);
    CHF(ido - 1, 2, k) = CCF(ido - 1, k, 0) - tr1;
    CHF(0, 1, k) = ti1 - CCF(ido - 1, k, 2);
    CHF(0, 3, k) = ti1 + CCF(ido - 1, k, 2);
  }
}

/* r1f5kf.f:31-143 */
static void radf5(int ido, int l1, const real *cc, real *ch, const real *wa1, const real *wa2,
                  const real *wa3, const real *wa4) {
  const int ip = 5;
  const real arg = R_LIT(2.0) * R_LIT(4.0) * R_ATAN(R_LIT(1.0)) / R_LIT(5.0);
  const real tr11 = R_COS(arg), ti11 = R_SIN(arg);
  const real tr12 = R_COS(R_LIT(2.0) * arg), ti12 = R_SIN(R_LIT(2.0) * arg);
  int i, k, ic;
  real cr2, ci2, cr3, ci3, cr4, ci4, cr5, ci5, dr2, di2, dr3, di3, dr4, di4, dr5, di5;
  real tr2, ti2, tr3, ti3, tr4, ti4, tr5, ti5;
  for (k = 0; k < l1; k++) {
    cr2 = CCF(0, k, 4) + CCF(0, k, 1);
    ci5 = CCF(0, k, 4) - CCF(0, k, 1);
    cr3 = CCF(0, k, 3) + CCF(0, k, 2);
    ci4 = CCF(0, k, 3) - CCF(0, k, 2);
    CHF(0, 0, k) = CCF(0, k, 0) + cr2 + cr3;
    CHF(ido - 1, 1, k) = CCF(0, k, 0) + tr11 * cr2 + tr12 * cr3;
    CHF(0, 2, k) = ti11 * ci5 + ti12 * ci4;
    CHF(ido - 1, 3, k) = CCF(0, k, 0) + tr12 * cr2 + tr11 * cr3;
    CHF(0, 4, k) = ti12 * ci5 - ti11 * ci4;
  }
  if (ido == 1) return;
  for (k = 0; k < l1; k++) {
    for (i = 2; i < ido; i += 2) {
      ic = ido - i;
      dr2 = wa1[i - 2] * CCF(i - 1, k, 1) + wa1[i - 1] * CCF(i, k, 1);
      di2 = wa1[i - 2] * CCF(i, k, 1) - wa1[i - 1] * CCF(i - 1, k, 1);
      dr3 = wa2[i - 2] * CCF(i - 1, k, 2) + wa2[i - 1] * CCF(i, k, 2);
      di3 = wa2[i - 2] * CCF(i, k, 2) - wa2[i - 1] * CCF(i - 1, k, 2);
      dr4 = wa3[i - 2] * CCF(i - 1, k, 3) + wa3[i - 1] * CCF(i, k, 3);
      di4 = wa3[i - 2] * CCF(i, k, 3) - wa3[i - 1] * CCF(i - 1, k, 3);
      dr5 = wa4[i - 2] * CCF(i - 1, k, 4) + wa4[i - 1] * CCF(i, k, 4);
      di5 = wa4[i - 2] * CCF(i, k, 4) - wa4[i - 1] * CCF(i - 1, k, 4);
      cr2 = dr2 + dr5;
      ci5 = dr5 - dr2;
      cr5 = di2 - di5;
      ci2 = di2 + di5;
      cr3 = dr3 + dr4;
      ci4 = dr4 - dr3;
      cr4 = di3 - di4;
      ci3 = di3 + di4;
      CHF(i - 1, 0, k) = CCF(i - 1, k, 0) + cr2 + cr3;
      CHF(i, 0, k) = CCF(i, k, 0) + ci2 + ci3;
      tr2 = CCF(i - 1, k, 0) + tr11 * cr2 + tr12 * cr3;
      ti2 = CCF(i, k, 0) + tr11 * ci2 + tr12 * ci3;
      tr3 = CCF(i - 1, k, 0) + tr12 * cr2 + tr11 * cr3;
      ti3 = CCF(i, k, 0) + tr12 * ci2 + tr11 * ci3;
      tr5 = ti11 * cr5 + ti12 * cr4;
      ti5 = ti11 * ci5 + ti12 * ci4;
      tr4 = ti12 * cr5 - ti11 * cr4;
      ti4 = ti12 * ci5 - ti11 * ci4;
      CHF(i - 1, 2, k) = tr2 + tr5;
      CHF(ic - 1, 1, k) = tr2 - tr5;
      CHF(i, 2, k) = ti2 + ti5;
      CHF(ic, 1, k) = ti5 - ti2;
      CHF(i - 1, 4, k) = tr3 + tr4;
      CHF(ic - 1, 3, k) = tr3 - tr4;
      CHF(i, 4, k) = ti3 + ti4;
      CHF(ic, 3, k) = ti4 - ti3;
    }
  }
}

/* r1f2kb.f:31-62 */
static void radb2(int ido, int l1, const real *cc, real *ch, const real *wa1) {
  const int ip = 2;
  int i, k, ic;
  real tr2, ti2;
  for (k = 0; k < l1; k++) {
    CHB(0, k, 0) = CCB(0, 0, k) + CCB(ido - 1, 1, k);
    CHB(0, k, 1) = CCB(0, 0, k) - CCB(ido - 1, 1, k);
  }
  if (ido < 2) return;
  if (ido > 2) {
    for (k = 0; k < l1; k++) {
      for (i = 2; i < ido; i += 2) {
        ic = ido - i;
        CHB(i - 1, k, 0) = CCB(i - 1, 0, k) + CCB(ic - 1, 1, k);
        CHB(i, k, 0) = CCB(i, 0, k) - CCB(ic, 1, k);
        tr2 = CCB(i - 1, 0, k) - CCB(ic - 1, 1, k);
        ti2 = CCB(i, 0, k) + CCB(ic, 1, k);
        CHB(i - 1, k, 1) = wa1[i - 2] * tr2 - wa1[i - 1] * ti2;
        CHB(i, k, 1) = wa1[i - 2] * ti2 + wa1[i - 1] * tr2;
      }
    }
    if (ido % 2 == 1) return;
  }
  for (k = 0; k < l1; k++) {
    CHB(ido - 1, k, 0) = CCB(ido - 1, 0, k) + CCB(ido - 1, 0, k);
    CHB(ido - 1, k, 1) = -(CCB(0, 1, k) + CCB(0, 1, k));
  }
}

/* r1f3kb.f:31-87 */
static void radb3(int ido, int l1, const real *cc, real *ch, const real *wa1, const real *wa2) {
  const int ip = 3;
  const real arg = R_LIT(2.0) * R_LIT(4.0) * R_ATAN(R_LIT(1.0)) / R_LIT(3.0);
  const real taur = R_COS(arg), taui = R_SIN(arg);
  int i, k, ic;
  real tr2, ti2, cr2, ci2, cr3, ci3, dr2, di2, dr3, di3;
  for (k = 0; k < l1; k++) {
    CHB(0, k, 0) = CCB(0, 0, k) + R_LIT(2.0) * CCB(ido - 1, 1, k);
    CHB(0, k, 1) = CCB(0, 0, k) + (R_LIT(2.0) * taur) * CCB(ido - 1, 1, k) -
                   (R_LIT(2.0) * taui) * CCB(0, 2, k);
    CHB(0, k, 2) = CCB(0, 0, k) + (R_LIT(2.0) * taur) * CCB(ido - 1, 1, k) +
                   R_LIT(2.0) * taui * CCB(0, 2, k);
  }
  if (ido == 1) return;
  for (k = 0; k < l1; k++) {
    for (i = 2; i < ido; i += 2) {
      ic = ido - i;
      tr2 = CCB(i - 1, 2, k) + CCB(ic - 1, 1, k);
      ti2 = CCB(i, 2, k) - CCB(ic, 1, k);
      CHB(i - 1, k, 0) = CCB(i - 1, 0, k) + tr2;
      CHB(i, k, 0) = CCB(i, 0, k) + ti2;
      cr2 = CCB(i - 1, 0, k) + taur * tr2;
      ci2 = CCB(i, 0, k) + taur * ti2;
      cr3 = taui * (CCB(i - 1, 2, k) - CCB(ic - 1, 1, k));
      ci3 = taui * (CCB(i, 2, k) + CCB(ic, 1, k));
      dr2 = cr2 - ci3;
      dr3 = cr2 + ci3;
      di2 = ci2 + cr3;
      di3 = ci2 - cr3;
      CHB(i - 1, k, 1) = wa1[i - 2] * dr2 - wa1[i - 1] * di2;
      CHB(i, k, 1) = wa1[i - 2] * di2 + wa1[i - 1] * dr2;
      CHB(i - 1, k, 2) = wa2[i - 2] * dr3 - wa2[i - 1] * di3;
      CHB(i, k, 2) = wa2[i - 2] * di3 + wa2[i - 1] * dr3;
    }
  }
}

/* r1f4kb.f:31-92 */
static void radb4(int ido, int l1, const real *cc, real *ch, const real *wa1, const real *wa2,
                  const real *wa3) {
  const int ip = 4;
  const real sqrt2 = R_SQRT(R_LIT(2.0));
  int i, k, ic;
  real tr1, ti1, tr2, ti2, tr3, ti3, tr4, ti4, cr2, ci2, cr3, ci3, cr4, ci4;
  for (k = 0; k < l1; k++) {
    tr3 = CCB(ido - 1, 1, k) + CCB(ido - 1, 1, k);
    tr2 = CCB(0, 0, k) + CCB(ido - 1, 3, k);
    tr1 = CCB(0, 0, k) - CCB(ido - 1, 3, k);
    tr4 = CCB(0, 2, k) + CCB(0, 2, k);
    CHB(0, k, 2) = tr2 - tr3;
    CHB(0, k, 0) = tr2 + tr3;
    CHB(0, k, 3) = tr1 + tr4;
    CHB(0, k, 1) = tr1 - tr4;
  }
  if (ido < 2) return;
  if (ido > 2) {
    for (k = 0; k < l1; k++) {
      for (i = 2; i < ido; i += 2) {
        ic = ido - i;
        tr2 = CCB(i - 1, 0, k) + CCB(ic - 1, 3, k);
        tr1 = CCB(i - 1, 0, k) - CCB(ic - 1, 3, k);
        tr3 = CCB(i - 1, 2, k) + CCB(ic - 1, 1, k);
        ti4 = CCB(i - 1, 2, k) - CCB(ic - 1, 1, k);
        ti2 = CCB(i, 0, k) - CCB(ic, 3, k);
        ti1 = CCB(i, 0, k) + CCB(ic, 3, k);
        ti3 = CCB(i, 2, k) - CCB(ic, 1, k);
        tr4 = CCB(i, 2, k) + CCB(ic, 1, k);
        CHB(i - 1, k, 0) = tr2 + tr3;
        CHB(i, k, 0) = ti2 + ti3;
        cr2 = tr1 - tr4;
        ci2 = ti1 + ti4;
        cr3 = tr2 - tr3;
        ci3 = ti2 - ti3;
        cr4 = tr1 + tr4;
        ci4 = ti1 - ti4;
        CHB(i - 1, k, 1) = wa1[i - 2] * cr2 - wa1[i - 1] * ci2;
        CHB(i, k, 1) = wa1[i - 2] * ci2 + wa1[i - 1] * cr2;
        CHB(i - 1, k, 2) = wa2[i - 2] * cr3 - wa2[i - 1] * ci3;
        CHB(i, k, 2) = wa2[i - 2] * ci3 + wa2[i - 1] * cr3;
        CHB(i - 1, k, 3) = wa3[i - 2] * cr4 - wa3[i - 1] * ci4;
        CHB(i, k, 3) = wa3[i - 2] * ci4 + wa3[i - 1] * cr4;
      }
    }
    if (ido % 2 == 1) return;
  }
  for (k = 0; k < l1; k++) {
    tr2 = CCB(ido - 1, 0, k) + CCB(ido - 1, 2, k);
    tr1 = CCB(ido - 1, 0, k) - CCB(ido - 1, 2, k);
    ti1 = CCB(0, 1, k) + CCB(0, 3, k);
    ti2 = CCB(0, 3, k) - CCB(0, 1, k);
    CHB(ido - 1, k, 0) = tr2 + tr2;
    CHB(ido - 1, k, 1) = sqrt2 * (tr1 - ti1);
    CHB(ido - 1, k, 2) = ti2 + ti2;
    CHB(ido - 1, k, 3) = (-sqrt2) * (tr1 + ti1);
  }
}

/* r1f5kb.f:31-145 */
static void radb5(int ido, int l1, const real *cc, real *ch, const real *wa1, const real *wa2,
                  const real *wa3, const real *wa4) {
  const int ip = 5;
  const real arg = R_LIT(2.0) * R_LIT(4.0) * R_ATAN(R_LIT(1.0)) / R_LIT(5.0);
  const real tr11 = R_COS(arg), ti11 = R_SIN(arg);
  const real tr12 = R_COS(R_LIT(2.0) * arg), ti12 = R_SIN(R_LIT(2.0) * arg);
  int i, k, ic;
  real ti5, ti4, tr2, tr3, cr2, cr3, ci5, ci4, ti2, ti3, tr5, tr4, ci2, ci3, cr5, cr4;
  real dr2, di2, dr3, di3, dr4, di4, dr5, di5;
  for (k = 0; k < l1; k++) {
    ti5 = R_LIT(2.0) * CCB(0, 2, k);
    ti4 = R_LIT(2.0) * CCB(0, 4, k);
    tr2 = R_LIT(2.0) * CCB(ido - 1, 1, k);
    tr3 = R_LIT(2.0) * CCB(ido - 1, 3, k);
    CHB(0, k, 0) = CCB(0, 0, k) + tr2 + tr3;
    cr2 = CCB(0, 0, k) + tr11 * tr2 + tr12 * tr3;
    cr3 = CCB(0, 0, k) + tr12 * tr2 + tr11 * tr3;
    ci5 = ti11 * ti5 + ti12 * ti4;
    ci4 = ti12 * ti5 - ti11 * ti4;
    CHB(0, k, 1) = cr2 - ci5;
    CHB(0, k, 2) = cr3 - ci4;
    CHB(0, k, 3) = cr3 + ci4;
    CHB(0, k, 4) = cr2 + ci5;
  }
  if (ido == 1) return;
  for (k = 0; k < l1; k++) {
    for (i = 2; i < ido; i += 2) {
      ic = ido - i;
      ti5 = CCB(i, 2, k) + CCB(ic, 1, k);
      ti2 = CCB(i, 2, k) - CCB(ic, 1, k);
      ti4 = CCB(i, 4, k) + CCB(ic, 3, k);
      ti3 = CCB(i, 4, k) - CCB(ic, 3, k);
      tr5 = CCB(i - 1, 2, k) - CCB(ic - 1, 1, k);
      tr2 = CCB(i - 1, 2, k) + CCB(ic - 1, 1, k);
      tr4 = CCB(i - 1, 4, k) - CCB(ic - 1, 3, k);
      tr3 = CCB(i - 1, 4, k) + CCB(ic - 1, 3, k);
      CHB(i - 1, k, 0) = CCB(i - 1, 0, k) + tr2 + tr3;
      CHB(i, k, 0) = CCB(i, 0, k) + ti2 + ti3;
      cr2 = CCB(i - 1, 0, k) + tr11 * tr2 + tr12 * tr3;
      ci2 = CCB(i, 0, k) + tr11 * ti2 + tr12 * ti3;
      cr3 = CCB(i - 1, 0, k) + tr12 * tr2 + tr11 * tr3;
      ci3 = CCB(i, 0, k) + tr12 * ti2 + tr11 * ti3;
      cr5 = ti11 * tr5 + ti12 * tr4;
      ci5 = ti11 * ti5 + ti12 * ti4;
      cr4 = ti12 * tr5 - ti11 * tr4;
      ci4 = ti12 * ti5 - ti11 * ti4;
      dr3 = cr3 - ci4;
      dr4 = cr3 + ci4;
      di3 = ci3 + cr4;
      di4 = ci3 - cr4;
      dr5 = cr2 + ci5;
      dr2 = cr2 - ci5;
      di5 = ci2 - cr5;
      di2 = ci2 + cr5;
      CHB(i - 1, k, 1) = wa1[i - 2] * dr2 - wa1[i - 1] * di2;
      CHB(i, k, 1) = wa1[i - 2] * di2 + wa1[i - 1] * dr2;
      CHB(i - 1, k, 2) = wa2[i - 2] * dr3 - wa2[i - 1] * di3;
      CHB(i, k, 2) = wa2[i - 2] * di3 + wa2[i - 1] * dr3;
      CHB(i - 1, k, 3) = wa3[i - 2] * dr4 - wa3[i - 1] * di4;
      CHB(i, k, 3) = wa3[i - 2] * di4 + wa3[i - 1] * dr4;
      CHB(i - 1, k, 4) = wa4[i - 2] * dr5 - wa4[i - 1] * di5;
      CHB(i, k, 4) = wa4[i - 2] * di5 + wa4[i - 1] * dr5;
    }
  }
}

/* ---- plan --------------------------------------------------------------------------------------- */

struct orc_rfft_plan {
  int n;
  int nf;
  int fac[32];
  real *wa; /* n twiddles, rffti1.f layout */
  real *ch; /* work array of n */
};

/* rffti1.f:31-90.  Returns 0, or 3 when n has a prime factor other than 2,3,5. */
int orc_rfft_plan_create(int n, orc_rfft_plan **out) {
  static const int ntryh[4] = {4, 2, 3, 5};
  orc_rfft_plan *p;
  int nl = n, nf = 0, j = 0, ntry = 0, i;
  *out = NULL;
  if (n < 1) return 1;
  p = (orc_rfft_plan *)calloc(1, sizeof(*p));
  p->n = n;
  /* factorisation, rffti1.f:38-58 */
  while (nl != 1) {
    ntry = (j < 4) ? ntryh[j] : ntry + 2;
    j++;
    while (nl % ntry == 0) {
      p->fac[nf++] = ntry;
      nl /= ntry;
      if (ntry == 2 && nf != 1) { /* move the factor 2 to the front */
        for (i = nf - 1; i >= 1; i--) p->fac[i] = p->fac[i - 1];
        p->fac[0] = 2;
      }
    }
    if (ntry > 5 && nl != 1) { free(p); return 3; }
  }
  for (i = 0; i < nf; i++)
    if (p->fac[i] > 5) { free(p); return 3; }
  p->nf = nf;
  p->wa = (real *)calloc((size_t)n + 1, sizeof(real));
  p->ch = (real *)calloc((size_t)n + 1, sizeof(real));
  /* twiddles, rffti1.f:61-88: angle arithmetic in binary128 (DOUBLE PRECISION under
     -fdefault-real-8), rounded to `real` once. */
  {
    const __float128 tpi = 8.0Q * atanq(1.0Q);
    const __float128 argh = tpi / (__float128)n;
    int is = 0, l1 = 1, k1;
    for (k1 = 0; k1 < nf - 1; k1++) {
      int ip = p->fac[k1], ld = 0, l2 = l1 * ip, ido = n / l2, jj;
      for (jj = 1; jj <= ip - 1; jj++) {
        __float128 argld, fi = 0.0Q;
        int ii, idx = is;
        ld += l1;
        argld = (__float128)ld * argh;
        for (ii = 3; ii <= ido; ii += 2) {
          __float128 a;
          idx += 2;
          fi += 1.0Q;
          a = fi * argld;
          p->wa[idx - 2] = (real)cosq(a);
          p->wa[idx - 1] = (real)sinq(a);
        }
        is += ido;
      }
      l1 = l2;
    }
  }
  *out = p;
  return 0;
}

void orc_rfft_plan_destroy(orc_rfft_plan *p) {
  if (!p) return;
  free(p->wa);
  free(p->ch);
  free(p);
}

int orc_rfft_plan_factors(const orc_rfft_plan *p, int *fac, int maxfac) {
  int i;
  for (i = 0; i < p->nf && i < maxfac; i++) fac[i] = p->fac[i];
  return p->nf;
}

/* rfftf1.f:31-112 (INC=1): in-place forward transform of c[0..n-1], 1/N normalised. */
void orc_rfft_forward(orc_rfft_plan *p, real *c) {
  const int n = p->n, nf = p->nf;
  real *ch = p->ch;
  const real *wa = p->wa;
  int na = 1, l2 = n, iw = n - 1, k1, j;
  if (n == 1) return;
  for (k1 = 1; k1 <= nf; k1++) {
    int kh = nf - k1, ip = p->fac[kh], l1 = l2 / ip, ido = n / l2;
    real *in, *outp;
    iw -= (ip - 1) * ido;
    na = 1 - na;
    in = (na == 0) ? c : ch;
    outp = (na == 0) ? ch : c;
    switch (ip) {
      case 4: radf4(ido, l1, in, outp, wa + iw, wa + iw + ido, wa + iw + 2 * ido); break;
      case 2: radf2(ido, l1, in, outp, wa + iw); break;
      case 3: radf3(ido, l1, in, outp, wa + iw, wa + iw + ido); break;
      default:
        radf5(ido, l1, in, outp, wa + iw, wa + iw + ido, wa + iw + 2 * ido, wa + iw + 3 * ido);
        break;
    }
    l2 = l1;
  }
  {
    /* normalisation, rfftf1.f:87-111 */
    const real sn = R_LIT(1.0) / (real)n, tsn = R_LIT(2.0) / (real)n, tsnm = -tsn;
    const int modn = n % 2, nl = modn ? n - 1 : n - 2;
    const real *src = (na == 0) ? ch : c;
    c[0] = sn * src[0];
    for (j = 1; j < nl; j += 2) { /* Fortran J=2,NL,2 */
      c[j] = tsn * src[j];
      c[j + 1] = tsnm * src[j + 1];
    }
    if (!modn) c[n - 1] = sn * src[n - 1];
  }
}

/* rfftb1.f:31-115 (INC=1): in-place backward transform of c[0..n-1], un-normalised. */
void orc_rfft_backward(orc_rfft_plan *p, real *c) {
  const int n = p->n, nf = p->nf;
  real *ch = p->ch;
  const real *wa = p->wa;
  const real half = R_LIT(0.5), halfm = -R_LIT(0.5);
  const int modn = n % 2, nl = modn ? n - 1 : n - 2;
  int na = 0, k1, j, l1 = 1, iw = 0;
  if (n == 1) return;
  for (k1 = 1; k1 <= nf; k1++) na = 1 - na; /* rfftb1.f:35-42 with every factor <= 5 */
  if (na != 0) {
    ch[0] = c[0];
    ch[n - 1] = c[n - 1];
    for (j = 1; j < nl; j += 2) {
      ch[j] = half * c[j];
      ch[j + 1] = halfm * c[j + 1];
    }
  } else {
    for (j = 1; j < nl; j += 2) {
      c[j] = half * c[j];
      c[j + 1] = halfm * c[j + 1];
    }
  }
  for (k1 = 1; k1 <= nf; k1++) {
    int ip = p->fac[k1 - 1], l2 = ip * l1, ido = n / l2;
    real *in = (na == 0) ? c : ch;
    real *outp = (na == 0) ? ch : c;
    switch (ip) {
      case 4: radb4(ido, l1, in, outp, wa + iw, wa + iw + ido, wa + iw + 2 * ido); break;
      case 2: radb2(ido, l1, in, outp, wa + iw); break;
      case 3: radb3(ido, l1, in, outp, wa + iw, wa + iw + ido); break;
      default:
        radb5(ido, l1, in, outp, wa + iw, wa + iw + ido, wa + iw + 2 * ido, wa + iw + 3 * ido);
        break;
    }
    na = 1 - na;
    l1 = l2;
    iw += (ip - 1) * ido;
  }
}

/* binary64 entry points for the test-suite (ctypes) */
int orc_rfft_forward_f64(int n, double *x) {
  orc_rfft_plan *p;
  real *t;
  int i, ier = orc_rfft_plan_create(n, &p);
  if (ier) return ier;
  t = (real *)malloc(sizeof(real) * (size_t)n);
  for (i = 0; i < n; i++) t[i] = (real)x[i];
  orc_rfft_forward(p, t);
  for (i = 0; i < n; i++) x[i] = (double)t[i];
  free(t);
  orc_rfft_plan_destroy(p);
  return 0;
}

int orc_rfft_backward_f64(int n, double *x) {
  orc_rfft_plan *p;
  real *t;
  int i, ier = orc_rfft_plan_create(n, &p);
  if (ier) return ier;
  t = (real *)malloc(sizeof(real) * (size_t)n);
  for (i = 0; i < n; i++) t[i] = (real)x[i];
  orc_rfft_backward(p, t);
  for (i = 0; i < n; i++) x[i] = (double)t[i];
  free(t);
  orc_rfft_plan_destroy(p);
  return 0;
}

int orc_rfft_factors(int n, int *fac, int maxfac) {
  orc_rfft_plan *p;
  int nf, ier = orc_rfft_plan_create(n, &p);
  if (ier) return -ier;
  nf = orc_rfft_plan_factors(p, fac, maxfac);
  orc_rfft_plan_destroy(p);
  return nf;
}

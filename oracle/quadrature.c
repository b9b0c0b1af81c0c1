/*
 * oracle/quadrature.c -- adaptive 21-point Gauss-Kronrod quadrature for the Galewsky-jet initial
 * condition (jet_zonal_flow_test_mod.F90:63 calls QUADPACK `qags`, lib/quadpack.f90:1877, with
 * epsabs=1e-10, epsrel=1e-3).
 *
 * TEST INFRASTRUCTURE ONLY (see orc_real.h).
 *
 * Restated: the 21-point rule and its error heuristic (lib/quadpack.f90:6763-6945, `qk21`), the
 * bisect-the-worst-interval loop and the stopping test errsum <= max(epsabs, epsrel*|result|) of
 * `qagse`.  NOT restated: the epsilon-algorithm extrapolation (`qelg`) that distinguishes QAGS from
 * QAG.  The integrand here is C-infinity with compact support, extrapolation never helps it, and the
 * reference only asks for 1e-3 relative accuracy; tests/test_oracle_ic.py compares this routine with
 * SciPy's QUADPACK QAGS (the same algorithm the reference links) on the committed golden profile.
 */
#include "orc_internal.h"
#include <float.h>
#include <stdlib.h>

static const double xgk[11] = {
    0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
    0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
    0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
    0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
    0.294392862701460198131126603103866, 0.148874338981631210884826001129720,
    0.000000000000000000000000000000000};
static const double wgk[11] = {
    0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
    0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
    0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
    0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
    0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
    0.149445554002916905664936468389821};
static const double wg[5] = {
    0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
    0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
    0.295524224714752870173815619188769};

/* lib/quadpack.f90:6763-6945 */
static void qk21(orc_integrand f, void *ctx, double a, double b, double *result, double *abserr,
                 double *resabs, double *resasc) {
  double fv1[10], fv2[10];
  const double centr = 0.5 * (a + b), hlgth = 0.5 * (b - a), dhlgth = fabs(hlgth);
  double resg = 0.0, fc = f(centr, ctx), resk = wgk[10] * fc, rabs = fabs(resk), reskh, rasc;
  int j;
  for (j = 0; j < 5; j++) {
    const int jtw = 2 * j + 1;
    const double absc = hlgth * xgk[jtw];
    const double f1 = f(centr - absc, ctx), f2 = f(centr + absc, ctx), fsum = f1 + f2;
    fv1[jtw] = f1;
    fv2[jtw] = f2;
    resg += wg[j] * fsum;
    resk += wgk[jtw] * fsum;
    rabs += wgk[jtw] * (fabs(f1) + fabs(f2));
  }
  for (j = 0; j < 5; j++) {
    const int jtwm1 = 2 * j;
    const double absc = hlgth * xgk[jtwm1];
    const double f1 = f(centr - absc, ctx), f2 = f(centr + absc, ctx), fsum = f1 + f2;
    fv1[jtwm1] = f1;
    fv2[jtwm1] = f2;
    resk += wgk[jtwm1] * fsum;
    rabs += wgk[jtwm1] * (fabs(f1) + fabs(f2));
  }
  reskh = resk * 0.5;
  rasc = wgk[10] * fabs(fc - reskh);
  for (j = 0; j < 10; j++) rasc += wgk[j] * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
  *result = resk * hlgth;
  rabs *= dhlgth;
  rasc *= dhlgth;
  *abserr = fabs((resk - resg) * hlgth);
  if (rasc != 0.0 && *abserr != 0.0) {
    double t = pow(200.0 * *abserr / rasc, 1.5);
    *abserr = rasc * (t < 1.0 ? t : 1.0);
  }
  if (rabs > DBL_MIN / (50.0 * DBL_EPSILON)) {
    double t = (DBL_EPSILON * 50.0) * rabs;
    if (t > *abserr) *abserr = t;
  }
  *resabs = rabs;
  *resasc = rasc;
}

int orc_qag21(orc_integrand f, void *ctx, double a, double b, double epsabs, double epsrel,
              double *result, double *abserr, int *neval) {
  enum { LIMIT = 500 };
  double alist[LIMIT], blist[LIMIT], rlist[LIMIT], elist[LIMIT];
  double resabs, resasc, errsum, res;
  int n = 1, i, ier = 0;
  qk21(f, ctx, a, b, &rlist[0], &elist[0], &resabs, &resasc);
  alist[0] = a;
  blist[0] = b;
  res = rlist[0];
  errsum = elist[0];
  *neval = 21;
  while (errsum > fmax(epsabs, epsrel * fabs(res))) {
    int worst = 0;
    double mid, r1, e1, r2, e2, d1, d2;
    if (n >= LIMIT) { ier = 1; break; }
    for (i = 1; i < n; i++)
      if (elist[i] > elist[worst]) worst = i;
    mid = 0.5 * (alist[worst] + blist[worst]);
    qk21(f, ctx, alist[worst], mid, &r1, &e1, &d1, &d2);
    qk21(f, ctx, mid, blist[worst], &r2, &e2, &d1, &d2);
    *neval += 42;
    alist[n] = mid;
    blist[n] = blist[worst];
    rlist[n] = r2;
    elist[n] = e2;
    blist[worst] = mid;
    rlist[worst] = r1;
    elist[worst] = e1;
    n++;
    res = 0.0;
    errsum = 0.0;
    for (i = 0; i < n; i++) {
      res += rlist[i];
      errsum += elist[i];
    }
  }
  *result = res;
  *abserr = errsum;
  return ier;
}

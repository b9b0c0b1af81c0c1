/* oracle/orc_internal.h -- private prototypes shared by the oracle's translation units.
 * TEST INFRASTRUCTURE ONLY (see orc_real.h). */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H
#include "orc_real.h"
#include "gmd_oracle.h"

void orc_rfft_forward(orc_rfft_plan *p, real *c);
void orc_rfft_backward(orc_rfft_plan *p, real *c);

/* adaptive 21-point Gauss-Kronrod quadrature (quadrature.c) */
typedef double (*orc_integrand)(double x, void *ctx);
int orc_qag21(orc_integrand f, void *ctx, double a, double b, double epsabs, double epsrel,
              double *result, double *abserr, int *neval);
#endif

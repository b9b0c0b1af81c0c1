/*
 * oracle/gmd_oracle.c -- CPU oracle: a sweep-by-sweep restatement of the barotropic shallow-water
 * time step of dongli/gamil-dycore (reference commit 5260294).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing here is part of the product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load this library.
 *
 * PARITY UNPINNED: the reference holds no golden vector or fixture for this path and cannot be
 * built in this image (see gmd_oracle.h).
 *
 * The code keeps the reference's array shapes (2-wide halos, (nlon+4) x (nlat+4), lon fastest), its
 * 1-based indices, its one-sweep-per-term loop structure and plain left-to-right sums, so that each
 * function can be audited against the Fortran it follows:
 *
 *   mesh_init                      src/mesh_mod.F90:44-114
 *   data_init (coefficients)       src/data_mod.F90:26-47
 *   fill_halo                      src/parallel_mod.F90:466-524   (periodic lon halo, lat halos stay 0)
 *   filter_init / filter_row       src/filter_mod.F90:35-167
 *   reset_cos_lat_at_poles         src/dycore_mod.F90:159-173
 *   iap_transform                  src/types_mod.F90:399-426
 *   inner_product_*                src/types_mod.F90:347-397
 *   space_operators + 7 operators  src/dycore_mod.F90:184-598
 *   update_state                   src/dycore_mod.F90:600-652
 *   time_integrate / csp2 / isp    src/dycore_mod.F90:654-752
 *   predict_correct                src/dycore_mod.F90:754-792
 *   check_antisymmetry             src/dycore_mod.F90:794-851
 *   ordinary_diffusion             src/diffusion_mod.F90:74-217
 *   weno_*                         src/weno_mod.F90:36-300
 *   diag_run / diag_total_energy   src/diag_mod.F90:42-121
 *   test-case plugins              src/test_cases/barotropic/{rossby_haurwitz_wave,steady_geostrophic_flow,
 *                                  mountain_zonal_flow,jet_zonal_flow}_test_mod.F90
 *
 * Deliberate deviations: B4 half_cos_lat(0), half_cos_lat(nlat) are defined as 0
 * (the reference reads them out of bounds and multiplies by a zero ghost v); B2's out-of-bounds
 * flag write is dropped; isp's tendency algebra runs on du,dv,dgd only (the other ten arrays of
 * types_mod.F90:229-345 are never read afterwards).
 */
#include "orc_internal.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- constants, params_mod.F90:6-11 ------------------------------------------------------------- */
#define PI_ (R_ATAN(R_LIT(1.0)) * R_LIT(4.0))
#define OMEGA_ (R_LIT(2.0) * PI_ / R_LIT(86400.0))
#define RADIUS_ R_LIT(6.37122e6)
#define G_ R_LIT(9.80616)

static char g_err[512];
const char *orc_last_error(void) { return g_err; }
int orc_real_bytes(void) { return (int)sizeof(real); }

typedef struct { real *u, *v, *gd, *iu, *iv, *igd; } state_t;
typedef struct {
  real *u_adv_lon, *u_adv_lat, *v_adv_lon, *v_adv_lat, *fu, *fv, *u_pgf, *v_pgf, *mass_div_lon,
      *mass_div_lat, *du, *dv, *dgd;
} tend_t;

struct orc_model {
  orc_config cfg;
  int nlon, nlat, LD;
  size_t NE;
  /* mesh (1-based 1-D tables; half_cos_lat has valid padded entries 0 and nlat) */
  real dlon, dlat;
  real *full_lon, *half_lon, *full_lat, *half_lat;
  real *full_cos_lon, *half_cos_lon, *full_sin_lon, *half_sin_lon;
  real *full_cos_lat, *half_cos_lat, *full_sin_lat, *half_sin_lat;
  /* coef */
  real *full_f, *half_f, *full_c, *half_c, *full_dlon, *half_dlon, *full_dlat, *half_dlat;
  state_t state_[4]; /* state(-1:2) */
  tend_t tend_[5];   /* tend(-2:2)  */
  real *ghs;
  /* filter */
  int *filter_full, *filter_half;   /* flags, 1-based */
  int *cutoff_full, *cutoff_half;   /* mask cutoff per row, -1 = all-zero mask */
  int *reduce_full, *reduce_half;   /* zonal reduction factor per row, 0 = none (specified extension) */
  real *reduce_tmp;
  orc_rfft_plan *plan;
  real *local_x;
  /* diffusion work arrays */
  real *dud, *dvd, *dgdd, *du_, *dv_, *dgd_;
  /* weno */
  real *fp_u, *fn_u, *f_u, *fp_v, *fn_v, *f_v;
  /* diag */
  real *vor, *div, total_mass, total_energy, beta;
  int old_idx, new_idx, step, run_inited, pole_reset;
};

#define STATE(m, k) (&(m)->state_[(k) + 1])
#define TEND(m, k) (&(m)->tend_[(k) + 2])
#define A2(a, i, j) (a)[(size_t)((j) + 1) * (size_t)m->LD + (size_t)((i) + 1)]

static real *alloc2(const orc_model *m) { return (real *)calloc(m->NE, sizeof(real)); }
static real *alloc1(int n) { return (real *)calloc((size_t)n + 3, sizeof(real)); }

/* parallel_fill_halo_1, src/parallel_mod.F90:466-524 (all_halo defaults to true: both sides, every row) */
static void fill_halo(const orc_model *m, real *f) {
  const int nlon = m->nlon;
  int j;
  for (j = -1; j <= m->nlat + 2; j++) {
    A2(f, -1, j) = A2(f, nlon - 1, j);
    A2(f, 0, j) = A2(f, nlon, j);
  }
  for (j = -1; j <= m->nlat + 2; j++) {
    A2(f, nlon + 1, j) = A2(f, 1, j);
    A2(f, nlon + 2, j) = A2(f, 2, j);
  }
}

/* ---- mesh_init, src/mesh_mod.F90:44-114 ----------------------------------------------------------- */
static void mesh_init(orc_model *m) {
  const int nlon = m->nlon, nlat = m->nlat, nhalf = nlat - 1;
  const real pi = PI_, rad_to_deg = R_LIT(180.0) / pi;
  int i, j;
  (void)rad_to_deg;
  m->full_lon = alloc1(nlon); m->half_lon = alloc1(nlon);
  m->full_lat = alloc1(nlat); m->half_lat = alloc1(nlat);
  m->full_cos_lon = alloc1(nlon); m->half_cos_lon = alloc1(nlon);
  m->full_sin_lon = alloc1(nlon); m->half_sin_lon = alloc1(nlon);
  m->full_cos_lat = alloc1(nlat); m->half_cos_lat = alloc1(nlat);
  m->full_sin_lat = alloc1(nlat); m->half_sin_lat = alloc1(nlat);
  m->dlon = 2 * pi / nlon;
  for (i = 1; i <= nlon; i++) {
    m->full_lon[i] = (i - 1) * m->dlon;
    m->half_lon[i] = m->full_lon[i] + R_LIT(0.5) * m->dlon;
  }
  m->dlat = pi / nhalf;
  for (j = 1; j <= nhalf; j++) {
    m->full_lat[j] = -R_LIT(0.5) * pi + (j - 1) * m->dlat;
    m->half_lat[j] = m->full_lat[j] + R_LIT(0.5) * m->dlat;
  }
  m->full_lat[nlat] = R_LIT(0.5) * pi;
  for (i = 1; i <= nlon; i++) {
    m->full_cos_lon[i] = R_COS(m->full_lon[i]);
    m->full_sin_lon[i] = R_SIN(m->full_lon[i]);
    m->half_cos_lon[i] = R_COS(m->half_lon[i]);
    m->half_sin_lon[i] = R_SIN(m->half_lon[i]);
  }
  for (j = 1; j <= nhalf; j++) {
    m->half_cos_lat[j] = R_COS(m->half_lat[j]);
    m->half_sin_lat[j] = R_SIN(m->half_lat[j]);
  }
  /* B4: padded entries the reference reads out of bounds */
  m->half_cos_lat[0] = 0;
  m->half_cos_lat[nlat] = 0;
  for (j = 1; j <= nlat; j++) {
    m->full_cos_lat[j] = R_COS(m->full_lat[j]);
    m->full_sin_lat[j] = R_SIN(m->full_lat[j]);
  }
  m->full_cos_lat[1] = 0;
  m->full_cos_lat[nlat] = 0;
  m->full_sin_lat[1] = -1;
  m->full_sin_lat[nlat] = 1;
}

/* ---- data_init, src/data_mod.F90:26-47 ------------------------------------------------------------ */
static void alloc_state(orc_model *m, state_t *s) {
  s->u = alloc2(m); s->v = alloc2(m); s->gd = alloc2(m);
  s->iu = alloc2(m); s->iv = alloc2(m); s->igd = alloc2(m);
}
static void alloc_tend(orc_model *m, tend_t *t) {
  t->u_adv_lon = alloc2(m); t->u_adv_lat = alloc2(m); t->v_adv_lon = alloc2(m);
  t->v_adv_lat = alloc2(m); t->fu = alloc2(m); t->fv = alloc2(m); t->u_pgf = alloc2(m);
  t->v_pgf = alloc2(m); t->mass_div_lon = alloc2(m); t->mass_div_lat = alloc2(m);
  t->du = alloc2(m); t->dv = alloc2(m); t->dgd = alloc2(m);
}
static void data_init(orc_model *m) {
  const int nlat = m->nlat;
  const real omega = OMEGA_, radius = RADIUS_;
  int j, k;
  m->full_f = alloc1(nlat); m->half_f = alloc1(nlat); m->full_c = alloc1(nlat);
  m->half_c = alloc1(nlat); m->full_dlon = alloc1(nlat); m->half_dlon = alloc1(nlat);
  m->full_dlat = alloc1(nlat); m->half_dlat = alloc1(nlat);
  for (j = 1; j <= nlat; j++) {
    m->full_f[j] = R_LIT(2.0) * omega * m->full_sin_lat[j];
    if (j == 1 || j == nlat) m->full_c[j] = 0;
    else m->full_c[j] = m->full_sin_lat[j] / m->full_cos_lat[j] / radius;
    m->full_dlon[j] = radius * m->dlon * m->full_cos_lat[j];
    m->full_dlat[j] = radius * m->dlat * m->full_cos_lat[j];
  }
  for (j = 1; j <= nlat - 1; j++) {
    m->half_f[j] = R_LIT(2.0) * omega * m->half_sin_lat[j];
    m->half_c[j] = m->half_sin_lat[j] / m->half_cos_lat[j] / radius;
    m->half_dlon[j] = radius * m->dlon * m->half_cos_lat[j];
    m->half_dlat[j] = radius * m->dlat * m->half_cos_lat[j];
  }
  for (k = 0; k < 4; k++) alloc_state(m, &m->state_[k]);
  for (k = 0; k < 5; k++) alloc_tend(m, &m->tend_[k]);
  m->ghs = alloc2(m);
}

/* ---- filter_init, src/filter_mod.F90:35-103 ------------------------------------------------------- */
static int filter_init(orc_model *m) {
  const int nlat = m->nlat, nhalf = nlat - 1;
  const int *cw = m->cfg.zonal_tend_filter_cutoff_wavenumber;
  int j, ier;
  m->filter_full = (int *)calloc((size_t)nlat + 3, sizeof(int));
  m->filter_half = (int *)calloc((size_t)nlat + 3, sizeof(int));
  m->cutoff_full = (int *)malloc(((size_t)nlat + 3) * sizeof(int));
  m->cutoff_half = (int *)malloc(((size_t)nlat + 3) * sizeof(int));
  for (j = 0; j < nlat + 3; j++) m->cutoff_full[j] = m->cutoff_half[j] = -1;
  if (m->cfg.use_zonal_tend_filter) {
    for (j = 1; j <= 20; j++) {
      if (cw[j - 1] != 0) {
        /* south, filter_mod.F90:44-51 */
        if (1 + j <= nlat) m->filter_full[1 + j] = 1;
        if (j <= nhalf) m->filter_half[j] = 1;
        /* north, filter_mod.F90:52-59; nlat-j+1 == nlat for j=1 is out of bounds in the reference (B2) */
        if (nlat - j >= 1) m->filter_full[nlat - j] = 1;
        if (nlat - j + 1 >= 1 && nlat - j + 1 <= nhalf) m->filter_half[nlat - j + 1] = 1;
      }
    }
  }
  ier = orc_rfft_plan_create(m->nlon, &m->plan);
  if (ier) {
    snprintf(g_err, sizeof g_err, "Failed to initialize FFTPACK! (num_lon=%d has a factor other than 2,3,5)", m->nlon);
    return ier;
  }
  m->local_x = (real *)calloc((size_t)m->nlon, sizeof(real));
  /* masks, filter_mod.F90:76-99: entries 1..2(c+1) of the halfcomplex array are kept; a row hit by
     several entries gets the union, i.e. the largest cutoff */
  for (j = 1; j <= 20; j++) {
    int c = cw[j - 1];
    if (c != 0) {
      if (1 + j <= nlat && c > m->cutoff_full[1 + j]) m->cutoff_full[1 + j] = c;
      if (j <= nhalf && c > m->cutoff_half[j]) m->cutoff_half[j] = c;
      if (nlat - j >= 1 && c > m->cutoff_full[nlat - j]) m->cutoff_full[nlat - j] = c;
      if (nhalf - j + 1 >= 1 && c > m->cutoff_half[nhalf - j + 1]) m->cutoff_half[nhalf - j + 1] = c;
    }
  }
  /* moving reduced tendency (specified extension): factor of the k-th full row next to each pole (pole row excluded)
     and of the k-th half row from each pole; a row is either filtered or reduced */
  m->reduce_full = (int *)calloc((size_t)nlat + 3, sizeof(int));
  m->reduce_half = (int *)calloc((size_t)nlat + 3, sizeof(int));
  m->reduce_tmp = (real *)calloc((size_t)m->nlon + 3, sizeof(real));
  if (m->cfg.use_zonal_reduce) {
    for (j = 1; j <= 20; j++) {
      const int r = m->cfg.zonal_reduce_factors[j - 1];
      if (r < 0 || (r > 1 && m->nlon % r != 0)) {
        snprintf(g_err, sizeof g_err, "zonal_reduce_factors(%d)=%d must be >= 0 and divide num_lon=%d", j, r, m->nlon);
        return 2;
      }
      if (r <= 1) continue;
      if (1 + j <= nlat - 1 && r > m->reduce_full[1 + j]) m->reduce_full[1 + j] = r;
      if (nlat - j >= 2 && r > m->reduce_full[nlat - j]) m->reduce_full[nlat - j] = r;
      if (j <= nhalf && r > m->reduce_half[j]) m->reduce_half[j] = r;
      if (nhalf - j + 1 >= 1 && r > m->reduce_half[nhalf - j + 1]) m->reduce_half[nhalf - j + 1] = r;
    }
    for (j = 1; j <= nlat; j++)
      if ((m->reduce_full[j] > 1 && m->filter_full[j]) || (j <= nhalf && m->reduce_half[j] > 1 && m->filter_half[j])) {
        snprintf(g_err, sizeof g_err, "row %d is both a zonal filter row and a reduced row", j);
        return 2;
      }
  }
  return 0;
}

/* filter_array_at_full_lat / filter_array_at_half_lat, src/filter_mod.F90:105-167.
   x points at element (i=1) of the row. */
static void filter_row_real(orc_model *m, int cutoff, real *x) {
  const int n = m->nlon;
  int i, keep = (cutoff < 0) ? 0 : 2 * (cutoff + 1);
  if (keep > n) keep = n;
  for (i = 0; i < n; i++) m->local_x[i] = x[i];
  orc_rfft_forward(m->plan, m->local_x);
  for (i = 0; i < n; i++) m->local_x[i] = m->local_x[i] * (i < keep ? R_LIT(1.0) : R_LIT(0.0));
  orc_rfft_backward(m->plan, m->local_x);
  for (i = 0; i < n; i++) x[i] = m->local_x[i];
}
static void filter_array_at_full_lat(orc_model *m, int j, real *field) {
  filter_row_real(m, m->cutoff_full[j], &A2(field, 1, j));
}
static void filter_array_at_half_lat(orc_model *m, int j, real *field) {
  filter_row_real(m, m->cutoff_half[j], &A2(field, 1, j));
}

/* ---- reset_cos_lat_at_poles, src/dycore_mod.F90:159-173 ------------------------------------------- */
static void reset_cos_lat_at_poles(orc_model *m) {
  const real radius = RADIUS_;
  int j = 1;
  m->full_cos_lat[j] = m->half_cos_lat[1] * R_LIT(0.25);
  m->full_dlon[j] = radius * m->dlon * m->full_cos_lat[j];
  m->full_dlat[j] = radius * m->dlat * m->full_cos_lat[j];
  j = m->nlat;
  m->full_cos_lat[j] = m->half_cos_lat[m->nlat - 1] * R_LIT(0.25);
  m->full_dlon[j] = radius * m->dlon * m->full_cos_lat[j];
  m->full_dlat[j] = radius * m->dlat * m->full_cos_lat[j];
  m->pole_reset = 1;
}

/* ---- iap_transform, src/types_mod.F90:399-426 ----------------------------------------------------- */
static void iap_transform(orc_model *m, state_t *s) {
  const int nlon = m->nlon, nlat = m->nlat;
  int i, j;
  for (j = 1; j <= nlat; j++)
    for (i = 1; i <= nlon; i++) A2(s->igd, i, j) = R_SQRT(A2(s->gd, i, j));
  fill_halo(m, s->igd);
  for (j = 1; j <= nlat; j++)
    for (i = 1; i <= nlon; i++)
      A2(s->iu, i, j) = R_LIT(0.5) * (A2(s->igd, i, j) + A2(s->igd, i + 1, j)) * A2(s->u, i, j);
  fill_halo(m, s->iu);
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++)
      A2(s->iv, i, j) = R_LIT(0.5) * (A2(s->igd, i, j) + A2(s->igd, i, j + 1)) * A2(s->v, i, j);
  fill_halo(m, s->iv);
}

/* ---- operators, src/dycore_mod.F90:367-598 -------------------------------------------------------- */
static void weno_zonal(orc_model *m, const state_t *s, tend_t *t);
static void weno_meridional(orc_model *m, const state_t *s, tend_t *t);

/* src/dycore_mod.F90:367-420 */
static void zonal_momentum_advection_operator(orc_model *m, const state_t *s, tend_t *t) {
  const int nlon = m->nlon, nlat = m->nlat;
  const real beta = (real)m->cfg.uv_adv_upwind_lon_beta;
  real u1, u2;
  int i, j;
  switch (m->cfg.uv_adv_scheme) {
    case ORC_ADV_CENTER_DIFF:
      for (j = 2; j <= nlat - 1; j++)
        for (i = 1; i <= nlon; i++) {
          u1 = A2(s->u, i, j) + A2(s->u, i - 1, j);
          u2 = A2(s->u, i, j) + A2(s->u, i + 1, j);
          A2(t->u_adv_lon, i, j) =
              R_LIT(0.25) / m->full_dlon[j] * (u2 * A2(s->iu, i + 1, j) - u1 * A2(s->iu, i - 1, j));
        }
      for (j = 1; j <= nlat - 1; j++)
        for (i = 1; i <= nlon; i++) {
          u1 = A2(s->u, i - 1, j) + A2(s->u, i - 1, j + 1);
          u2 = A2(s->u, i, j) + A2(s->u, i, j + 1);
          A2(t->v_adv_lon, i, j) =
              R_LIT(0.25) / m->half_dlon[j] * (u2 * A2(s->iv, i + 1, j) - u1 * A2(s->iv, i - 1, j));
        }
      break;
    case ORC_ADV_UPWIND:
      for (j = 2; j <= nlat - 1; j++)
        for (i = 1; i <= nlon; i++) {
          u1 = A2(s->u, i, j) + A2(s->u, i - 1, j);
          u2 = A2(s->u, i, j) + A2(s->u, i + 1, j);
          A2(t->u_adv_lon, i, j) =
              R_LIT(0.25) / m->full_dlon[j] *
              (u2 * (A2(s->iu, i, j) + A2(s->iu, i + 1, j)) -
               beta * R_FABS(u2) * (A2(s->iu, i + 1, j) - A2(s->iu, i, j)) -
               u1 * (A2(s->iu, i, j) + A2(s->iu, i - 1, j)) +
               beta * R_FABS(u1) * (A2(s->iu, i, j) - A2(s->iu, i - 1, j)) -
               (u2 - u1) * A2(s->iu, i, j));
        }
      for (j = 1; j <= nlat - 1; j++)
        for (i = 1; i <= nlon; i++) {
          u1 = A2(s->u, i - 1, j) + A2(s->u, i - 1, j + 1);
          u2 = A2(s->u, i, j) + A2(s->u, i, j + 1);
          A2(t->v_adv_lon, i, j) =
              R_LIT(0.25) / m->half_dlon[j] *
              (u2 * (A2(s->iv, i, j) + A2(s->iv, i + 1, j)) -
               beta * R_FABS(u2) * (A2(s->iv, i + 1, j) - A2(s->iv, i, j)) -
               u1 * (A2(s->iv, i, j) + A2(s->iv, i - 1, j)) +
               beta * R_FABS(u1) * (A2(s->iv, i, j) - A2(s->iv, i - 1, j)) -
               (u2 - u1) * A2(s->iv, i, j));
        }
      break;
    default:
      weno_zonal(m, s, t);
  }
}

/* src/dycore_mod.F90:422-475 */
static void meridional_momentum_advection_operator(orc_model *m, const state_t *s, tend_t *t) {
  const int nlon = m->nlon, nlat = m->nlat;
  const real beta = (real)m->cfg.uv_adv_upwind_lat_beta;
  const real *hc = m->half_cos_lat;
  real v1, v2;
  int i, j;
  switch (m->cfg.uv_adv_scheme) {
    case ORC_ADV_CENTER_DIFF:
      for (j = 2; j <= nlat - 1; j++)
        for (i = 1; i <= nlon; i++) {
          v1 = (A2(s->v, i, j - 1) + A2(s->v, i + 1, j - 1)) * hc[j - 1];
          v2 = (A2(s->v, i, j) + A2(s->v, i + 1, j)) * hc[j];
          A2(t->u_adv_lat, i, j) =
              R_LIT(0.25) / m->full_dlat[j] * (v2 * A2(s->iu, i, j + 1) - v1 * A2(s->iu, i, j - 1));
        }
      for (j = 1; j <= nlat - 1; j++)
        for (i = 1; i <= nlon; i++) {
          v1 = A2(s->v, i, j) * hc[j] + A2(s->v, i, j - 1) * hc[j - 1];
          v2 = A2(s->v, i, j) * hc[j] + A2(s->v, i, j + 1) * hc[j + 1];
          A2(t->v_adv_lat, i, j) =
              R_LIT(0.25) / m->half_dlat[j] * (v2 * A2(s->iv, i, j + 1) - v1 * A2(s->iv, i, j - 1));
        }
      break;
    case ORC_ADV_UPWIND:
      for (j = 2; j <= nlat - 1; j++)
        for (i = 1; i <= nlon; i++) {
          v1 = (A2(s->v, i, j - 1) + A2(s->v, i + 1, j - 1)) * hc[j - 1];
          v2 = (A2(s->v, i, j) + A2(s->v, i + 1, j)) * hc[j];
          A2(t->u_adv_lat, i, j) =
              R_LIT(0.25) / m->full_dlat[j] *
              (v2 * (A2(s->iu, i, j) + A2(s->iu, i, j + 1)) -
               beta * R_FABS(v2) * (A2(s->iu, i, j + 1) - A2(s->iu, i, j)) -
               v1 * (A2(s->iu, i, j) + A2(s->iu, i, j - 1)) +
               beta * R_FABS(v1) * (A2(s->iu, i, j) - A2(s->iu, i, j - 1)) -
               (v2 - v1) * A2(s->iu, i, j));
        }
      for (j = 1; j <= nlat - 1; j++)
        for (i = 1; i <= nlon; i++) {
          v1 = A2(s->v, i, j) * hc[j] + A2(s->v, i, j - 1) * hc[j - 1];
          v2 = A2(s->v, i, j) * hc[j] + A2(s->v, i, j + 1) * hc[j + 1];
          A2(t->v_adv_lat, i, j) =
              R_LIT(0.25) / m->half_dlat[j] *
              (v2 * (A2(s->iv, i, j) + A2(s->iv, i, j + 1)) -
               beta * R_FABS(v2) * (A2(s->iv, i, j + 1) - A2(s->iv, i, j)) -
               v1 * (A2(s->iv, i, j) + A2(s->iv, i, j - 1)) +
               beta * R_FABS(v1) * (A2(s->iv, i, j) - A2(s->iv, i, j - 1)) -
               (v2 - v1) * A2(s->iv, i, j));
        }
      break;
    default:
      weno_meridional(m, s, t);
  }
}

/* src/dycore_mod.F90:477-505 */
static void coriolis_operator(orc_model *m, const state_t *s, tend_t *t) {
  const int nlon = m->nlon, nlat = m->nlat;
  const real *ff = m->full_f, *fc = m->full_c;
  real c1, c2;
  int i, j;
  for (j = 2; j <= nlat - 1; j++) {
    c1 = m->half_cos_lat[j - 1] / m->full_cos_lat[j];
    c2 = m->half_cos_lat[j] / m->full_cos_lat[j];
    for (i = 1; i <= nlon; i++)
      A2(t->fv, i, j) = R_LIT(0.25) * (ff[j] + fc[j] * A2(s->u, i, j)) *
                        (c1 * (A2(s->iv, i, j - 1) + A2(s->iv, i + 1, j - 1)) +
                         c2 * (A2(s->iv, i, j) + A2(s->iv, i + 1, j)));
  }
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++)
      A2(t->fu, i, j) =
          R_LIT(0.25) * ((ff[j] + fc[j] * A2(s->u, i, j)) * A2(s->iu, i, j) +
                         (ff[j] + fc[j] * A2(s->u, i - 1, j)) * A2(s->iu, i - 1, j) +
                         (ff[j + 1] + fc[j + 1] * A2(s->u, i, j + 1)) * A2(s->iu, i, j + 1) +
                         (ff[j + 1] + fc[j + 1] * A2(s->u, i - 1, j + 1)) * A2(s->iu, i - 1, j + 1));
}

/* src/dycore_mod.F90:507-537 */
static void zonal_pressure_gradient_force_operator(orc_model *m, const state_t *s, tend_t *t) {
  int i, j;
  for (j = 2; j <= m->nlat - 1; j++)
    for (i = 1; i <= m->nlon; i++)
      A2(t->u_pgf, i, j) =
          R_LIT(0.5) * (A2(s->igd, i, j) + A2(s->igd, i + 1, j)) / m->full_dlon[j] *
          (A2(s->gd, i + 1, j) + A2(m->ghs, i + 1, j) - A2(s->gd, i, j) - A2(m->ghs, i, j));
}
static void meridional_pressure_gradient_force_operator(orc_model *m, const state_t *s, tend_t *t) {
  int i, j;
  for (j = 1; j <= m->nlat - 1; j++)
    for (i = 1; i <= m->nlon; i++)
      A2(t->v_pgf, i, j) =
          R_LIT(0.5) * (A2(s->igd, i, j) + A2(s->igd, i, j + 1)) / m->half_dlat[j] *
          m->half_cos_lat[j] *
          (A2(s->gd, i, j + 1) + A2(m->ghs, i, j + 1) - A2(s->gd, i, j) - A2(m->ghs, i, j));
}

/* src/dycore_mod.F90:539-598 */
static void zonal_mass_divergence_operator(orc_model *m, const state_t *s, tend_t *t) {
  int i, j;
  for (j = 2; j <= m->nlat - 1; j++)
    for (i = 1; i <= m->nlon; i++)
      A2(t->mass_div_lon, i, j) =
          ((A2(s->igd, i, j) + A2(s->igd, i + 1, j)) * A2(s->iu, i, j) -
           (A2(s->igd, i, j) + A2(s->igd, i - 1, j)) * A2(s->iu, i - 1, j)) *
          R_LIT(0.5) / m->full_dlon[j];
}
static void meridional_mass_divergence_operator(orc_model *m, const state_t *s, tend_t *t) {
  const int nlon = m->nlon, nlat = m->nlat;
  const real radius = RADIUS_;
  real sp, np;
  int i, j;
  for (j = 2; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++)
      A2(t->mass_div_lat, i, j) =
          ((A2(s->igd, i, j) + A2(s->igd, i, j + 1)) * A2(s->iv, i, j) * m->half_cos_lat[j] -
           (A2(s->igd, i, j) + A2(s->igd, i, j - 1)) * A2(s->iv, i, j - 1) * m->half_cos_lat[j - 1]) *
          R_LIT(0.5) / m->full_dlat[j];
  j = 1;
  sp = 0;
  for (i = 1; i <= nlon; i++) sp = sp + (A2(s->igd, i, j) + A2(s->igd, i, j + 1)) * A2(s->iv, i, j);
  sp = sp * R_LIT(2.0) / nlon / radius / m->dlat;
  for (i = 1; i <= nlon; i++) A2(t->mass_div_lat, i, j) = sp;
  j = nlat;
  np = 0;
  for (i = 1; i <= nlon; i++)
    np = np - (A2(s->igd, i, j) + A2(s->igd, i, j - 1)) * A2(s->iv, i, j - 1);
  np = np * R_LIT(2.0) / nlon / radius / m->dlat;
  for (i = 1; i <= nlon; i++) A2(t->mass_div_lat, i, j) = np;
}

/* ---- weno, src/weno_mod.F90:36-300 (weno_order = 2, the only reachable order) ----------------------- */
static const real weno_eps = R_LIT(1.0e-6), weno_umax = R_LIT(20.0), weno_vmax = R_LIT(20.0);

/* src/weno_mod.F90:248-300 */
static real weno_2nd_order_pass(real fp1, real fp2, real fp3, real fn2, real fn3, real fn4) {
  const real c11 = -R_LIT(1.0) / R_LIT(2.0), c21 = R_LIT(3.0) / R_LIT(2.0);
  const real c12 = R_LIT(1.0) / R_LIT(2.0), c22 = R_LIT(1.0) / R_LIT(2.0);
  const real wo1 = R_LIT(1.0) / R_LIT(3.0), wo2 = R_LIT(2.0) / R_LIT(3.0);
  real fs1, fs2, b1, b2, w1, w2, sw, f;
  fs1 = c11 * fp1 + c21 * fp2;
  fs2 = c12 * fp2 + c22 * fp3;
  b1 = (fp2 - fp1) * (fp2 - fp1);
  b2 = (fp3 - fp2) * (fp3 - fp2);
  w1 = wo1 / ((weno_eps + b1) * (weno_eps + b1));
  w2 = wo2 / ((weno_eps + b2) * (weno_eps + b2));
  sw = w1 + w2;
  w1 = w1 / sw;
  w2 = w2 / sw;
  f = w1 * fs1 + w2 * fs2;
  fs1 = c11 * fn4 + c21 * fn3;
  fs2 = c12 * fn3 + c22 * fn2;
  b1 = (fn3 - fn4) * (fn3 - fn4);
  b2 = (fn2 - fn3) * (fn2 - fn3);
  w1 = wo1 / ((weno_eps + b1) * (weno_eps + b1));
  w2 = wo2 / ((weno_eps + b2) * (weno_eps + b2));
  sw = w1 + w2;
  w1 = w1 / sw;
  w2 = w2 / sw;
  f = f + (w1 * fs1 + w2 * fs2);
  return f;
}

static void weno_alloc(orc_model *m) {
  if (m->fp_u) return;
  m->fp_u = alloc2(m); m->fn_u = alloc2(m); m->f_u = alloc2(m);
  m->fp_v = alloc2(m); m->fn_v = alloc2(m); m->f_v = alloc2(m);
}

/* src/weno_mod.F90:69-164 */
static void weno_zonal(orc_model *m, const state_t *s, tend_t *t) {
  const int nlon = m->nlon, nlat = m->nlat;
  real u;
  int i, j;
  weno_alloc(m);
  for (j = 2; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) {
      A2(m->fp_u, i, j) = R_LIT(0.5) * (A2(s->u, i, j) + weno_umax) * A2(s->iu, i, j);
      A2(m->fn_u, i, j) = R_LIT(0.5) * (A2(s->u, i, j) - weno_umax) * A2(s->iu, i, j);
    }
  fill_halo(m, m->fp_u);
  fill_halo(m, m->fn_u);
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) {
      u = R_LIT(0.25) * (A2(s->u, i - 1, j) + A2(s->u, i - 1, j + 1) + A2(s->u, i, j) + A2(s->u, i, j + 1));
      A2(m->fp_v, i, j) = R_LIT(0.5) * (u + weno_umax) * A2(s->iv, i, j);
      A2(m->fn_v, i, j) = R_LIT(0.5) * (u - weno_umax) * A2(s->iv, i, j);
    }
  fill_halo(m, m->fp_v);
  fill_halo(m, m->fn_v);
  for (j = 2; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++)
      A2(m->f_u, i, j) = weno_2nd_order_pass(A2(m->fp_u, i - 1, j), A2(m->fp_u, i, j), A2(m->fp_u, i + 1, j),
                                             A2(m->fn_u, i, j), A2(m->fn_u, i + 1, j), A2(m->fn_u, i + 2, j));
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++)
      A2(m->f_v, i, j) = weno_2nd_order_pass(A2(m->fp_v, i - 1, j), A2(m->fp_v, i, j), A2(m->fp_v, i + 1, j),
                                             A2(m->fn_v, i, j), A2(m->fn_v, i + 1, j), A2(m->fn_v, i + 2, j));
  fill_halo(m, m->f_u);
  fill_halo(m, m->f_v);
  for (j = 2; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) /* B8: half_dlon(j) on a full row, weno_mod.F90:151-155 */
      A2(t->u_adv_lon, i, j) = (A2(m->f_u, i, j) - A2(m->f_u, i - 1, j) -
                                (A2(s->u, i + 1, j) - A2(s->u, i - 1, j)) * A2(s->iu, i, j) * R_LIT(0.25)) /
                               m->half_dlon[j];
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++)
      A2(t->v_adv_lon, i, j) =
          (A2(m->f_v, i, j) - A2(m->f_v, i - 1, j) -
           (A2(s->u, i, j) + A2(s->u, i, j + 1) - A2(s->u, i - 1, j) - A2(s->u, i - 1, j + 1)) *
               A2(s->iv, i, j) * R_LIT(0.25)) /
          m->half_dlon[j];
}

/* src/weno_mod.F90:166-233 */
static void weno_meridional(orc_model *m, const state_t *s, tend_t *t) {
  const int nlon = m->nlon, nlat = m->nlat;
  real v;
  int i, j;
  weno_alloc(m);
  for (j = 2; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) {
      v = R_LIT(0.25) * (A2(s->v, i, j - 1) + A2(s->v, i, j) + A2(s->v, i + 1, j - 1) + A2(s->v, i + 1, j));
      A2(m->fp_u, i, j) = R_LIT(0.5) * (v + weno_vmax) * A2(s->iu, i, j);
      A2(m->fn_u, i, j) = R_LIT(0.5) * (v - weno_vmax) * A2(s->iu, i, j);
    }
  fill_halo(m, m->fp_u);
  fill_halo(m, m->fn_u);
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) {
      A2(m->fp_v, i, j) = R_LIT(0.5) * (A2(s->v, i, j) + weno_vmax) * A2(s->iv, i, j);
      A2(m->fn_v, i, j) = R_LIT(0.5) * (A2(s->v, i, j) - weno_vmax) * A2(s->iv, i, j);
    }
  fill_halo(m, m->fp_v);
  fill_halo(m, m->fn_v);
  for (j = 2; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++)
      A2(m->f_u, i, j) = weno_2nd_order_pass(A2(m->fp_u, i, j - 1), A2(m->fp_u, i, j), A2(m->fp_u, i, j + 1),
                                             A2(m->fn_u, i, j), A2(m->fn_u, i, j + 1), A2(m->fn_u, i, j + 2));
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++)
      A2(m->f_v, i, j) = weno_2nd_order_pass(A2(m->fp_v, i, j - 1), A2(m->fp_v, i, j), A2(m->fp_v, i, j + 1),
                                             A2(m->fn_v, i, j), A2(m->fn_v, i, j + 1), A2(m->fn_v, i, j + 2));
  fill_halo(m, m->f_u);
  fill_halo(m, m->f_v);
  for (j = 2; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++)
      A2(t->u_adv_lat, i, j) =
          (A2(m->f_u, i, j) - A2(m->f_u, i, j - 1) -
           (A2(s->v, i - 1, j) + A2(s->v, i, j) - A2(s->v, i - 1, j - 1) - A2(s->v, i, j - 1)) *
               A2(s->iu, i, j) * R_LIT(0.25)) /
          m->full_dlat[j];
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++)
      A2(t->v_adv_lat, i, j) = (A2(m->f_v, i, j) - A2(m->f_v, i, j - 1) -
                                (A2(s->v, i, j + 1) - A2(s->v, i, j - 1)) * A2(s->iv, i, j) * R_LIT(0.25)) /
                               m->half_dlat[j];
}

/* ---- space_operators, src/dycore_mod.F90:184-365 ------------------------------------------------- */
static const real filter_inner_product_threshold = R_LIT(1.0e-16); /* src/filter_mod.F90:31 */

/* the SMOOTHING block, e.g. src/dycore_mod.F90:212-219: w2 may be NULL (weight = w1) */
static void smooth_row(orc_model *m, int half, int j, real *d, const real *w1, const real *w2) {
  const int nlon = m->nlon;
  real s1 = 0, s2 = 0;
  int i;
  for (i = 1; i <= nlon; i++)
    s1 = s1 + A2(d, i, j) * (w2 ? (A2(w1, i, j) + A2(w2, i, j)) : A2(w1, i, j));
  if (R_FABS(s1) > filter_inner_product_threshold) {
    if (half) filter_array_at_half_lat(m, j, d);
    else filter_array_at_full_lat(m, j, d);
    for (i = 1; i <= nlon; i++)
      s2 = s2 + A2(d, i, j) * (w2 ? (A2(w1, i, j) + A2(w2, i, j)) : A2(w1, i, j));
    for (i = 1; i <= nlon; i++) A2(d, i, j) = A2(d, i, j) * s1 / s2;
  }
}

/* Moving reduced tendency of one row -- SPECIFIED extension, not in the reference commit.
   On a row with reduction factor r the tendency is averaged over the nlon / r cells of the zonally reduced grid and
   handed back to the fine cells, for each of the r possible offsets of the reduced grid, and the r results are
   averaged ("moving"): T'(i) = sum_{|d| < r} (r - |d|) T(i + d) / r^2, periodic in i.  sum_i T' = sum_i T.
   With use_reduce_tend_smooth the row is then rescaled by s1 / s2 exactly as the filter blocks do
   (src/dycore_mod.F90:212-219), including the |s1| threshold around the whole block. */
static void reduce_average(orc_model *m, int j, int r, real *d) {
  const int nlon = m->nlon;
  real acc;
  int i, dd, ii;
  for (i = 1; i <= nlon; i++) {
    acc = 0;
    for (dd = -(r - 1); dd <= r - 1; dd++) {
      ii = i + dd;
      if (ii < 1) ii += nlon;
      if (ii > nlon) ii -= nlon;
      acc = acc + (real)(r - (dd < 0 ? -dd : dd)) * A2(d, ii, j);
    }
    m->reduce_tmp[i] = acc / (real)(r * r);
  }
  for (i = 1; i <= nlon; i++) A2(d, i, j) = m->reduce_tmp[i];
}
static void reduce_row(orc_model *m, int j, int r, real *d, const real *w1, const real *w2) {
  const int nlon = m->nlon, smooth = m->cfg.use_reduce_tend_smooth;
  real s1 = 0, s2 = 0;
  int i;
  if (smooth) {
    for (i = 1; i <= nlon; i++)
      s1 = s1 + A2(d, i, j) * (w2 ? (A2(w1, i, j) + A2(w2, i, j)) : A2(w1, i, j));
    if (!(R_FABS(s1) > filter_inner_product_threshold)) return;
  }
  reduce_average(m, j, r, d);
  if (smooth) {
    for (i = 1; i <= nlon; i++)
      s2 = s2 + A2(d, i, j) * (w2 ? (A2(w1, i, j) + A2(w2, i, j)) : A2(w1, i, j));
    for (i = 1; i <= nlon; i++) A2(d, i, j) = A2(d, i, j) * s1 / s2;
  }
}
/* what space_operators does with a finished tendency row: the reference's SMOOTHING block on a filter row, the
   moving reduced tendency on a reduced row (fast and unsplit passes; the slow pass only with reduce_adv_lon) */
static void finish_row(orc_model *m, int half, int j, int pass, real *d, const real *w1, const real *w2) {
  const int r = half ? m->reduce_half[j] : m->reduce_full[j];
  if (half ? m->filter_half[j] : m->filter_full[j]) smooth_row(m, half, j, d, w1, w2);
  else if (r > 1 && (pass != ORC_PASS_SLOW || m->cfg.reduce_adv_lon)) reduce_row(m, j, r, d, w1, w2);
}

static void zero2(const orc_model *m, real *a) { memset(a, 0, m->NE * sizeof(real)); }

static void space_operators(orc_model *m, state_t *s, tend_t *t, int pass) {
  const int nlon = m->nlon, nlat = m->nlat;
  int i, j;
  switch (pass) {
    case ORC_PASS_ALL:
      zonal_momentum_advection_operator(m, s, t);
      meridional_momentum_advection_operator(m, s, t);
      coriolis_operator(m, s, t);
      zonal_pressure_gradient_force_operator(m, s, t);
      meridional_pressure_gradient_force_operator(m, s, t);
      zonal_mass_divergence_operator(m, s, t);
      meridional_mass_divergence_operator(m, s, t);
      for (j = 2; j <= nlat - 1; j++) {
        for (i = 1; i <= nlon; i++)
          A2(t->du, i, j) = -A2(t->u_adv_lon, i, j) - A2(t->u_adv_lat, i, j) + A2(t->fv, i, j) - A2(t->u_pgf, i, j);
        finish_row(m, 0, j, pass, t->du, s->iu, NULL);
      }
      for (j = 1; j <= nlat - 1; j++) {
        for (i = 1; i <= nlon; i++)
          A2(t->dv, i, j) = -A2(t->v_adv_lon, i, j) - A2(t->v_adv_lat, i, j) - A2(t->fu, i, j) - A2(t->v_pgf, i, j);
        finish_row(m, 1, j, pass, t->dv, s->iv, NULL);
      }
      for (j = 1; j <= nlat; j++) {
        for (i = 1; i <= nlon; i++)
          A2(t->dgd, i, j) = -A2(t->mass_div_lon, i, j) - A2(t->mass_div_lat, i, j);
        finish_row(m, 0, j, pass, t->dgd, s->gd, m->ghs);
      }
      break;
    case ORC_PASS_SLOW:
      zonal_momentum_advection_operator(m, s, t);
      meridional_momentum_advection_operator(m, s, t);
      for (j = 2; j <= nlat - 1; j++) {
        for (i = 1; i <= nlon; i++) A2(t->du, i, j) = -A2(t->u_adv_lon, i, j) - A2(t->u_adv_lat, i, j);
        finish_row(m, 0, j, pass, t->du, s->iu, NULL);
      }
      for (j = 1; j <= nlat - 1; j++) {
        for (i = 1; i <= nlon; i++) A2(t->dv, i, j) = -A2(t->v_adv_lon, i, j) - A2(t->v_adv_lat, i, j);
        finish_row(m, 1, j, pass, t->dv, s->iv, NULL);
      }
      zero2(m, t->dgd);
      break;
    case ORC_PASS_FAST:
      coriolis_operator(m, s, t);
      zonal_pressure_gradient_force_operator(m, s, t);
      meridional_pressure_gradient_force_operator(m, s, t);
      zonal_mass_divergence_operator(m, s, t);
      meridional_mass_divergence_operator(m, s, t);
      for (j = 1; j <= nlat; j++) {
        for (i = 1; i <= nlon; i++)
          A2(t->dgd, i, j) = -A2(t->mass_div_lon, i, j) - A2(t->mass_div_lat, i, j);
        finish_row(m, 0, j, pass, t->dgd, s->gd, m->ghs);
      }
      for (j = 2; j <= nlat - 1; j++) {
        for (i = 1; i <= nlon; i++) A2(t->du, i, j) = A2(t->fv, i, j) - A2(t->u_pgf, i, j);
        finish_row(m, 0, j, pass, t->du, s->iu, NULL);
      }
      for (j = 1; j <= nlat - 1; j++) {
        for (i = 1; i <= nlon; i++) A2(t->dv, i, j) = -A2(t->fu, i, j) - A2(t->v_pgf, i, j);
        finish_row(m, 1, j, pass, t->dv, s->iv, NULL);
      }
      break;
  }
}

/* ---- update_state, src/dycore_mod.F90:600-652 ------------------------------------------------------ */
static void update_state(orc_model *m, real dt, const tend_t *t, const state_t *old, state_t *new_) {
  const int nlon = m->nlon, nlat = m->nlat;
  int i, j;
  for (j = 1; j <= nlat; j++)
    for (i = 1; i <= nlon; i++) A2(new_->gd, i, j) = A2(old->gd, i, j) + dt * A2(t->dgd, i, j);
  fill_halo(m, new_->gd);
  for (j = -1; j <= nlat + 2; j++)
    for (i = -1; i <= nlon + 2; i++) A2(new_->igd, i, j) = R_SQRT(A2(new_->gd, i, j));
  for (j = 1; j <= nlat; j++)
    for (i = 1; i <= nlon; i++) A2(new_->iu, i, j) = A2(old->iu, i, j) + dt * A2(t->du, i, j);
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) A2(new_->iv, i, j) = A2(old->iv, i, j) + dt * A2(t->dv, i, j);
  for (j = 2; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++)
      A2(new_->u, i, j) = A2(new_->iu, i, j) * R_LIT(2.0) / (A2(new_->igd, i, j) + A2(new_->igd, i + 1, j));
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++)
      A2(new_->v, i, j) = A2(new_->iv, i, j) * R_LIT(2.0) / (A2(new_->igd, i, j) + A2(new_->igd, i, j + 1));
  fill_halo(m, new_->iu);
  fill_halo(m, new_->iv);
  fill_halo(m, new_->u);
  fill_halo(m, new_->v);
}

/* ---- inner products, src/types_mod.F90:347-397 ----------------------------------------------------- */
static real inner_product_tend_tend(const orc_model *m, const tend_t *a, const tend_t *b) {
  const int nlon = m->nlon, nlat = m->nlat;
  real res = 0;
  int i, j;
  for (j = 2; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) res = res + A2(a->du, i, j) * A2(b->du, i, j) * m->full_cos_lat[j];
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) res = res + A2(a->dv, i, j) * A2(b->dv, i, j) * m->half_cos_lat[j];
  for (j = 1; j <= nlat; j++)
    for (i = 1; i <= nlon; i++) res = res + A2(a->dgd, i, j) * A2(b->dgd, i, j) * m->full_cos_lat[j];
  return res;
}
static real inner_product_tend_state(const orc_model *m, const tend_t *a, const state_t *s) {
  const int nlon = m->nlon, nlat = m->nlat;
  real res = 0;
  int i, j;
  for (j = 2; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) res = res + A2(a->du, i, j) * A2(s->iu, i, j) * m->full_cos_lat[j];
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) res = res + A2(a->dv, i, j) * A2(s->iv, i, j) * m->half_cos_lat[j];
  for (j = 1; j <= nlat; j++)
    for (i = 1; i <= nlon; i++) res = res + A2(a->dgd, i, j) * A2(s->gd, i, j) * m->full_cos_lat[j];
  return res;
}

/* ---- predict_correct, src/dycore_mod.F90:754-792 --------------------------------------------------- */
static void predict_correct(orc_model *m, real time_step_size, int old, int new_, int pass) {
  real dt, ip1, ip2, beta;
  dt = time_step_size * R_LIT(0.5);
  space_operators(m, STATE(m, old), TEND(m, old), pass);
  update_state(m, dt, TEND(m, old), STATE(m, old), STATE(m, new_));
  space_operators(m, STATE(m, new_), TEND(m, old), pass);
  update_state(m, dt, TEND(m, old), STATE(m, old), STATE(m, new_));
  space_operators(m, STATE(m, new_), TEND(m, new_), pass);
  ip1 = inner_product_tend_tend(m, TEND(m, old), TEND(m, new_));
  ip2 = inner_product_tend_tend(m, TEND(m, new_), TEND(m, new_));
  beta = (m->cfg.qcon_modified && ip1 != 0 && ip2 != 0) ? ip1 / ip2 : R_LIT(1.0);
  m->beta = beta;
  dt = time_step_size * beta;
  update_state(m, dt, TEND(m, new_), STATE(m, old), STATE(m, new_));
}

/* ---- tend algebra on du,dv,dgd, src/types_mod.F90:229-345 ------------------------------------------ */
static void tend_zero(const orc_model *m, tend_t *t) { zero2(m, t->du); zero2(m, t->dv); zero2(m, t->dgd); }
static void tend_add(const orc_model *m, const tend_t *a, tend_t *b) { /* b = b + a */
  size_t k;
  for (k = 0; k < m->NE; k++) { b->du[k] = b->du[k] + a->du[k]; b->dv[k] = b->dv[k] + a->dv[k]; b->dgd[k] = b->dgd[k] + a->dgd[k]; }
}
static void tend_sub(const orc_model *m, const tend_t *a, tend_t *b) { /* b = b - a */
  size_t k;
  for (k = 0; k < m->NE; k++) { b->du[k] = b->du[k] - a->du[k]; b->dv[k] = b->dv[k] - a->dv[k]; b->dgd[k] = b->dgd[k] - a->dgd[k]; }
}
static void tend_scale(const orc_model *m, real sc, tend_t *t) {
  size_t k;
  for (k = 0; k < m->NE; k++) { t->du[k] = t->du[k] * sc; t->dv[k] = t->dv[k] * sc; t->dgd[k] = t->dgd[k] * sc; }
}
static void state_copy(const orc_model *m, const state_t *a, state_t *b) {
  const size_t nb = m->NE * sizeof(real);
  memcpy(b->u, a->u, nb); memcpy(b->v, a->v, nb); memcpy(b->gd, a->gd, nb);
  memcpy(b->iu, a->iu, nb); memcpy(b->iv, a->iv, nb); memcpy(b->igd, a->igd, nb);
}

static void tend_copy(const orc_model *m, const tend_t *a, tend_t *b) {
  const size_t nb = m->NE * sizeof(real);
  memcpy(b->du, a->du, nb); memcpy(b->dv, a->dv, nb); memcpy(b->dgd, a->dgd, nb);
}
static void tend_axpby(const orc_model *m, real alpha, const tend_t *x, real beta, tend_t *y) { /* y = beta y + alpha x */
  size_t k;
  for (k = 0; k < m->NE; k++) {
    y->du[k] = beta * y->du[k] + alpha * x->du[k];
    y->dv[k] = beta * y->dv[k] + alpha * x->dv[k];
    y->dgd[k] = beta * y->dgd[k] + alpha * x->dgd[k];
  }
}

/* ---- runge_kutta: SPECIFIED extension, not in the reference commit -------------
   The integrator slot of dycore_mod.F90:43-53,78-83 (`time_scheme = 'runge_kutta'`, params_mod.F90:40-44) filled with
   an explicit Runge-Kutta step in increment form, phi' = phi + beta dt K, K = sum b_i k_i, k_i = L(phi_i) from the same
   space_operators / update_state / tend algebra the reference has, and the energy fix of predict_correct carried
   over.  E(phi + beta dt K) = E(phi) + 2 beta dt <K, phi> + beta^2 dt^2 <K, K> asks for beta = -2 <K, phi> / (dt <K, K>);
   with the antisymmetry <L(psi), psi> = 0 at every stage state, <K, phi> is a combination of tendency products -- the
   form predict_correct itself uses (beta = <k2, k3> / <k3, k3>, :779-781), free of the cancellation in <K, phi>:
     time_order 3 (Shu-Osher SSP-RK3): phi1 = phi + dt k1; phi2 = phi + dt/4 (k1 + k2); K = 1/6 (k1 + k2) + 2/3 k3;
                                       beta = (<k1, k2> + <k1 + k2, k3>) / (3 <K, K>)
     time_order 4 (classical RK4):     phi1 = phi + dt/2 k1; phi2 = phi + dt/2 k2; phi3 = phi + dt k3;
                                       K = 1/6 ((k1 + 2 k2 + 2 k3) + k4); beta = (<k1,k2> + <k2,k3> + <k3,k4>) / (3 <K, K>)
   Scratch: tend slots -2 (K), -1 and `old` (k_i) -- the first two are isp's, which never runs through this slot. */
static void runge_kutta(orc_model *m, real dt, int old, int new_, int pass) {
  tend_t *K = TEND(m, -2), *ka = TEND(m, -1), *kb = TEND(m, old);
  state_t *S0 = STATE(m, old);
  state_t *S1 = STATE(m, new_);
  real ip1, ip2, beta;
  space_operators(m, S0, ka, pass);                      /* k1 */
  tend_copy(m, ka, K);
  if (m->cfg.time_order == 4) {
    update_state(m, dt * R_LIT(0.5), ka, S0, S1);
    space_operators(m, S1, kb, pass);                    /* k2 */
    ip1 = inner_product_tend_tend(m, ka, kb);
    tend_axpby(m, R_LIT(2.0), kb, R_LIT(1.0), K);
    update_state(m, dt * R_LIT(0.5), kb, S0, S1);
    space_operators(m, S1, ka, pass);                    /* k3 */
    ip1 = ip1 + inner_product_tend_tend(m, kb, ka);
    tend_axpby(m, R_LIT(2.0), ka, R_LIT(1.0), K);
    update_state(m, dt, ka, S0, S1);
    space_operators(m, S1, kb, pass);                    /* k4 */
    ip1 = ip1 + inner_product_tend_tend(m, ka, kb);
    tend_axpby(m, R_LIT(1.0) / R_LIT(6.0), kb, R_LIT(1.0) / R_LIT(6.0), K);
  } else {
    update_state(m, dt, ka, S0, S1);
    space_operators(m, S1, ka, pass);                    /* k2 */
    ip1 = inner_product_tend_tend(m, K, ka);
    tend_axpby(m, R_LIT(1.0), ka, R_LIT(1.0), K);
    update_state(m, dt * R_LIT(0.25), K, S0, S1);
    space_operators(m, S1, ka, pass);                    /* k3 */
    ip1 = ip1 + inner_product_tend_tend(m, K, ka);
    tend_axpby(m, R_LIT(2.0) / R_LIT(3.0), ka, R_LIT(1.0) / R_LIT(6.0), K);
  }
  ip2 = inner_product_tend_tend(m, K, K);
  beta = (m->cfg.qcon_modified && ip1 != 0 && ip2 != 0) ? ip1 / (R_LIT(3.0) * ip2) : R_LIT(1.0);
  m->beta = beta;
  update_state(m, dt * beta, K, S0, S1);
}
static void integrator(orc_model *m, real dt, int old, int new_, int pass) {   /* dycore_mod.F90:43-53 */
  if (m->cfg.time_scheme == ORC_TIME_RUNGE_KUTTA) runge_kutta(m, dt, old, new_, pass);
  else predict_correct(m, dt, old, new_, pass);
}

/* ---- csp2_splitting, src/dycore_mod.F90:671-687 ---------------------------------------------------- */
static void csp2_splitting(orc_model *m) {
  const real dtm = (real)m->cfg.time_step_size;
  const real fast_dt = dtm / m->cfg.subcycles;
  int t1 = 0, t2 = m->old_idx, sub, tmp;
  integrator(m, R_LIT(0.5) * dtm, m->old_idx, t1, ORC_PASS_SLOW);
  for (sub = 1; sub <= m->cfg.subcycles; sub++) {
    integrator(m, fast_dt, t1, t2, ORC_PASS_FAST);
    tmp = t1; t1 = t2; t2 = tmp;
  }
  integrator(m, R_LIT(0.5) * dtm, t1, m->new_idx, ORC_PASS_SLOW);
}

/* ---- isp_splitting, src/dycore_mod.F90:689-752 ----------------------------------------------------- */
static void isp_splitting(orc_model *m) {
  const real dtm = (real)m->cfg.time_step_size;
  const real fast_dt = dtm / m->cfg.subcycles;
  real half_dt = dtm * R_LIT(0.5), ip1, ip2, beta;
  int t1 = m->old_idx, t2 = 0, sub, tmp;
  const int new_ = m->new_idx;
  state_t *saved_state = STATE(m, -1);
  tend_t *slow_tend = TEND(m, -2), *accum_fast_tend = TEND(m, -1);
  state_copy(m, STATE(m, m->old_idx), saved_state);
  tend_zero(m, accum_fast_tend);
  space_operators(m, STATE(m, m->old_idx), slow_tend, ORC_PASS_SLOW);
  for (sub = 1; sub <= m->cfg.subcycles; sub++) {
    space_operators(m, STATE(m, t1), TEND(m, t1), ORC_PASS_FAST);
    tend_add(m, slow_tend, TEND(m, t1));
    update_state(m, fast_dt * R_LIT(0.5), TEND(m, t1), STATE(m, t1), STATE(m, t2));
    space_operators(m, STATE(m, t2), TEND(m, t1), ORC_PASS_FAST);
    tend_add(m, slow_tend, TEND(m, t1));
    update_state(m, fast_dt * R_LIT(0.5), TEND(m, t1), STATE(m, t1), STATE(m, t2));
    space_operators(m, STATE(m, t2), TEND(m, t2), ORC_PASS_FAST);
    tend_add(m, TEND(m, t2), accum_fast_tend);
    tend_add(m, slow_tend, TEND(m, t2));
    update_state(m, fast_dt, TEND(m, t2), STATE(m, t1), STATE(m, t2));
    tmp = t1; t1 = t2; t2 = tmp;
  }
  tend_scale(m, R_LIT(2.0) / m->cfg.subcycles, accum_fast_tend);
  space_operators(m, STATE(m, t1), TEND(m, t1), ORC_PASS_SLOW);
  tend_sub(m, slow_tend, TEND(m, t1));
  update_state(m, half_dt, TEND(m, t1), STATE(m, t1), STATE(m, new_));
  space_operators(m, STATE(m, new_), TEND(m, t1), ORC_PASS_SLOW);
  tend_sub(m, slow_tend, TEND(m, t1));
  update_state(m, half_dt, TEND(m, t1), STATE(m, t1), STATE(m, new_));
  space_operators(m, STATE(m, new_), TEND(m, new_), ORC_PASS_SLOW);
  tend_add(m, slow_tend, TEND(m, new_));
  tend_add(m, accum_fast_tend, TEND(m, new_));
  ip1 = inner_product_tend_state(m, TEND(m, new_), saved_state);
  ip2 = inner_product_tend_tend(m, TEND(m, new_), TEND(m, new_));
  beta = (m->cfg.qcon_modified && ip1 != 0 && ip2 != 0) ? ip1 / ip2 : R_LIT(1.0);
  beta = beta * R_LIT(4.0) / dtm;
  m->beta = beta;
  half_dt = half_dt * beta;
  update_state(m, half_dt, TEND(m, new_), saved_state, STATE(m, new_));
}

/* ---- ordinary_diffusion, src/diffusion_mod.F90:74-217 ----------------------------------------------- */
static void ordinary_diffusion(orc_model *m, real dt, state_t *s) {
  const int nlon = m->nlon, nlat = m->nlat, norder = m->cfg.diffusion_order / 2;
  const real coef = (real)m->cfg.diffusion_coef;
  const real *fcl = m->full_cos_lat, *hcl = m->half_cos_lat;
  real *ud = m->dud, *vd = m->dvd, *gdd = m->dgdd, *u = m->du_, *v = m->dv_, *gd = m->dgd_;
  const size_t nb = m->NE * sizeof(real);
  real sp, np;
  int i, j, order, sign;
  memcpy(u, s->u, nb); memcpy(v, s->v, nb); memcpy(gd, s->gd, nb);
  for (order = 1; order <= norder; order++) {
    for (j = 2; j <= nlat - 1; j++)
      for (i = 1; i <= nlon; i++)
        A2(gdd, i, j) = (A2(gd, i + 1, j) - 2 * A2(gd, i, j) + A2(gd, i - 1, j)) / (m->full_dlon[j] * m->full_dlon[j]) +
                        ((A2(gd, i, j + 1) - A2(gd, i, j)) * hcl[j] - (A2(gd, i, j) - A2(gd, i, j - 1)) * hcl[j - 1]) /
                            (m->full_dlat[j] * m->full_dlat[j]) * fcl[j];
    j = 1;
    sp = 0;
    for (i = 1; i <= nlon; i++) sp = sp + A2(gd, i, j + 1) - A2(gd, i, j);
    sp = sp * hcl[j] / (m->full_dlat[j] * m->full_dlat[j]) * fcl[j] / nlon;
    for (i = -1; i <= nlon + 2; i++) A2(gdd, i, j) = sp;
    j = nlat;
    np = 0;
    for (i = 1; i <= nlon; i++) np = np - (A2(gd, i, j) - A2(gd, i, j - 1));
    np = np * hcl[j - 1] / (m->full_dlat[j] * m->full_dlat[j]) * fcl[j] / nlon;
    for (i = -1; i <= nlon + 2; i++) A2(gdd, i, j) = np;
    for (j = 2; j <= nlat - 1; j++)
      for (i = 1; i <= nlon; i++)
        A2(ud, i, j) = (A2(u, i + 1, j) - 2 * A2(u, i, j) + A2(u, i - 1, j)) / (m->full_dlon[j] * m->full_dlon[j]) +
                       ((A2(u, i, j + 1) - A2(u, i, j)) * hcl[j] - (A2(u, i, j) - A2(u, i, j - 1)) * hcl[j - 1]) /
                           (m->full_dlat[j] * m->full_dlat[j]) * fcl[j];
    for (j = 1; j <= nlat - 1; j++)
      for (i = 1; i <= nlon; i++)
        A2(vd, i, j) = (A2(v, i + 1, j) - 2 * A2(v, i, j) + A2(v, i - 1, j)) / (m->half_dlon[j] * m->half_dlon[j]);
    for (j = 2; j <= nlat - 2; j++)
      for (i = 1; i <= nlon; i++)
        A2(vd, i, j) = A2(vd, i, j) +
                       ((A2(v, i, j + 1) - A2(v, i, j)) * fcl[j + 1] - (A2(v, i, j) - A2(v, i, j - 1)) * fcl[j]) /
                           (m->half_dlat[j] * m->half_dlat[j]) * hcl[j];
    j = 1;
    for (i = 1; i <= nlon; i++)
      A2(vd, i, j) = A2(vd, i, j) + (A2(v, i, j + 1) - A2(v, i, j)) * fcl[j + 1] / (m->half_dlat[j] * m->half_dlat[j]) * hcl[j];
    j = nlat - 1;
    for (i = 1; i <= nlon; i++)
      A2(vd, i, j) = A2(vd, i, j) - (A2(v, i, j) - A2(v, i, j - 1)) * fcl[j] / (m->half_dlat[j] * m->half_dlat[j]) * hcl[j];
    if (order != norder) {
      fill_halo(m, gdd); fill_halo(m, ud); fill_halo(m, vd);
      memcpy(gd, gdd, nb); memcpy(u, ud, nb); memcpy(v, vd, nb);
    }
  }
  for (j = 2; j <= nlat - 1; j++)
    if (m->filter_full[j]) {
      filter_array_at_full_lat(m, j, gdd);
      filter_array_at_full_lat(m, j, ud);
    }
  for (j = 1; j <= nlat - 1; j++)
    if (m->filter_half[j]) filter_array_at_half_lat(m, j, vd);
  /* specified extension: the diffusion tendencies of reduced rows get the moving reduced average (no rescale), as the
     filter rows get the plain filter above */
  for (j = 2; j <= nlat - 1; j++)
    if (m->reduce_full[j] > 1) {
      reduce_average(m, j, m->reduce_full[j], gdd);
      reduce_average(m, j, m->reduce_full[j], ud);
    }
  for (j = 1; j <= nlat - 1; j++)
    if (m->reduce_half[j] > 1) reduce_average(m, j, m->reduce_half[j], vd);
  sign = ((norder + 1) % 2 == 0) ? 1 : -1;
  for (j = 1; j <= nlat; j++) {
    for (i = 1; i <= nlon; i++) A2(s->gd, i, j) = A2(s->gd, i, j) + sign * dt * coef * A2(gdd, i, j);
    for (i = 1; i <= nlon; i++) A2(s->u, i, j) = A2(s->u, i, j) + sign * dt * coef * A2(ud, i, j);
  }
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) A2(s->v, i, j) = A2(s->v, i, j) + sign * dt * coef * A2(vd, i, j);
  fill_halo(m, s->gd); fill_halo(m, s->u); fill_halo(m, s->v);
  iap_transform(m, s);
}

/* ---- time_integrate, src/dycore_mod.F90:654-669 ---------------------------------------------------- */
static void time_integrate(orc_model *m) {
  switch (m->cfg.split_scheme) {
    case ORC_SPLIT_CSP2: csp2_splitting(m); break;
    case ORC_SPLIT_ISP: isp_splitting(m); break;
    default: integrator(m, (real)m->cfg.time_step_size, m->old_idx, m->new_idx, ORC_PASS_ALL);
  }
  if (m->cfg.use_diffusion) ordinary_diffusion(m, (real)m->cfg.time_step_size, STATE(m, m->new_idx));
}

/* ---- diag_run, src/diag_mod.F90:42-121 -------------------------------------------------------------- */
static real diag_total_energy(const orc_model *m, const state_t *s) {
  const int nlon = m->nlon, nlat = m->nlat;
  real res = 0;
  int i, j;
  for (j = 2; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) res = res + A2(s->iu, i, j) * A2(s->iu, i, j) * m->full_cos_lat[j];
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) res = res + A2(s->iv, i, j) * A2(s->iv, i, j) * m->half_cos_lat[j];
  for (j = 1; j <= nlat; j++)
    for (i = 1; i <= nlon; i++)
      res = res + (A2(s->gd, i, j) + A2(m->ghs, i, j)) * (A2(s->gd, i, j) + A2(m->ghs, i, j)) * m->full_cos_lat[j];
  return res;
}
static int diag_run(orc_model *m, const state_t *s) {
  const int nlon = m->nlon, nlat = m->nlat;
  const real radius = RADIUS_;
  real um1, up1, vm1, vp1;
  int i, j;
  for (j = 2; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) {
      um1 = A2(s->u, i - 1, j);
      up1 = A2(s->u, i, j);
      vm1 = A2(s->v, i, j - 1) * m->half_cos_lat[j - 1];
      vp1 = A2(s->v, i, j) * m->half_cos_lat[j];
      A2(m->div, i, j) = (up1 - um1) / m->full_dlon[j] + (vp1 - vm1) / m->full_dlat[j];
    }
  fill_halo(m, m->div);
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) {
      um1 = A2(s->u, i, j);
      up1 = A2(s->u, i, j + 1);
      vm1 = A2(s->v, i, j);
      vp1 = A2(s->v, i + 1, j);
      A2(m->vor, i, j) = (vp1 - vm1) / m->half_dlon[j] - (up1 - um1) / m->half_dlat[j];
    }
  fill_halo(m, m->vor);
  m->total_mass = 0;
  for (j = 1; j <= nlat; j++)
    for (i = 1; i <= nlon; i++)
      m->total_mass = m->total_mass + m->full_cos_lat[j] * m->dlon * m->dlat * A2(s->gd, i, j);
  m->total_mass = m->total_mass * (radius * radius);
  m->total_energy = diag_total_energy(m, s);
  if (R_ISNAN(m->total_mass)) { snprintf(g_err, sizeof g_err, "Total mass is NaN!"); return 1; }
  if (R_ISNAN(m->total_energy)) { snprintf(g_err, sizeof g_err, "Total energy is NaN!"); return 1; }
  return 0;
}

/* ---- test-case plugins ------------------------------------------------------------------------------ */
/* rossby_haurwitz_wave_test_mod.F90:33-85 */
static void ic_rossby_haurwitz(orc_model *m, real R, real omg, real gd0) {
  const int nlon = m->nlon, nlat = m->nlat;
  const real radius = RADIUS_, omega = OMEGA_;
  state_t *s = STATE(m, 1);
  real lon, cos_lat, sin_lat, a, b, c;
  int i, j;
  zero2(m, m->ghs);
  for (j = 1; j <= nlat; j++) {
    cos_lat = m->full_cos_lat[j];
    sin_lat = m->full_sin_lat[j];
    for (i = 1; i <= nlon; i++) {
      lon = m->half_lon[i];
      a = cos_lat;
      b = R * R_POW(cos_lat, R - 1) * (sin_lat * sin_lat) * R_COS(R * lon);
      c = R_POW(cos_lat, R + 1) * R_COS(R * lon);
      A2(s->u, i, j) = radius * omg * (a + b - c);
    }
  }
  fill_halo(m, s->u);
  for (j = 1; j <= nlat - 1; j++) {
    cos_lat = m->half_cos_lat[j];
    sin_lat = m->half_sin_lat[j];
    for (i = 1; i <= nlon; i++) {
      lon = m->full_lon[i];
      a = R * R_POW(cos_lat, R - 1) * sin_lat * R_SIN(R * lon);
      A2(s->v, i, j) = -radius * omg * a;
    }
  }
  fill_halo(m, s->v);
  for (j = 1; j <= nlat; j++) {
    cos_lat = m->full_cos_lat[j];
    a = R_LIT(0.5) * omg * (2 * omega + omg) * (cos_lat * cos_lat) +
        R_LIT(0.25) * (omg * omg) *
            ((R + 1) * R_POW(cos_lat, 2 * R + 2) + (2 * (R * R) - R - 2) * R_POW(cos_lat, 2 * R) -
             2 * (R * R) * R_POW(cos_lat, 2 * R - 2));
    b = 2 * (omega + omg) * omg * R_POW(cos_lat, R) * (R * R + 2 * R + 2 - (R + 1) * (R + 1) * (cos_lat * cos_lat)) /
        (R + 1) / (R + 2);
    c = R_LIT(0.25) * (omg * omg) * R_POW(cos_lat, 2 * R) * ((R + 1) * (cos_lat * cos_lat) - R - 2);
    for (i = 1; i <= nlon; i++) {
      lon = m->full_lon[i];
      A2(s->gd, i, j) = gd0 + (radius * radius) * (a + b * R_COS(R * lon) + c * R_COS(2 * R * lon));
    }
  }
  fill_halo(m, s->gd);
}

/* steady_geostrophic_flow_test_mod.F90:20-62 and the flow part of mountain_zonal_flow_test_mod.F90:69-96.
   B11: `sin_lon = full_cos_lon` in the reference is harmless because sin(alpha) = 0. */
static void ic_zonal_flow(orc_model *m, real u0, real gd0) {
  const int nlon = m->nlon, nlat = m->nlat;
  const real radius = RADIUS_, omega = OMEGA_, alpha = 0;
  const real cos_alpha = R_COS(alpha), sin_alpha = R_SIN(alpha);
  state_t *s = STATE(m, 1);
  real cos_lat, sin_lat, cos_lon, sin_lon, t;
  int i, j;
  for (j = 2; j <= nlat - 1; j++) {
    cos_lat = m->full_cos_lat[j];
    sin_lat = m->full_sin_lat[j];
    for (i = 1; i <= nlon; i++) {
      cos_lon = m->half_cos_lon[i];
      A2(s->u, i, j) = u0 * (cos_lat * cos_alpha + cos_lon * sin_lat * sin_alpha);
    }
  }
  fill_halo(m, s->u);
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) {
      sin_lon = m->full_cos_lon[i];
      A2(s->v, i, j) = -u0 * sin_lon * sin_alpha;
    }
  fill_halo(m, s->v);
  for (j = 1; j <= nlat; j++) {
    cos_lat = m->full_cos_lat[j];
    sin_lat = m->full_sin_lat[j];
    for (i = 1; i <= nlon; i++) {
      cos_lon = m->full_cos_lon[i];
      t = sin_lat * cos_alpha - cos_lon * cos_lat * sin_alpha;
      A2(s->gd, i, j) = gd0 - (radius * omega * u0 + (u0 * u0) * R_LIT(0.5)) * (t * t) - A2(m->ghs, i, j);
    }
  }
  fill_halo(m, s->gd);
}

static void ic_steady_geostrophic(orc_model *m) {
  const real u0 = 2 * PI_ * RADIUS_ / (12 * R_LIT(86400.0));
  zero2(m, m->ghs);
  ic_zonal_flow(m, u0, R_LIT(2.94e4));
}

/* mountain_zonal_flow_test_mod.F90:28-98 */
static void ic_mountain_zonal(orc_model *m, int smooth_mountain) {
  const int nlon = m->nlon, nlat = m->nlat;
  const real pi = PI_, lon0 = pi * R_LIT(1.5), lat0 = pi / R_LIT(6.0), ghs0 = R_LIT(2000.0) * G_;
  const real R = pi / R_LIT(9.0);
  real dlon, d, t;
  int i, j, k;
  zero2(m, m->ghs);
  for (j = 1; j <= nlat; j++)
    for (i = 1; i <= nlon; i++) {
      dlon = R_FABS(m->full_lon[i] - lon0);
      t = 2 * pi - dlon;
      dlon = dlon < t ? dlon : t;
      d = R_SQRT(dlon * dlon + (m->full_lat[j] - lat0) * (m->full_lat[j] - lat0));
      d = R < d ? R : d;
      A2(m->ghs, i, j) = ghs0 * (R_LIT(1.0) - d / R);
    }
  if (smooth_mountain) {
    for (k = 1; k <= 30; k++) {
      fill_halo(m, m->ghs);
      for (j = 2; j <= nlat - 1; j++)
        for (i = 1; i <= nlon; i++)
          A2(m->ghs, i, j) =
              A2(m->ghs, i, j) +
              (R_LIT(0.5) / 4) * (A2(m->ghs, i - 1, j) + A2(m->ghs, i, j + 1) + A2(m->ghs, i + 1, j) +
                                  A2(m->ghs, i, j - 1) - 4 * A2(m->ghs, i, j)) +
              (R_LIT(0.25) / 4) * (A2(m->ghs, i - 1, j - 1) + A2(m->ghs, i - 1, j + 1) + A2(m->ghs, i + 1, j + 1) +
                                   A2(m->ghs, i + 1, j - 1) - 4 * A2(m->ghs, i, j));
    }
  }
  fill_halo(m, m->ghs);
  ic_zonal_flow(m, R_LIT(20.0), R_LIT(5960.0) * G_);
}

/* jet_zonal_flow_test_mod.F90:16-105 */
static double jet_u_function(double lat) {
  const double pi = atan(1.0) * 4.0, u_max = 80.0, lat0 = pi / 7.0, lat1 = pi / 2.0 - lat0;
  const double en = exp(-4.0 / ((lat1 - lat0) * (lat1 - lat0)));
  if (lat <= lat0 || lat >= lat1) return 0.0;
  return u_max / en * exp(1 / (lat - lat0) / (lat - lat1));
}
static double jet_gh_integrand(double lat, void *ctx) {
  const double pi = atan(1.0) * 4.0, omega = 2.0 * pi / 86400.0, radius = 6.37122e6;
  double u = jet_u_function(lat), f = 2 * omega * sin(lat);
  (void)ctx;
  return radius * u * (f + tan(lat) / radius * u);
}
double orc_jet_gd_profile(double lat) {
  const double pi = atan(1.0) * 4.0, gh0 = 9.80616 * 1.0e4;
  double res, abserr;
  int neval;
  if (lat <= -0.5 * pi) return gh0;
  orc_qag21(jet_gh_integrand, NULL, -0.5 * pi, lat, 1.0e-10, 1.0e-3, &res, &abserr, &neval);
  return gh0 - res;
}
static void ic_jet_zonal(orc_model *m) {
  const int nlon = m->nlon, nlat = m->nlat;
  const real pi = PI_, ghd = G_ * 120, lat2 = pi / R_LIT(4.0), alpha = R_LIT(1.0) / R_LIT(3.0),
             beta = R_LIT(1.0) / R_LIT(15.0);
  state_t *s = STATE(m, 1);
  real base, t1, t2;
  int i, j;
  zero2(m, m->ghs);
  for (j = 1; j <= nlat; j++)
    for (i = 1; i <= nlon; i++) A2(s->u, i, j) = (real)jet_u_function((double)m->full_lat[j]);
  fill_halo(m, s->u);
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) A2(s->v, i, j) = 0;
  fill_halo(m, s->v);
  for (j = 1; j <= nlat; j++) {
    base = (j == 1) ? (real)(9.80616 * 1.0e4) : (real)orc_jet_gd_profile((double)m->full_lat[j]);
    for (i = 1; i <= nlon; i++) {
      t1 = (m->full_lon[i] - pi) / alpha;
      t2 = (lat2 - m->full_lat[j]) / beta;
      A2(s->gd, i, j) = base + ghd * R_COS(m->full_lat[j]) * R_EXP(-(t1 * t1)) * R_EXP(-(t2 * t2));
    }
  }
  fill_halo(m, s->gd);
}

/* ---- public API -------------------------------------------------------------------------------------- */
int orc_create(const orc_config *cfg, orc_model **out) {
  orc_model *m;
  int ier;
  *out = NULL;
  if (cfg->num_lon < 4 || cfg->num_lat < 5) {
    snprintf(g_err, sizeof g_err, "grid too small: %d x %d", cfg->num_lon, cfg->num_lat);
    return 2;
  }
  m = (orc_model *)calloc(1, sizeof(*m));
  m->cfg = *cfg;
  m->nlon = cfg->num_lon;
  m->nlat = cfg->num_lat;
  m->LD = m->nlon + 4;
  m->NE = (size_t)m->LD * (size_t)(m->nlat + 4);
  mesh_init(m);
  data_init(m);
  m->dud = alloc2(m); m->dvd = alloc2(m); m->dgdd = alloc2(m);
  m->du_ = alloc2(m); m->dv_ = alloc2(m); m->dgd_ = alloc2(m);
  m->vor = alloc2(m); m->div = alloc2(m);
  ier = filter_init(m);
  if (ier) { orc_destroy(m); return ier; }
  m->old_idx = 1; /* time_init, src/time_mod.F90:66-69 */
  m->new_idx = 2;
  m->beta = 1;
  *out = m;
  return 0;
}

static void free_state(state_t *s) { free(s->u); free(s->v); free(s->gd); free(s->iu); free(s->iv); free(s->igd); }
static void free_tend(tend_t *t) {
  free(t->u_adv_lon); free(t->u_adv_lat); free(t->v_adv_lon); free(t->v_adv_lat); free(t->fu); free(t->fv);
  free(t->u_pgf); free(t->v_pgf); free(t->mass_div_lon); free(t->mass_div_lat); free(t->du); free(t->dv); free(t->dgd);
}
void orc_destroy(orc_model *m) {
  int k;
  if (!m) return;
  free(m->full_lon); free(m->half_lon); free(m->full_lat); free(m->half_lat);
  free(m->full_cos_lon); free(m->half_cos_lon); free(m->full_sin_lon); free(m->half_sin_lon);
  free(m->full_cos_lat); free(m->half_cos_lat); free(m->full_sin_lat); free(m->half_sin_lat);
  free(m->full_f); free(m->half_f); free(m->full_c); free(m->half_c);
  free(m->full_dlon); free(m->half_dlon); free(m->full_dlat); free(m->half_dlat);
  for (k = 0; k < 4; k++) free_state(&m->state_[k]);
  for (k = 0; k < 5; k++) free_tend(&m->tend_[k]);
  free(m->ghs);
  free(m->filter_full); free(m->filter_half); free(m->cutoff_full); free(m->cutoff_half);
  free(m->reduce_full); free(m->reduce_half); free(m->reduce_tmp);
  orc_rfft_plan_destroy(m->plan);
  free(m->local_x);
  free(m->dud); free(m->dvd); free(m->dgdd); free(m->du_); free(m->dv_); free(m->dgd_);
  free(m->fp_u); free(m->fn_u); free(m->f_u); free(m->fp_v); free(m->fn_v); free(m->f_v);
  free(m->vor); free(m->div);
  free(m);
}

/* shallow_water_waves_test_mod.F90: Shamir & Paldor (2016) analytic Rossby wave, gH = 5e4, (n, k) = (5, 10).
   getPhaseSpeed :130-171 (Cardano's formula in complex binary64, as the reference computes it), getPsi :176-210,
   getAmplitudes :215-271 (waveFlag 0), getFields :276-321 at time 0, set_initial_condition :95-135.
   The amplitudes are evaluated at the FULL latitudes only; v on half row j uses vTilde(j) of full row j (:307-313). */
#include <complex.h>
double orc_swe_phase_speed(int wave_flag) {
  const double omega = 7.29212e-5, g = 9.80616, a = 6371220.0, H0 = 5.0e3, pi = 3.14159265358979323;
  const int n = 5, k = 10;
  const double sigma = 0.5 + pow(0.25 + k * k, 0.5);
  const double En = g * H0 / (a * a) * ((n + sigma) * (n + sigma));
  const double Delta0 = 3.0 * (k * k) * En;
  const double Delta4 = -54.0 * (k * k * k * k) * g * H0 * omega / (a * a);
  const double r2 = 0.5, r3 = 1.0 / 3.0;
  double Cj[3], mn, mx, mabs;
  int j;
  for (j = 1; j <= 3; j++) {
    double complex D = cpow((double complex)(Delta4 * Delta4 - 4.0 * (Delta0 * Delta0 * Delta0)), r2);
    D = cpow(r2 * (Delta4 + D), r3);
    D = D * cexp(2.0 * pi * I * j * r3);
    Cj[j - 1] = creal(-r3 / (k * k) * (D + Delta0 / D));
  }
  mn = mx = Cj[0];
  mabs = fabs(Cj[0]);
  for (j = 1; j < 3; j++) {
    if (Cj[j] < mn) mn = Cj[j];
    if (Cj[j] > mx) mx = Cj[j];
    if (fabs(Cj[j]) < mabs) mabs = fabs(Cj[j]);
  }
  return wave_flag == 0 ? -mabs : (wave_flag == 1 ? mx : mn);
}
static void ic_shallow_water_waves(orc_model *m) {
  const int nlon = m->nlon, nlat = m->nlat;
  const real omega = R_LIT(7.29212e-5), g = R_LIT(9.80616), a = R_LIT(6371220.0), H0 = R_LIT(5.0e3);
  const real pi = R_LIT(3.14159265358979323);
  const int k = 10;
  const real sigma = R_LIT(0.5) + R_POW(R_LIT(0.25) + k * k, R_LIT(0.5)), amp = R_LIT(1.0e-8), o2 = 2 * omega;
  const real a3 = sigma * (sigma + 1) * (sigma + 2), a4 = a3 * (sigma + 3), a5 = a4 * (sigma + 4);
  const real C = (real)orc_swe_phase_speed(0);
  state_t *s = STATE(m, 1);
  real *ut = alloc1(nlat), *vt = alloc1(nlat), *ht = alloc1(nlat);
  int i, j;
  zero2(m, m->ghs);
  for (j = 1; j <= nlat; j++) {
    const real lat = m->full_lat[j];
    const real sl = R_SIN(lat), cl = R_COS(lat), tl = R_TAN(lat);
    const real s2 = sl * sl, s4 = s2 * s2;   /* sin(lat)**2, **4: integer powers */
    const real C5 = (4 * a5 * s4 - 20 * a4 * s2 + 15 * a3) * sl / 15;
    const real C5p = (4 * a5 * s4 - 12 * a4 * s2 + 3 * a3) * cl / 3;
    const real psi = amp * R_POW(cl, sigma) * C5;
    const real dpsi = amp * R_POW(cl, sigma) * (-sigma * tl * C5 + C5p);
    const real Kp = (g * H0 + a * a * (C * C) * (cl * cl)) / (C * cl);
    const real Km = (g * H0 - a * a * (C * C) * (cl * cl)) / (C * cl);
    real v = R_POW(o2 * R_FABS(Km) / (cl * cl), R_LIT(0.5)) * psi;
    const real h = R_POW(o2 * R_FABS(Km) * (a * a) * (H0 * H0), R_LIT(0.5)) / Km *
                   (dpsi + tl * (R_LIT(0.5) * Kp / Km - o2 / C) * psi);
    ut[j] = (o2 * sl / C) * v + (g / a / cl / C) * h;
    vt[j] = k * v;
    ht[j] = h;
  }
  for (j = 1; j <= nlat; j++)
    for (i = 1; i <= nlon; i++) {
      A2(s->u, i, j) = ut[j] * R_COS(k * m->half_lon[i] - k * C * 0);
      A2(s->gd, i, j) = g * (ht[j] * R_COS(k * m->full_lon[i] - k * C * 0)) + R_LIT(5.0e4);
    }
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) A2(s->v, i, j) = vt[j] * R_COS(k * m->full_lon[i] - k * C * 0 - R_LIT(0.5) * pi);
  fill_halo(m, s->gd);
  fill_halo(m, s->u);
  fill_halo(m, s->v);
  free(ut); free(vt); free(ht);
}

int orc_set_initial_condition(orc_model *m, int test_case, const double *params, int nparams) {
  if (m->pole_reset) {
    snprintf(g_err, sizeof g_err, "initial condition must be set before orc_run_init (B11)");
    return 2;
  }
  switch (test_case) {
    case ORC_IC_ROSSBY_HAURWITZ: {
      real R = R_LIT(4.0), omg = R_LIT(7.848e-6), gd0 = R_LIT(8.0e3) * G_;
      if (params && nparams >= 3) { R = (real)params[0]; omg = (real)params[1]; gd0 = (real)params[2]; }
      ic_rossby_haurwitz(m, R, omg, gd0);
      break;
    }
    case ORC_IC_STEADY_GEOSTROPHIC: ic_steady_geostrophic(m); break;
    case ORC_IC_MOUNTAIN_ZONAL: ic_mountain_zonal(m, (params && nparams >= 1) ? (params[0] != 0.0) : 0); break;
    case ORC_IC_JET_ZONAL: ic_jet_zonal(m); break;
    case ORC_IC_SHALLOW_WATER_WAVES: ic_shallow_water_waves(m); break;
    default:
      snprintf(g_err, sizeof g_err, "Unknown test case %d!", test_case);
      return 2;
  }
  return 0;
}

static void put_full(const orc_model *m, real *dst, const double *src) {
  int i, j;
  for (j = 1; j <= m->nlat; j++)
    for (i = 1; i <= m->nlon; i++) A2(dst, i, j) = (real)src[(size_t)(j - 1) * m->nlon + (i - 1)];
}
static void put_half(const orc_model *m, real *dst, const double *src) {
  int i, j;
  for (j = 1; j <= m->nlat - 1; j++)
    for (i = 1; i <= m->nlon; i++) A2(dst, i, j) = (real)src[(size_t)(j - 1) * m->nlon + (i - 1)];
}
static void get_full(const orc_model *m, const real *src, double *dst) {
  int i, j;
  if (!dst) return;
  for (j = 1; j <= m->nlat; j++)
    for (i = 1; i <= m->nlon; i++) dst[(size_t)(j - 1) * m->nlon + (i - 1)] = (double)A2(src, i, j);
}
static void get_half(const orc_model *m, const real *src, double *dst) {
  int i, j;
  if (!dst) return;
  for (j = 1; j <= m->nlat - 1; j++)
    for (i = 1; i <= m->nlon; i++) dst[(size_t)(j - 1) * m->nlon + (i - 1)] = (double)A2(src, i, j);
}

int orc_set_state(orc_model *m, const double *u, const double *v, const double *gd, const double *ghs) {
  state_t *s = STATE(m, m->old_idx);
  put_full(m, s->u, u); fill_halo(m, s->u);
  put_half(m, s->v, v); fill_halo(m, s->v);
  put_full(m, s->gd, gd); fill_halo(m, s->gd);
  if (ghs) { put_full(m, m->ghs, ghs); fill_halo(m, m->ghs); }
  else zero2(m, m->ghs);
  if (m->run_inited) { iap_transform(m, s); return diag_run(m, s); }
  return 0;
}

int orc_run_init(orc_model *m) {
  if (!m->pole_reset) reset_cos_lat_at_poles(m);
  iap_transform(m, STATE(m, m->old_idx));
  m->run_inited = 1;
  return diag_run(m, STATE(m, m->old_idx));
}

int orc_step(orc_model *m, int nsteps) {
  int n, tmp;
  if (!m->run_inited) { snprintf(g_err, sizeof g_err, "orc_run_init not called"); return 2; }
  for (n = 0; n < nsteps; n++) {
    time_integrate(m);
    tmp = m->old_idx; m->old_idx = m->new_idx; m->new_idx = tmp; /* time_advance, src/time_mod.F90:122 */
    m->step++;
    if (diag_run(m, STATE(m, m->old_idx))) return 1;
  }
  return 0;
}

int orc_get_state(const orc_model *m, double *u, double *v, double *gd) {
  const state_t *s = &m->state_[m->old_idx + 1];
  get_full(m, s->u, u); get_half(m, s->v, v); get_full(m, s->gd, gd);
  return 0;
}
int orc_get_iap_state(const orc_model *m, double *iu, double *iv, double *igd) {
  const state_t *s = &m->state_[m->old_idx + 1];
  get_full(m, s->iu, iu); get_half(m, s->iv, iv); get_full(m, s->igd, igd);
  return 0;
}
int orc_get_ghs(const orc_model *m, double *ghs) { get_full(m, m->ghs, ghs); return 0; }
int orc_get_diag(const orc_model *m, double *mass, double *energy, double *beta) {
  if (mass) *mass = (double)m->total_mass;
  if (energy) *energy = (double)m->total_energy;
  if (beta) *beta = (double)m->beta;
  return 0;
}
int orc_get_vor_div(const orc_model *m, double *vor, double *div) {
  get_half(m, m->vor, vor); get_full(m, m->div, div);
  return 0;
}
int orc_get_step_count(const orc_model *m) { return m->step; }

int orc_space_operators(orc_model *m, int pass, double *du, double *dv, double *dgd) {
  tend_t *t = TEND(m, m->old_idx);
  if (!m->run_inited) { snprintf(g_err, sizeof g_err, "orc_run_init not called"); return 2; }
  space_operators(m, STATE(m, m->old_idx), t, pass);
  get_full(m, t->du, du); get_half(m, t->dv, dv); get_full(m, t->dgd, dgd);
  return 0;
}

/* check_antisymmetry, src/dycore_mod.F90:794-851 (plus the magnitude of each sum) */
int orc_check_antisymmetry(orc_model *m, double *sums) {
  const int nlon = m->nlon, nlat = m->nlat;
  const state_t *s = STATE(m, m->old_idx);
  const tend_t *t = TEND(m, m->old_idx);
  real ip[10], ab[10], w, x;
  int i, j, k;
  for (k = 0; k < 10; k++) ip[k] = ab[k] = 0;
#define ACC(k, term) do { x = (term); ip[k] += x; ab[k] += R_FABS(x); } while (0)
  for (j = 1; j <= nlat; j++)
    for (i = 1; i <= nlon; i++) {
      w = A2(s->iu, i, j) * m->full_cos_lat[j];
      ACC(0, A2(t->u_adv_lon, i, j) * w);
      ACC(1, A2(t->u_adv_lat, i, j) * w);
      ACC(2, A2(t->fv, i, j) * w);
      ACC(3, A2(t->u_pgf, i, j) * w);
    }
  for (j = 1; j <= nlat - 1; j++)
    for (i = 1; i <= nlon; i++) {
      w = A2(s->iv, i, j) * m->half_cos_lat[j];
      ACC(4, A2(t->v_adv_lon, i, j) * w);
      ACC(5, A2(t->v_adv_lat, i, j) * w);
      ACC(6, A2(t->fu, i, j) * w);
      ACC(7, A2(t->v_pgf, i, j) * w);
    }
  for (j = 1; j <= nlat; j++)
    for (i = 1; i <= nlon; i++) {
      w = (A2(s->gd, i, j) + A2(m->ghs, i, j)) * m->full_cos_lat[j];
      ACC(8, A2(t->mass_div_lon, i, j) * w);
      ACC(9, A2(t->mass_div_lat, i, j) * w);
    }
#undef ACC
  sums[0] = (double)(ip[0] + ip[4] + ip[1] + ip[5]);
  sums[1] = (double)(ip[3] + ip[8]);
  sums[2] = (double)(ip[2] - ip[6]);
  sums[3] = (double)(ip[7] + ip[9]);
  sums[4] = (double)(ab[0] + ab[4] + ab[1] + ab[5]);
  sums[5] = (double)(ab[3] + ab[8]);
  sums[6] = (double)(ab[2] + ab[6]);
  sums[7] = (double)(ab[7] + ab[9]);
  return 0;
}

int orc_update_state_preview(orc_model *m, double dt, double *u, double *v, double *gd, double *iu,
                             double *iv, double *igd) {
  state_t *n = STATE(m, m->new_idx);
  update_state(m, (real)dt, TEND(m, m->old_idx), STATE(m, m->old_idx), n);
  get_full(m, n->u, u); get_half(m, n->v, v); get_full(m, n->gd, gd);
  get_full(m, n->iu, iu); get_half(m, n->iv, iv); get_full(m, n->igd, igd);
  return 0;
}

int orc_predict_correct(orc_model *m, double dt, int pass) {
  int tmp;
  if (!m->run_inited) { snprintf(g_err, sizeof g_err, "orc_run_init not called"); return 2; }
  predict_correct(m, (real)dt, m->old_idx, m->new_idx, pass);
  tmp = m->old_idx; m->old_idx = m->new_idx; m->new_idx = tmp;
  return 0;
}

int orc_ordinary_diffusion(orc_model *m, double dt) {
  if (!m->run_inited) { snprintf(g_err, sizeof g_err, "orc_run_init not called"); return 2; }
  ordinary_diffusion(m, (real)dt, STATE(m, m->old_idx));
  return 0;
}

int orc_get_table(const orc_model *m, int which, double *out) {
  const real *t;
  int n = m->nlat, j;
  switch (which) {
    case 0: t = m->full_cos_lat; break;
    case 1: t = m->half_cos_lat; n = m->nlat - 1; break;
    case 2: t = m->full_f; break;
    case 3: t = m->full_c; break;
    case 4: t = m->full_dlon; break;
    case 5: t = m->half_dlon; n = m->nlat - 1; break;
    case 6: t = m->full_dlat; break;
    case 7: t = m->half_dlat; n = m->nlat - 1; break;
    case 8: t = m->full_lat; break;
    case 9: t = m->half_lat; n = m->nlat - 1; break;
    default: return 2;
  }
  for (j = 1; j <= n; j++) out[j - 1] = (double)t[j];
  return 0;
}

int orc_get_filter_rows(const orc_model *m, int *full_flag, int *full_cutoff, int *half_flag, int *half_cutoff) {
  int j;
  for (j = 1; j <= m->nlat; j++) {
    if (full_flag) full_flag[j - 1] = m->filter_full[j];
    if (full_cutoff) full_cutoff[j - 1] = m->cutoff_full[j];
  }
  for (j = 1; j <= m->nlat - 1; j++) {
    if (half_flag) half_flag[j - 1] = m->filter_half[j];
    if (half_cutoff) half_cutoff[j - 1] = m->cutoff_half[j];
  }
  return 0;
}

int orc_filter_row(orc_model *m, int half, int row0, double *x) {
  real *t = (real *)malloc(sizeof(real) * (size_t)m->nlon);
  int i, c = half ? m->cutoff_half[row0 + 1] : m->cutoff_full[row0 + 1];
  for (i = 0; i < m->nlon; i++) t[i] = (real)x[i];
  filter_row_real(m, c, t);
  for (i = 0; i < m->nlon; i++) x[i] = (double)t[i];
  free(t);
  return 0;
}

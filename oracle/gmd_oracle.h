/*
 * oracle/gmd_oracle.h -- C API of the CPU oracle (a sweep-by-sweep restatement of the reference's
 * barotropic shallow-water time step; see gmd_oracle.c for the file:line map).
 *
 * TEST INFRASTRUCTURE ONLY.  Loaded through ctypes by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  The product (include/gmd.h, libgmd.so) never
 * includes, links or calls anything declared here.
 *
 * PARITY UNPINNED: the reference ships no golden vector, known-answer test or fixture for this path
 *  and cannot be compiled in this image (Fortran only; no gfortran, no netCDF).
 * The oracle is pinned instead by (1) an independent NumPy restatement in tests/np_restatement.py,
 * (2) the invariants the scheme guarantees (mass/energy conservation, operator antisymmetry,
 * steady-state stationarity), (3) a binary128 build of this same source (-DORC_QUAD) and (4) known answers of
 * the continuous problem: second-order convergence to the steady geostrophic solution, the analytic
 * Rossby-Haurwitz phase speed, the balanced jet without its bump (tests/test_oracle.py).
 *
 * Array exchange format ("compact"): C-contiguous, longitude fastest, no halos, 0-based.
 *   full-lat fields  u, gd, ghs, du, dgd, div : [num_lat][num_lon]
 *   half-lat fields  v, dv, vor               : [num_lat-1][num_lon]
 * Staggering (mesh_mod.F90:60-80): u(i,j) is east of gd(i,j); v(i,j) is north of gd(i,j).
 */
#ifndef GMD_ORACLE_H
#define GMD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_SPLIT_NONE = 0, ORC_SPLIT_CSP1 = 1, ORC_SPLIT_CSP2 = 2, ORC_SPLIT_ISP = 3 };
enum { ORC_ADV_CENTER_DIFF = 0, ORC_ADV_UPWIND = 1, ORC_ADV_WENO = 2 };
enum { ORC_PASS_ALL = 0, ORC_PASS_FAST = 1, ORC_PASS_SLOW = 2 };
enum {
  ORC_IC_ROSSBY_HAURWITZ = 0,
  ORC_IC_STEADY_GEOSTROPHIC = 1,
  ORC_IC_MOUNTAIN_ZONAL = 2,
  ORC_IC_JET_ZONAL = 3,
  ORC_IC_SHALLOW_WATER_WAVES = 4
};

/* mirrors /dycore_params/ (params_mod.F90:13-98), numeric keys only */
typedef struct orc_config {
  int num_lon;
  int num_lat;
  int subcycles;               /* default 4 */
  double time_step_size;
  int qcon_modified;
  int split_scheme;            /* ORC_SPLIT_* */
  int uv_adv_scheme;           /* ORC_ADV_* */
  double uv_adv_upwind_lon_beta; /* default 0.0 */
  double uv_adv_upwind_lat_beta; /* default 0.5 */
  int use_zonal_tend_filter;   /* default 1 */
  int zonal_tend_filter_cutoff_wavenumber[20];
  int use_diffusion;
  int diffusion_order;         /* default 2 */
  double diffusion_coef;
  /* time_scheme (params_mod.F90:40-44 lists 'predict-correct' and 'runge-kutta'; the commit implements the first only,
     dycore_mod.F90:78-83).  ORC_TIME_RUNGE_KUTTA is the SPECIFIED extension of DESIGN.md section 8: parity unpinned. */
  int time_scheme;             /* ORC_TIME_* ; default predict_correct */
  int time_order;              /* runge_kutta: 3 (SSP-RK3, default) or 4 (classical RK4) */
  /* moving reduced tendency (README.md:13, run/namelist.jz_test:16-19; absent from the commit: SPECIFIED extension of
     DESIGN.md section 8, parity unpinned) */
  int use_zonal_reduce;
  int reduce_adv_lon;
  int use_reduce_tend_smooth;
  int zonal_reduce_factors[20];
} orc_config;
enum { ORC_TIME_PREDICT_CORRECT = 0, ORC_TIME_RUNGE_KUTTA = 1 };

typedef struct orc_model orc_model;
typedef struct orc_rfft_plan orc_rfft_plan;

/* dycore_init (dycore_mod.F90:60-111): mesh_init, parallel_init, data_init, filter_init ... */
int orc_create(const orc_config *cfg, orc_model **out);
void orc_destroy(orc_model *m);
const char *orc_last_error(void);

/* test-case plugins (src/test_cases/barotropic/...): write u,v,gd into state(1), ghs into static.
   params: RH -> {R, omg, gd0} (NULL = defaults); mountain -> {smooth_mountain}. */
int orc_set_initial_condition(orc_model *m, int test_case, const double *params, int nparams);
/* phase speed (rad/s) of the Shamir-Paldor wave: wave_flag -1 WIG, 0 Rossby, 1 EIG (shallow_water_waves_test_mod.F90:130-171) */
double orc_swe_phase_speed(int wave_flag);
/* arbitrary state in compact layout; fills the periodic lon halos as the plugins do */
int orc_set_state(orc_model *m, const double *u, const double *v, const double *gd, const double *ghs);

/* head of dycore_run (dycore_mod.F90:121-129): reset_cos_lat_at_poles, iap_transform, diag_run */
int orc_run_init(orc_model *m);
/* nsteps x { time_integrate; time_advance; diag_run } (dycore_mod.F90:131-140).
   Returns 0, or 1 when total mass / energy became NaN (diag_mod.F90:79-87). */
int orc_step(orc_model *m, int nsteps);

/* current state(old) in compact layout; any pointer may be NULL */
int orc_get_state(const orc_model *m, double *u, double *v, double *gd);
int orc_get_iap_state(const orc_model *m, double *iap_u, double *iap_v, double *iap_gd);
int orc_get_ghs(const orc_model *m, double *ghs);
int orc_get_diag(const orc_model *m, double *total_mass, double *total_energy, double *beta);
int orc_get_vor_div(const orc_model *m, double *vor, double *div);
int orc_get_step_count(const orc_model *m);

/* one space_operators(state(old), tend, pass) evaluation (dycore_mod.F90:184-365); outputs the
   combined (filtered) tendencies.  orc_run_init must have been called. */
int orc_space_operators(orc_model *m, int pass, double *du, double *dv, double *dgd);
/* the ten separate terms of the last orc_space_operators call, for check_antisymmetry
   (dycore_mod.F90:794-851): sums[0..3] = the four printed sums, sums[4..7] = their scale
   (sum of absolute values of the summands). */
int orc_check_antisymmetry(orc_model *m, double *sums);
/* one update_state(dt, tend, state(old) -> state(new)) with the tendencies of the last
   orc_space_operators call; returns the new state without committing it. */
int orc_update_state_preview(orc_model *m, double dt, double *u, double *v, double *gd,
                             double *iap_u, double *iap_v, double *iap_gd);
/* one predict_correct(dt, old -> new, pass) followed by a swap of old/new, no diag */
int orc_predict_correct(orc_model *m, double dt, int pass);
/* one ordinary_diffusion(dt, state(old)) */
int orc_ordinary_diffusion(orc_model *m, double dt);

/* mesh / coefficient tables after orc_run_init (i.e. with reset pole cosines); n = num_lat
   for full tables, num_lat-1 for half tables.  which: 0 full_cos_lat 1 half_cos_lat 2 full_f
   3 full_c 4 full_dlon 5 half_dlon 6 full_dlat 7 half_dlat 8 full_lat 9 half_lat */
int orc_get_table(const orc_model *m, int which, double *out);
/* filter row map after filter_init: flag (0/1) and effective cutoff (-1 = all-zero mask) per row */
int orc_get_filter_rows(const orc_model *m, int *full_flag, int *full_cutoff, int *half_flag,
                        int *half_cutoff);
/* filter_array_at_full_lat / _half_lat on one compact row (filter_mod.F90:105-167) */
int orc_filter_row(orc_model *m, int half, int row0, double *x);

/* jet_zonal_flow balanced gd profile (jet_zonal_flow_test_mod.F90:60-70) at latitude `lat` */
double orc_jet_gd_profile(double lat);

/* FFTPACK restatement (fftpack_rfft.c) */
int orc_rfft_plan_create(int n, orc_rfft_plan **out);
void orc_rfft_plan_destroy(orc_rfft_plan *p);
int orc_rfft_plan_factors(const orc_rfft_plan *p, int *fac, int maxfac);
int orc_rfft_forward_f64(int n, double *x);
int orc_rfft_backward_f64(int n, double *x);
int orc_rfft_factors(int n, int *fac, int maxfac);

/* sizeof(real) of this build: 8 (binary64) or 16 (binary128) */
int orc_real_bytes(void);

#ifdef __cplusplus
}
#endif
#endif

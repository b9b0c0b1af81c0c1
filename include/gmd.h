/*
 * include/gmd.h -- C ABI of the B200-native barotropic shallow-water time step ("gmd" = gamil dycore).
 *
 * This is the drop-in boundary for the ONE hot path of dongli/gamil-dycore that this repository
 * rebuilds: everything `dycore_run` executes between two output alerts.  The reference exposes that
 * path as four parameterless Fortran subroutines that talk through module-global data
 *     dycore_init / dycore_restart / dycore_run / dycore_final      (src/dycore_mod.F90:22-25,60-157)
 * with inputs in the namelist globals (src/params_mod.F90:13-98) and in state(old)%{u,v,gd}, static%ghs
 * (src/data_mod.F90:19-22).  A Fortran `dycore_mod` keeps those four names and calls the functions below
 * through ISO_C_BINDING (fortran/gmd_c.F90); the C++ host driver in
 * gamil_dycore_b200/host/ does the same where no Fortran compiler exists.
 *
 * Plain C: pointers and sizes only, no CUDA / torch types.  All arithmetic is IEEE binary64 on the GPU.
 * There is NO CPU fallback: every compute entry point returns GMD_ERR_CUDA when no device is usable.
 *
 * Array exchange layouts (host memory, borrowed for the duration of the call):
 *   GMD_LAYOUT_COMPACT    C-contiguous, longitude fastest, no halos, 0-based:
 *                         full-lat fields (u, gd, ghs, du, dgd, div) [num_lat][num_lon];
 *                         half-lat fields (v, dv, vor)               [num_lat-1][num_lon].
 *   GMD_LAYOUT_REFERENCE  the reference's own allocation (parallel_allocate_2, src/parallel_mod.F90:283-315):
 *                         column-major (lon fastest) with a 2-wide halo in both directions, i.e.
 *                         [num_lat+4][num_lon+4] doubles for every field, interior at [j+1][i+1] for the
 *                         1-based Fortran (i,j).  A `state_type` component can be passed with c_loc.
 *                         On output the periodic longitude halos are filled as parallel_fill_halo does
 *                         (src/parallel_mod.F90:466-524); latitude halos are left untouched (zero).
 * Staggering (src/mesh_mod.F90:60-80): u(i,j) is east of gd(i,j); v(i,j) is north of gd(i,j).
 *
 * Multi-GPU: one process per GPU.  Rank r of n owns a contiguous band of full latitude rows
 * (gmd_get_band); set/get calls always take GLOBAL arrays, each rank reads / fills only its own rows.
 */
#ifndef GMD_H
#define GMD_H

#ifdef __cplusplus
extern "C" {
#endif

#define GMD_VERSION 100
/* |u| on a pole row up to this bound (m/s) is accepted by gmd_set_state and stored as the 0 it stands for (the
   Shamir-Paldor wave IC evaluates ~1e-146 there, shallow_water_waves_test_mod.F90:215-271) */
#define GMD_POLE_U_TINY 1.0e-100

enum { GMD_OK = 0, GMD_ERR_NAN = 1, GMD_ERR_ARG = 2, GMD_ERR_CUDA = 3, GMD_ERR_COMM = 4, GMD_ERR_STATE = 5 };

/* split_scheme (src/dycore_mod.F90:27-34,85-95); csp1 is parsed but runs unsplit (time_integrate :654-669) */
enum { GMD_SPLIT_NONE = 0, GMD_SPLIT_CSP1 = 1, GMD_SPLIT_CSP2 = 2, GMD_SPLIT_ISP = 3 };
/* uv_adv_scheme (src/dycore_mod.F90:39-42,97-107) */
enum { GMD_ADV_CENTER_DIFF = 0, GMD_ADV_UPWIND = 1, GMD_ADV_WENO = 2 };
/* pass (src/dycore_mod.F90:35-37) */
enum { GMD_PASS_ALL = 0, GMD_PASS_FAST = 1, GMD_PASS_SLOW = 2 };
enum { GMD_LAYOUT_COMPACT = 0, GMD_LAYOUT_REFERENCE = 1 };

/* Numeric keys of namelist /dycore_params/ (src/params_mod.F90:13-98) plus the decomposition. */
typedef struct gmd_config {
  int num_lon;                     /* params_mod.F90:21 */
  int num_lat;                     /* params_mod.F90:22 */
  int subcycles;                   /* :23, default 4 */
  double time_step_size;           /* :33 */
  int qcon_modified;               /* :44 */
  int split_scheme;                /* GMD_SPLIT_*  (:47) */
  int uv_adv_scheme;               /* GMD_ADV_*    (:52) */
  double uv_adv_upwind_lon_beta;   /* :55, default 0.0 */
  double uv_adv_upwind_lat_beta;   /* :56, default 0.5 */
  int use_zonal_tend_filter;       /* :65, default 1 */
  int zonal_tend_filter_cutoff_wavenumber[20]; /* :66, default 0 */
  int use_diffusion;               /* :58, default 0 */
  int diffusion_order;             /* :59, default 2 (2 or 4) */
  double diffusion_coef;           /* :60 */
  /* decomposition / placement (new: the reference's parallel_mod is a serial stub, parallel_mod.F90:75-113) */
  int rank;                        /* 0 .. nranks-1 */
  int nranks;                      /* latitude bands; 1 = whole globe */
  int device;                      /* CUDA device ordinal, -1 = current device */
  int polar_band_rows;             /* nranks >= 3: rows of the first and the last band (they also carry the polar
                                      filter rows and pole caps, whose cost does not shrink with the band), the
                                      other ranks share the rest evenly; 0 = all bands even (default) */
  /* time_scheme (params_mod.F90:40-44 names 'predict-correct' and 'runge-kutta'; the reference commit implements the
     first only, dycore_mod.F90:78-83).  GMD_TIME_RUNGE_KUTTA is the specified extension of DESIGN.md section 8. */
  int time_scheme;                 /* GMD_TIME_*, default predict_correct */
  int time_order;                  /* params_mod.F90:45; runge_kutta: 3 = SSP-RK3 (default), 4 = classical RK4 */
  /* moving reduced tendency near the poles (README.md:13, run/namelist.jz_test:16-19; absent from the reference
     commit: specified in DESIGN.md section 8) */
  int use_zonal_reduce;            /* default 0 */
  int reduce_adv_lon;              /* the slow (advection) pass is reduced too; default 0 */
  int use_reduce_tend_smooth;      /* inner-product rescale of the reduced rows, as the filter blocks do; default 0 */
  int zonal_reduce_factors[20];    /* factor of the k-th row from each pole; 0 or 1 = not reduced */
} gmd_config;
enum { GMD_TIME_PREDICT_CORRECT = 0, GMD_TIME_RUNGE_KUTTA = 1 };

typedef struct gmd_model gmd_model;

/* reference defaults for every key that has one (params_mod.F90:13-66); rank 0 of 1, device -1 */
void gmd_config_defaults(gmd_config *cfg);

/* dycore_init (src/dycore_mod.F90:60-111): mesh_init, parallel_init, data_init, filter_init,
   diffusion_init, weno_init -- builds the coefficient tables, the filter row map and all device buffers. */
int gmd_create(const gmd_config *cfg, gmd_model **out);
/* dycore_final (src/dycore_mod.F90:144-157) */
void gmd_destroy(gmd_model *m);
/* message of the last failing call on this thread (log_error text, src/log_mod.F90:66-79) */
const char *gmd_last_error(void);
int gmd_version(void);

/* NCCL bootstrap for nranks > 1: rank 0 calls gmd_comm_unique_id, the host program distributes the 128
   bytes (MPI_Bcast / torch.distributed / a file), every rank calls gmd_comm_init.  No reference
   counterpart (parallel_init, src/parallel_mod.F90:75-113, is serial). */
int gmd_comm_unique_id(void *id128);
int gmd_comm_init(gmd_model *m, const void *id128);

/* Peer-memory data path for nranks > 1 (one process per GPU on one NVLink / NVSwitch node): every rank exports a
   blob describing its field slab and its signal page (CUDA IPC handles), the host program gathers the blobs of
   all ranks in rank order (MPI_Allgather / torch.distributed) and every rank calls gmd_peer_connect.  From then
   on halo rows are stored straight into the neighbour's ghost rows by this rank's kernels, the neighbour's next
   boundary launch spins on a release/acquire flag, and the two-scalar all-reduces are one-shot peer exchanges --
   NCCL is no longer on the step path (gmd_comm_init becomes optional).  No reference counterpart.
   All ranks must issue the same sequence of gmd_* calls after connecting (the slot index of a buffer is what
   identifies "the same buffer" on a neighbour, and gmd_step / gmd_predict_correct / gmd_ordinary_diffusion wait for
   the neighbours inside their kernels). */
#define GMD_PEER_BLOB_BYTES 256
int gmd_peer_export(gmd_model *m, void *blob);
int gmd_peer_connect(gmd_model *m, const void *blobs, int nblobs);
/* back to "not connected" (unmaps the neighbours): for a host program that wants to fall back to gmd_comm_init
   when gmd_peer_connect failed on some rank.  Collective in the same sense as gmd_peer_connect. */
int gmd_peer_disconnect(gmd_model *m);
/* full-latitude rows [row_begin, row_end) (0-based) owned by this rank */
int gmd_get_band(const gmd_model *m, int *row_begin, int *row_end);
/* rows [row_begin, row_end) of this rank whose predict_correct runs as the fused wavefront kernel k_pc (the rest --
   filter / reduced / pole rows and the plain rows next to them -- runs the three-sweep k_stage + k_polar chain);
   row_begin == row_end == 0 when the configuration does not use it (WENO, runge_kutta, the strict build, polar bands of
   a multi-band run, GMD_FUSED=0).  Informational: results do not depend on it beyond rounding of the inner products. */
int gmd_get_fused_rows(const gmd_model *m, int *row_begin, int *row_end);

/* The IC plugins / restart_read write state(old)%{u,v,gd} and static%ghs (e.g.
   rossby_haurwitz_wave_test_mod.F90:45-83); this uploads them.  ghs may be NULL (= 0).
   u on the two pole rows must be 0 (true for every reference IC; |u| <= GMD_POLE_U_TINY is taken as 0; see
   DESIGN.md "pole rows"). */
int gmd_set_state(gmd_model *m, const double *u, const double *v, const double *gd, const double *ghs,
                  int layout);
/* head of dycore_run (src/dycore_mod.F90:121-129): reset_cos_lat_at_poles, iap_transform, diag_run */
int gmd_run_init(gmd_model *m);
/* nsteps x { time_integrate; time_advance; diag_run } (src/dycore_mod.F90:131-140), batched between
   output alerts.  Returns GMD_ERR_NAN when total mass or energy became NaN (src/diag_mod.F90:79-87). */
int gmd_step(gmd_model *m, int nsteps);

/* state(old) (u, v, gd) / its IAP transform (U, V, sqrt(gd)); any pointer may be NULL */
int gmd_get_state(gmd_model *m, double *u, double *v, double *gd, int layout);
int gmd_get_iap_state(gmd_model *m, double *iap_u, double *iap_v, double *iap_gd, int layout);
/* diag%total_mass, diag%total_energy (src/diag_mod.F90:71-77,98-121) and the last beta
   (src/dycore_mod.F90:786-787) after the most recent step */
int gmd_get_diag(gmd_model *m, double *total_mass, double *total_energy, double *beta);
/* the same three scalars for the last n steps (oldest first); n <= steps since gmd_run_init, n <= 4096 */
int gmd_get_diag_series(gmd_model *m, int n, double *total_mass, double *total_energy, double *beta);
/* diag%vor (half rows) and diag%div (full rows, pole rows 0) of state(old) (src/diag_mod.F90:47-69) */
int gmd_get_vor_div(gmd_model *m, double *vor, double *div, int layout);
int gmd_get_step_count(const gmd_model *m);

/* ---- finer-grained entry points: the internal routines of dycore_mod, exposed for parity tests ---- */
/* space_operators(state(old), tend, pass) (src/dycore_mod.F90:184-365): combined, filtered tendencies */
int gmd_space_operators(gmd_model *m, int pass, double *du, double *dv, double *dgd, int layout);
/* predict_correct(dt, old -> new, pass) (src/dycore_mod.F90:754-792) followed by the old/new swap */
int gmd_predict_correct(gmd_model *m, double dt, int pass);
/* ordinary_diffusion(dt, state(old)) (src/diffusion_mod.F90:74-217) */
int gmd_ordinary_diffusion(gmd_model *m, double dt);
/* filter_array_at_full_lat / _half_lat on one row of num_lon values (src/filter_mod.F90:105-167) */
int gmd_filter_row(gmd_model *m, int half, int row0, double *x);
/* filter row map after filter_init (src/filter_mod.F90:35-103): flag and effective cutoff (-1 = none) */
int gmd_get_filter_rows(const gmd_model *m, int *full_flag, int *full_cutoff, int *half_flag,
                        int *half_cutoff);
/* coefficient tables after gmd_run_init; `which` as in oracle: 0 full_cos_lat 1 half_cos_lat 2 full_f
   3 full_c 4 full_dlon 5 half_dlon 6 full_dlat 7 half_dlat 8 full_lat 9 half_lat */
int gmd_get_table(const gmd_model *m, int which, double *out);

/* ---- execution control / measurement ---- */
/* run all work on this cudaStream_t (e.g. the host framework's current stream); NULL = own stream */
int gmd_set_stream(gmd_model *m, void *cuda_stream);
/* replay each model step from a captured CUDA graph (default on); 0 = launch kernels directly */
int gmd_set_graph_mode(gmd_model *m, int on);
/* like gmd_step but asynchronous: enqueues nsteps on the stream and returns; no NaN check.
   gmd_sync waits and performs the check. */
int gmd_step_async(gmd_model *m, int nsteps);
int gmd_sync(gmd_model *m);
/* device time of the last `gmd_step`/`gmd_step_async`+`gmd_sync` span, CUDA events on the model's stream */
int gmd_last_step_ms(gmd_model *m, float *ms);
/* kernels launched by this model so far (graph replays count their kernel nodes) */
long long gmd_kernel_launches(const gmd_model *m);
/* algorithmic HBM bytes per grid column per model step for this configuration:
   312 n_fast + 216 n_slow + 48 [diffusion] */
double gmd_algorithmic_bytes_per_column_step(const gmd_model *m);
/* time the dominant kernel (the fused stage kernel, S2 variant of the configured fast/all pass) alone:
   `reps` back-to-back launches between CUDA events on the model's stream; returns average ms per
   launch and the algorithmic bytes one launch moves on this rank */
int gmd_time_stage_kernel(gmd_model *m, int reps, float *ms_per_launch, double *alg_bytes_per_launch);
/* the same for any (pass, mode) instantiation: mode 0 = S1 (predict), 1 = S2 (predict + store tendency),
   2 = S3a (tendency + inner products), 3 = evaluation only, 4 = S1 with the deferred update folded in,
   5 = the fused predict_correct kernel k_pc (S1 + S2 + S3a as one wavefront, deferred update folded in) over the rows
   it covers in a step; GMD_ERR_STATE when the configuration does not use it.  alg_bytes of mode 5: the 13 (slow pass:
   11) words per column the kernel has to move, not the 36 of the three sweeps it replaces */
int gmd_time_stage_variant(gmd_model *m, int pass, int mode, int reps, float *ms_per_launch,
                           double *alg_bytes_per_launch);
/* Device timeline of the following model steps (libgmd_trace.so, the -DGMD_TRACE=1 build of the same sources; the
   product build returns GMD_ERR_STATE): every kernel launch of a step records the %globaltimer of its first CTA's
   start, its last CTA's end and the longest in-kernel wait for a neighbour rank.  gmd_trace_dump writes the last
   (up to 4) steps as JSON: {"steps": [{"step": n, "launches": [{"seq", "kernel", "start_ns", "end_ns", "wait_ns"}]}]} */
int gmd_trace_begin(gmd_model *m);
int gmd_trace_dump(gmd_model *m, const char *path);

#ifdef __cplusplus
}
#endif
#endif

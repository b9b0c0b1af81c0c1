// gmd.cu -- model object, step driver and C ABI (include/gmd.h) of the B200-native barotropic step.
//
// Control flow mirrors src/dycore_mod.F90: time_integrate (:654-669) -> csp2_splitting (:671-687) /
// isp_splitting (:689-752) / predict_correct (:754-792) -> space_operators + update_state, followed by
// ordinary_diffusion (src/diffusion_mod.F90:74-217) and diag_run (src/diag_mod.F90:42-89).  Everything
// between two host calls is stream-ordered on one CUDA stream; beta never leaves the device.
#include "../../include/gmd.h"
#include "gmd_kernels.cuh"
#if !GMD_STRICT
#include "gmd_pc.cuh"
#endif
#include "gmd_mesh.h"

#include <dlfcn.h>
#include <unistd.h>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <vector>

using namespace gmd;

// ---------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fail(GMD_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, \
                  cudaGetErrorString(e_));                                                               \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// NCCL, bound at run time (only when nranks > 1) so that the library loads on hosts without it
// ---------------------------------------------------------------------------------------------------------
struct NcclId128 {  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value
  char b[128];
};
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, NcclId128, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.lib) return 0;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(GMD_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                        \
  *(void **)(&g_nccl.field) = dlsym(h, name);                                   \
  if (!g_nccl.field) return fail(GMD_ERR_COMM, "libnccl lacks symbol %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.lib = h;
  return 0;
}
#define NK(call)                                                                                              \
  do {                                                                                                        \
    int r_ = (call);                                                                                          \
    if (r_ != 0) return fail(GMD_ERR_COMM, "NCCL error %d (%s) at %s:%d", r_, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
  } while (0)
enum { NCCL_F64 = 8, NCCL_SUM = 0 };

// ---------------------------------------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------------------------------------
enum { KIND_U = 0, KIND_V = 1, KIND_G = 2 };

struct State {
  double *U = nullptr, *V = nullptr, *gd = nullptr;
};
struct Tend {
  double *U = nullptr, *V = nullptr, *gd = nullptr;
};

struct gmd_model {
  gmd_config cfg;
  HostMesh mesh;
  Geo geo;
  int nr = 0;            // owned rows
  size_t fld_elems = 0;  // (nr + 2 GHOST) * nlon
  int dev = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool run_inited = false, have_state = false;
  int step = 0;
  long long launches = 0;

  // device tables
  std::vector<double *> tab_allocs;
  unsigned char *d_flags_alloc = nullptr;
  Tab tab;
  double *d_basis = nullptr;
  int ncoef_max = 0;

  // polar items: [0] all/fast pass, [1] slow pass, [2] plain filter (diffusion)
  std::vector<unsigned> items[3];  // packed (kind, row, cutoff), passed in the kernel parameters
  int n_items[3] = {0, 0, 0};
  double *d_rot = nullptr;         // [PQ][KF][2] rotation table of the fast projector

  // buffers: every field lives in ONE slab (fixed slots), so that one CUDA-IPC handle maps the whole pool into
  // the neighbour ranks and "my buffer X" and "the neighbour's buffer X" are the same slot index
  std::vector<double *> free_[3];
  double *slab = nullptr;
  int slab_cap = 0, slab_next = 0;
  bool ew_rows = false;    // row-pair form of the diffusion / derive sweeps
  unsigned ew2_bx = 1;     // its CTAs per row
  std::map<double *, int> refc;  // base pointer (row r0) -> refcount
  std::map<double *, int> kind_of;

  State cur;
  double *ghs = nullptr;
  Tend tendOld, tendNew, tendNew2, tendA, tendB;  // tendA/B: isp only
  // (also with the fused predict_correct kernel on one band: its S3a warps write the new tendency while the S1 warps of
  // other CTAs still read the previous one -- the deferred update -- so the two must be different buffers)
  // the new tendency of predict_correct alternates between tendNew and tendNew2: a neighbour band stores the ghost
  // rows of predict_correct k+1 while this band may still be reading those of k (deferred update)
  int tn_idx = 0;
  double *d_partials = nullptr;
  int n_partials = 0;
  double *d_ip = nullptr;    // {ip1, ip2}
  double *d_sums = nullptr;  // {mass, energy}
  double *d_beta = nullptr;
  double *d_ring = nullptr;  // [RING][3]
  int *d_ctr = nullptr;      // device step counter (ring slot = ctr % RING)
  static const int RING = 4096;
  // weno / diffusion scratch (lazily acquired, kept)
  double *w_u = nullptr, *w_v = nullptr;
  double *w_fpu = nullptr, *w_fnu = nullptr, *w_fpv = nullptr, *w_fnv = nullptr, *w_fu = nullptr, *w_fv = nullptr;
  double *w_alon_u = nullptr, *w_alat_u = nullptr, *w_alon_v = nullptr, *w_alat_v = nullptr;
  double *d_ud = nullptr, *d_vd = nullptr, *d_gdd = nullptr, *d_ud2 = nullptr, *d_vd2 = nullptr, *d_gdd2 = nullptr;

  // stage launch geometry
  int nbx = 0, nchunks = 0, rows_per_cta = 0;
  int rows_per_cta_s3a = 0;   // S3a launches: their own one-wave geometry (the S3a kernels may be capped at fewer CTAs per SM)
  size_t stage_smem = 0;  // dynamic shared memory of k_stage: (rows_per_cta + 2) row records
  // boundary / interior split: rows [r0, r0+bs) and [r1-bn, r1) are evaluated first on the
  // main stream, followed by the polar-row kernel and the halo exchange, while rows [r0+bs, r1-bn) run on stream2
  int bs = 0, bn = 0, rows_per_cta_b = 0, nchunks_b = 0, nchunks_i = 0;
  size_t stage_smem_b = 0;
  cudaStream_t stream2 = nullptr;
  std::vector<cudaEvent_t> evpool;
  size_t evnext = 0;
  cudaEvent_t last_eI = nullptr;  // completion of the last interior launch on stream2
  cudaEvent_t ev_polar_side = nullptr;  // completion of the last polar-side stage launch on the main stream
  bool split = true;
  int ew_blocks = 0;  // grid of element-wise kernels
  // fused polar cap (k_cap): the sweep over the rows next to a pole and the polar rows of that sweep in one launch
  bool cap = false;
  unsigned *d_bar = nullptr;   // grid barrier counters of k_cap
  int cap_ctas = 0;            // CTAs of a k_cap launch (co-resident; a multiple of the cluster size)
  int prio_hi = 0;             // launch priority of k_cap
  int pdl = 0;                 // GMD_PDL: bit 0: polar rows after their sweep, bit 1: the next sweep after the polar rows,
                               // bit 2: the stage chain of a band without polar rows
  // fused predict_correct (k_pc, gmd_pc.cuh): rows [fz_I0, fz_I1) run the three sweeps as one wavefront kernel on
  // stream2; the rows next to a pole -- [r0, fz_I0) and [fz_I1, r1), which hold the filter / reduced / pole rows plus
  // the (2, 4) plain rows the wavefront must keep away from them -- run the k_stage + k_polar chain on the main stream,
  // each sweep widened towards the fused rows by what the later sweeps of the chain read (nothing is exchanged)
  bool fused = false;
  int fz_I0 = 0, fz_I1 = 0, fz_rpc = 0, fz_nchunks = 0, fz_nstrips = 0;
  int fz_ncb = 0;              // row chunks of a (widened) polar-side launch
  bool fz_active = false;      // inside the three stage() calls of a fused predict_correct
  Fold fz_fold;                // the inner-product fold shared by k_pc and the S3a polar-side launches
  cudaEvent_t fz_ev = nullptr; // completion of the k_pc launch on stream2

  // comm: NCCL (optional) and the peer-memory path (gmd_peer_connect)
  void *comm = nullptr;
  u64 *page = nullptr;             // my signal page
  bool p2p = false;
  u64 *peer_page[MAXR] = {};       // every rank's signal page as mapped here
  double *peer_slab[2] = {nullptr, nullptr};   // south / north neighbour's slab as mapped here
  int peer_r0[2] = {0, 0};
  size_t peer_fld[2] = {0, 0};
  std::vector<void *> ipc_opened;
  unsigned xk = 0, rk = 0, xwaited = 0;  // halo epochs released / reductions done / halo epoch waited for, this unit
  bool fuse_push = false;     // band-edge rows of the new tendency are stored to the neighbours by the S3a launches
  // wide-halo predict_correct: the three sweeps run on rows shrinking by (1 south, 2 north) per
  // sweep, the bands exchange only the new tendency, once per predict_correct
  bool wide = false;
  // A wide-halo predict_correct synchronises the bands only through its all-reduce, so after it a band may still be
  // reading buffers (its k_update) that a faster neighbour has already released, re-acquired and would now store
  // ghost rows into.  The first generic halo push after such a predict_correct is therefore preceded by a
  // signal-only exchange: it completes once BOTH neighbours have reached the same point of their streams.
  bool need_fence = false;

  // graphs
  bool graph_mode = true;
  struct GraphEntry {
    std::vector<double *> key;
    cudaGraphExec_t exec;
    long long launches;
    State out;
  };
  std::vector<GraphEntry> graphs;
  bool capturing = false;
  bool dry = false;  // bookkeeping-only pass of the step logic (graph replay): no launches

  // host <-> device transfer lanes (pinned double buffers, one copy stream each)
  struct Lane {
    cudaStream_t s = nullptr;
    double *pin[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
  };
  Lane lanes[4];
  bool lanes_ready = false;
  int lane_rows = 0;

  float last_ms = 0.f;
  bool span_open = false;
  int pending_steps = 0;

  // device timeline (GMD_TRACE builds, gmd_trace_begin / gmd_trace_dump)
  bool in_step = false;
  long long step_launch0 = 0;
  u64 *d_trace = nullptr;
  std::vector<std::string> trace_names;
  int trace_step0 = 0;
  static const int TRACE_STEPS = 4, TRACE_PER_STEP = 768;
};

// timeline slot of the launch that is about to be issued: its position inside the model step (GMD_TRACE builds)
static inline int tseq(gmd_model *m, const char *name) {
#if GMD_TRACE
  if (!m->in_step) return 0;
  const long long q = m->launches - m->step_launch0;
  if (q < 0 || q >= gmd_model::TRACE_PER_STEP) return 0;
  if ((long long)m->trace_names.size() <= q) m->trace_names.resize((size_t)q + 1);
  m->trace_names[(size_t)q] = name;
  return (int)q + 1;   // 0 = no slot
#else
  (void)m; (void)name;
  return 0;
#endif
}

static int set_dev(gmd_model *m) {
  CK(cudaSetDevice(m->dev));
  return 0;
}

// rows [r0, r1) of band `rank` (mirrored by gamil_dycore_b200/parallel.py: band)
static void band_rows(int nlat, int nranks, int polar_rows, int rank, int *r0, int *r1) {
  if (nranks >= 3 && polar_rows > 0 && 2 * polar_rows < nlat) {
    const int mid = nlat - 2 * polar_rows, nm = nranks - 2;
    const int base = mid / nm, rem = mid % nm;
    if (rank == 0) { *r0 = 0; *r1 = polar_rows; return; }
    if (rank == nranks - 1) { *r0 = nlat - polar_rows; *r1 = nlat; return; }
    const int q = rank - 1;
    *r0 = polar_rows + q * base + std::min(q, rem);
    *r1 = *r0 + base + (q < rem ? 1 : 0);
    return;
  }
  const int base = nlat / nranks, rem = nlat % nranks;
  *r0 = rank * base + std::min(rank, rem);
  *r1 = *r0 + base + (rank < rem ? 1 : 0);
}

// ---- buffer pool --------------------------------------------------------------------------------------
static int acquire(gmd_model *m, int kind, double **out) {
  if (!m->free_[kind].empty()) {
    *out = m->free_[kind].back();
    m->free_[kind].pop_back();
  } else {
    if (m->slab_next >= m->slab_cap)
      return fail(GMD_ERR_STATE, "field pool exhausted (%d fields of %zu bytes); set GMD_POOL_FIELDS", m->slab_cap,
                  m->fld_elems * sizeof(double));
    double *p = m->slab + (size_t)m->slab_next++ * m->fld_elems;  // zero since gmd_create
    *out = p + (size_t)GHOST * m->geo.nlon;
    m->kind_of[*out] = kind;
  }
  m->refc[*out] = 1;
  return 0;
}
static void retain(gmd_model *m, double *p) { m->refc[p]++; }
static void release(gmd_model *m, double *p) {
  if (!p) return;
  if (--m->refc[p] == 0) m->free_[m->kind_of[p]].push_back(p);
}
static int new_state(gmd_model *m, State *s, double *shared_gd) {
  int r;
  if ((r = acquire(m, KIND_U, &s->U))) return r;
  if ((r = acquire(m, KIND_V, &s->V))) return r;
  if (shared_gd) {
    s->gd = shared_gd;
    retain(m, shared_gd);
  } else if ((r = acquire(m, KIND_G, &s->gd)))
    return r;
  return 0;
}
static void release_state(gmd_model *m, State *s) {
  release(m, s->U);
  release(m, s->V);
  release(m, s->gd);
  s->U = s->V = s->gd = nullptr;
}
static int new_tend(gmd_model *m, Tend *t) {
  int r;
  if ((r = acquire(m, KIND_U, &t->U))) return r;
  if ((r = acquire(m, KIND_V, &t->V))) return r;
  if ((r = acquire(m, KIND_G, &t->gd))) return r;
  return 0;
}

// ---- tables -------------------------------------------------------------------------------------------
static int upload_table(gmd_model *m, const std::vector<double> &h, const double **dptr) {
  double *d = nullptr;
  CK(cudaMalloc(&d, h.size() * sizeof(double)));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  m->tab_allocs.push_back(d);
  *dptr = d + TPAD;
  return 0;
}

static int build_tables(gmd_model *m) {
  HostMesh &M = m->mesh;
  const int nlat = M.nlat;
  const size_t n = (size_t)nlat + 2 * TPAD;
  auto mk = [&](double fill) { return std::vector<double>(n, fill); };
  std::vector<double> q_fdlon = mk(0), q_hdlon = mk(0), q_fdlat = mk(0), q_hdlat = mk(0), cor1 = mk(0), cor2 = mk(0),
                      r_fdlon = mk(0), hc_hdlat = mk(0), h_fdlon = mk(0), h_fdlat = mk(0);
  for (int j = 0; j < nlat; j++) {
    const size_t k = (size_t)(j + TPAD);
    q_fdlon[k] = 0.25 / M.full_dlon[k];
    q_fdlat[k] = 0.25 / M.full_dlat[k];
    r_fdlon[k] = 1.0 / M.full_dlon[k];
    h_fdlon[k] = 0.5 / M.full_dlon[k];
    h_fdlat[k] = 0.5 / M.full_dlat[k];
    cor1[k] = M.half_cos[k - 1] / M.full_cos[k];
    cor2[k] = M.half_cos[k] / M.full_cos[k];
    if (j < nlat - 1) {
      q_hdlon[k] = 0.25 / M.half_dlon[k];
      q_hdlat[k] = 0.25 / M.half_dlat[k];
      hc_hdlat[k] = M.half_cos[k] / M.half_dlat[k];
    }
  }
  int r;
  // padded entries of the divisor tables must not be 0 (they are multiplied by exact zeros, never used)
  std::vector<double> fdlon = M.full_dlon, hdlon = M.half_dlon, fdlat = M.full_dlat, hdlat = M.half_dlat;
  for (size_t k = 0; k < n; k++) {
    if (fdlon[k] == 0.0) fdlon[k] = 1.0;
    if (hdlon[k] == 0.0) hdlon[k] = 1.0;
    if (fdlat[k] == 0.0) fdlat[k] = 1.0;
    if (hdlat[k] == 0.0) hdlat[k] = 1.0;
  }
  Tab &t = m->tab;
  memset(&t, 0, sizeof t);
  if ((r = upload_table(m, M.full_cos, &t.cosf))) return r;
  if ((r = upload_table(m, M.half_cos, &t.cosh))) return r;
  if ((r = upload_table(m, M.full_f, &t.ff))) return r;
  if ((r = upload_table(m, M.full_c, &t.fc))) return r;
  if ((r = upload_table(m, fdlon, &t.fdlon))) return r;
  if ((r = upload_table(m, hdlon, &t.hdlon))) return r;
  if ((r = upload_table(m, fdlat, &t.fdlat))) return r;
  if ((r = upload_table(m, hdlat, &t.hdlat))) return r;
  if ((r = upload_table(m, q_fdlon, &t.q_fdlon))) return r;
  if ((r = upload_table(m, q_hdlon, &t.q_hdlon))) return r;
  if ((r = upload_table(m, q_fdlat, &t.q_fdlat))) return r;
  if ((r = upload_table(m, q_hdlat, &t.q_hdlat))) return r;
  if ((r = upload_table(m, cor1, &t.cor1))) return r;
  if ((r = upload_table(m, cor2, &t.cor2))) return r;
  if ((r = upload_table(m, r_fdlon, &t.r_fdlon))) return r;
  if ((r = upload_table(m, hc_hdlat, &t.hc_hdlat))) return r;
  if ((r = upload_table(m, h_fdlon, &t.h_fdlon))) return r;
  if ((r = upload_table(m, h_fdlat, &t.h_fdlat))) return r;

  // row flags
  std::vector<unsigned char> fl(n, 0);
  for (int j = 0; j < nlat; j++) {
    unsigned char f = 0;
    if (j >= 1 && j <= nlat - 2 && M.flag_full[(size_t)j]) f |= FL_DU | FL_DGD;
    if (j <= nlat - 2 && M.flag_half[(size_t)j]) f |= FL_DV;
    if (j >= 1 && j <= nlat - 2 && M.red_full[(size_t)j] > 1) f |= (FL_DU | FL_DGD) << FL_REDUCE_SHIFT;
    if (j <= nlat - 2 && M.red_half[(size_t)j] > 1) f |= FL_DV << FL_REDUCE_SHIFT;
    if (j == 0 || j == nlat - 1) f |= FL_POLE;
    fl[(size_t)(j + TPAD)] = f;
  }
  CK(cudaMalloc(&m->d_flags_alloc, n));
  CK(cudaMemcpy(m->d_flags_alloc, fl.data(), n, cudaMemcpyHostToDevice));
  t.flags = m->d_flags_alloc + TPAD;
  {  // packed per-row records for the stage kernel (one 128-byte line per row)
    std::vector<double> rec(n * RC_N, 0.0);
    for (size_t k = 0; k < n; k++) {
      double *r = &rec[k * RC_N];
      r[RC_COSF] = M.full_cos[k]; r[RC_COSH] = M.half_cos[k]; r[RC_FF] = M.full_f[k]; r[RC_FC] = M.full_c[k];
      r[RC_Q_FDLON] = q_fdlon[k]; r[RC_Q_FDLAT] = q_fdlat[k]; r[RC_Q_HDLON] = q_hdlon[k]; r[RC_Q_HDLAT] = q_hdlat[k];
      r[RC_COR1] = cor1[k]; r[RC_COR2] = cor2[k];
#if GMD_STRICT
      r[RC_PGFU] = fdlon[k]; r[RC_PGFV] = hdlat[k]; r[RC_MLON] = fdlon[k]; r[RC_MLAT] = fdlat[k];
#else
      r[RC_PGFU] = r_fdlon[k]; r[RC_PGFV] = hc_hdlat[k]; r[RC_MLON] = h_fdlon[k]; r[RC_MLAT] = h_fdlat[k];
#endif
      r[RC_FLAGS] = (double)fl[k];
    }
    const double *d = nullptr;
    if ((r = upload_table(m, rec, &d))) return r;
    t.rowrec = d - TPAD + (size_t)TPAD * RC_N;  // upload_table offsets by TPAD doubles; records need TPAD rows
  }

  // filter basis: row 0 = 1, row 2k-1 = cos(k x_i), row 2k = sin(k x_i), x_i = 2 pi i / nlon
  const int nlon = M.nlon;
  const int cmax = std::max(M.cutoff_max, 0);
  m->ncoef_max = 2 * (cmax + 1);
  {
    // one row more than the projector keeps (sin of the highest wavenumber): the fast path of k_polar rotates
    // (cos, sin) pairs
    const int nrows_b = m->ncoef_max + 1;
    std::vector<double> B((size_t)nrows_b * nlon);
    const long double twopi = 8.0L * atanl(1.0L);
    for (int mm = 0; mm < nrows_b; mm++) {
      const int k = (mm + 1) / 2;
      for (int i = 0; i < nlon; i++) {
        const long long rr = ((long long)k * i) % nlon;  // exact argument reduction
        const long double ang = twopi * (long double)rr / (long double)nlon;
        B[(size_t)mm * nlon + i] = (mm == 0) ? 1.0 : ((mm & 1) ? (double)cosl(ang) : (double)sinl(ang));
      }
    }
    CK(cudaMalloc(&m->d_basis, B.size() * sizeof(double)));
    CK(cudaMemcpy(m->d_basis, B.data(), B.size() * sizeof(double), cudaMemcpyHostToDevice));
    // rot[q][k-1] = (cos, sin)(2 pi k (q PT) / nlon): basis(i + q PT) = basis(i) rotated
    std::vector<double> R((size_t)PQ * KF * 2);
    for (int q = 0; q < PQ; q++)
      for (int k = 1; k <= KF; k++) {
        const long long rr = ((long long)k * q * PT) % nlon;
        const long double ang = twopi * (long double)rr / (long double)nlon;
        R[((size_t)q * KF + (k - 1)) * 2] = (double)cosl(ang);
        R[((size_t)q * KF + (k - 1)) * 2 + 1] = (double)sinl(ang);
      }
    CK(cudaMalloc(&m->d_rot, R.size() * sizeof(double)));
    CK(cudaMemcpy(m->d_rot, R.data(), R.size() * sizeof(double), cudaMemcpyHostToDevice));
  }

  // polar items for this band
  const int r0 = m->geo.r0, r1 = m->geo.r1;
  std::vector<unsigned> *it = m->items;
  for (int j = r0; j < r1; j++) {
    const bool fullrow = (j >= 1 && j <= nlat - 2);
    if (fullrow && M.flag_full[(size_t)j]) {
      const unsigned a = pack_item(IT_DU, j, M.cut_full[(size_t)j]);
      const unsigned g = pack_item(IT_DGD, j, M.cut_full[(size_t)j]);
      it[0].push_back(a);
      it[0].push_back(g);
      it[1].push_back(a);
      it[2].push_back(a);
      it[2].push_back(g);
    }
    if (j <= nlat - 2 && M.flag_half[(size_t)j]) {
      const unsigned v = pack_item(IT_DV, j, M.cut_half[(size_t)j]);
      it[0].push_back(v);
      it[1].push_back(v);
      it[2].push_back(v);
    }
    // rows of the moving reduced tendency (never also filter rows: gmd_create): fast / unsplit passes, the slow pass
    // with reduce_adv_lon, the diffusion tendencies always
    if (fullrow && M.red_full[(size_t)j] > 1) {
      const unsigned a = pack_reduce_item(IT_DU, j, M.red_full[(size_t)j]);
      const unsigned g = pack_reduce_item(IT_DGD, j, M.red_full[(size_t)j]);
      it[0].push_back(a);
      it[0].push_back(g);
      if (m->cfg.reduce_adv_lon) it[1].push_back(a);
      it[2].push_back(a);
      it[2].push_back(g);
    }
    if (j <= nlat - 2 && M.red_half[(size_t)j] > 1) {
      const unsigned v = pack_reduce_item(IT_DV, j, M.red_half[(size_t)j]);
      it[0].push_back(v);
      if (m->cfg.reduce_adv_lon) it[1].push_back(v);
      it[2].push_back(v);
    }
    if (j == 0) it[0].push_back(pack_item(IT_POLE_S, j, -1));
    if (j == nlat - 1) it[0].push_back(pack_item(IT_POLE_N, j, -1));
  }
  for (int k = 0; k < 3; k++) {
    m->n_items[k] = (int)it[k].size();
    if (m->n_items[k] > MAX_ITEMS) return fail(GMD_ERR_STATE, "internal: %d polar-row items exceed %d", m->n_items[k], MAX_ITEMS);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
typedef void (*stage_fn)(const StageArgs);
// dynamic shared memory of a stage launch: the row records of rows ja-1 .. jb and the CTA's packet ring
static inline size_t stage_smem_bytes(int rows_per_cta) {
  return (size_t)(rows_per_cta + 2) * RC_N * sizeof(double) + (GMD_RING ? RING_BYTES : 0);
}
// more than 48 KB of dynamic shared memory is opt-in per kernel
static int allow_smem(const void *fn) {
  static std::set<const void *> done;
  if (!fn || done.count(fn)) return 0;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  done.insert(fn);
  return 0;
}
// PUSH instantiations exist for MODE_S3A of the schemes a wide-halo predict_correct supports (not WENO)
template <int MODE, bool PUSH>
static stage_fn pick_stage_mode(int pass, int adv) {
  if (pass == PASS_FAST) return k_stage<PASS_FAST, ADV_CENTER, MODE, 0, PUSH>;
  if (pass == PASS_ALL) {
    if (adv == ADV_CENTER) return k_stage<PASS_ALL, ADV_CENTER, MODE, 0, PUSH>;
    if (adv == ADV_UPWIND) return k_stage<PASS_ALL, ADV_UPWIND, MODE, 0, PUSH>;
    if constexpr (PUSH) return nullptr;
    else return k_stage<PASS_ALL, ADV_WENO, MODE, 0, false>;
  }
  if (adv == ADV_CENTER) return k_stage<PASS_SLOW, ADV_CENTER, MODE, 0, PUSH>;
  if (adv == ADV_UPWIND) return k_stage<PASS_SLOW, ADV_UPWIND, MODE, 0, PUSH>;
  if constexpr (PUSH) return nullptr;
  else return k_stage<PASS_SLOW, ADV_WENO, MODE, 0, false>;
}
static stage_fn pick_stage(int pass, int adv, int mode, bool push = false) {
  if (push) return mode == MODE_S3A ? pick_stage_mode<MODE_S3A, true>(pass, adv) : nullptr;
  switch (mode) {
    case MODE_S1: return pick_stage_mode<MODE_S1, false>(pass, adv);
    case MODE_S2: return pick_stage_mode<MODE_S2, false>(pass, adv);
    case MODE_S3A: return pick_stage_mode<MODE_S3A, false>(pass, adv);
    default: return pick_stage_mode<MODE_EVAL, false>(pass, adv);
  }
}

// MODE_S1 with the previous predict_correct's update folded in (k_stage LAZY = 1 / 2); never with WENO (its
// advection terms come from separate sweeps over a stored state)
template <int LZ>
static stage_fn pick_stage_lazy_t(int pass, int adv) {
  if (pass == PASS_FAST) return k_stage<PASS_FAST, ADV_CENTER, MODE_S1, LZ, false>;
  if (pass == PASS_ALL)
    return adv == ADV_UPWIND ? k_stage<PASS_ALL, ADV_UPWIND, MODE_S1, LZ, false> : k_stage<PASS_ALL, ADV_CENTER, MODE_S1, LZ, false>;
  return adv == ADV_UPWIND ? k_stage<PASS_SLOW, ADV_UPWIND, MODE_S1, LZ, false> : k_stage<PASS_SLOW, ADV_CENTER, MODE_S1, LZ, false>;
}
static stage_fn pick_stage_lazy(int pass, int adv, int lazy) {
  return lazy == 1 ? pick_stage_lazy_t<1>(pass, adv) : pick_stage_lazy_t<2>(pass, adv);
}

// the fused predict_correct kernel (gmd_pc.cuh; product build, never WENO)
#if !GMD_STRICT
template <int LZ, bool PUSH>
static stage_fn pick_pc_t(int pass, int adv) {
  if (pass == PASS_FAST) return k_pc<PASS_FAST, ADV_CENTER, LZ, PUSH>;
  if (pass == PASS_ALL) return adv == ADV_UPWIND ? k_pc<PASS_ALL, ADV_UPWIND, LZ, PUSH> : k_pc<PASS_ALL, ADV_CENTER, LZ, PUSH>;
  return adv == ADV_UPWIND ? k_pc<PASS_SLOW, ADV_UPWIND, LZ, PUSH> : k_pc<PASS_SLOW, ADV_CENTER, LZ, PUSH>;
}
static stage_fn pick_pc(int pass, int adv, int lazy, bool push) {
  if (adv == ADV_WENO) return nullptr;
  if (push) return lazy == 0 ? pick_pc_t<0, true>(pass, adv) : (lazy == 1 ? pick_pc_t<1, true>(pass, adv) : pick_pc_t<2, true>(pass, adv));
  return lazy == 0 ? pick_pc_t<0, false>(pass, adv) : (lazy == 1 ? pick_pc_t<1, false>(pass, adv) : pick_pc_t<2, false>(pass, adv));
}
static const size_t PC_SMEM_BYTES = PC_SMEM;
#else
static stage_fn pick_pc(int, int, int, bool) { return nullptr; }
static const size_t PC_SMEM_BYTES = 0;
static const int WOUT3 = 54, PC_BX = 96;
#endif

// the fused polar-cap kernel (never WENO: its advection terms come from separate sweeps)
typedef void (*cap_fn)(const StageArgs, const PolarArgs, const CapArgs);
template <int MODE, int LZ>
static cap_fn pick_cap_t(int pass, int adv) {
  if (pass == PASS_FAST) return k_cap<PASS_FAST, ADV_CENTER, MODE, LZ>;
  if (pass == PASS_ALL) return adv == ADV_UPWIND ? k_cap<PASS_ALL, ADV_UPWIND, MODE, LZ> : k_cap<PASS_ALL, ADV_CENTER, MODE, LZ>;
  return adv == ADV_UPWIND ? k_cap<PASS_SLOW, ADV_UPWIND, MODE, LZ> : k_cap<PASS_SLOW, ADV_CENTER, MODE, LZ>;
}
static cap_fn pick_cap(int pass, int adv, int mode, int lazy) {
  if (adv == ADV_WENO && pass != PASS_FAST) return nullptr;
  if (lazy) return lazy == 1 ? pick_cap_t<MODE_S1, 1>(pass, adv) : pick_cap_t<MODE_S1, 2>(pass, adv);
  switch (mode) {
    case MODE_S1: return pick_cap_t<MODE_S1, 0>(pass, adv);
    case MODE_S2: return pick_cap_t<MODE_S2, 0>(pass, adv);
    case MODE_S3A: return pick_cap_t<MODE_S3A, 0>(pass, adv);
    default: return pick_cap_t<MODE_EVAL, 0>(pass, adv);
  }
}
static int launch_cap(gmd_model *m, cap_fn fn, int grid, const StageArgs &b, const PolarArgs &p, const CapArgs &c) {
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid, 1, 1);
  lc.blockDim = dim3(BX, 1, 1);
  lc.dynamicSmemBytes = m->stage_smem_b;
  lc.stream = m->stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributePriority;
  at[1].val.priority = m->prio_hi;
  lc.attrs = at;
  lc.numAttrs = 2;
  if (int r = allow_smem((const void *)fn)) return r;
  CK(cudaLaunchKernelEx(&lc, fn, b, p, c));
  return 0;
}

static int post_launch(gmd_model *m) {
  m->launches++;
  if (m->dry) return 0;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) return fail(GMD_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// halo rows of one field: send my top `ns` rows north / bottom `nn` rows south, receive the matching ghosts
static int exchange_field(gmd_model *m, double *f, int ns, int nn) {
  const int nlon = m->geo.nlon, nr = m->nr;
  const int rank = m->cfg.rank, np = m->cfg.nranks;
  if (rank + 1 < np) {
    if (ns) NK(g_nccl.Send(f + (size_t)(nr - ns) * nlon, (size_t)ns * nlon, NCCL_F64, rank + 1, m->comm, m->stream));
    if (nn) NK(g_nccl.Recv(f + (size_t)nr * nlon, (size_t)nn * nlon, NCCL_F64, rank + 1, m->comm, m->stream));
  }
  if (rank > 0) {
    if (nn) NK(g_nccl.Send(f, (size_t)nn * nlon, NCCL_F64, rank - 1, m->comm, m->stream));
    if (ns) NK(g_nccl.Recv(f - (ptrdiff_t)ns * nlon, (size_t)ns * nlon, NCCL_F64, rank - 1, m->comm, m->stream));
  }
  return 0;
}
// ---- peer-memory path ------------------------------------------------------------------------------------
// the neighbour's copy of my buffer `mine` (same slab slot), shifted so that the element offset of (row j, column
// i) is the one I use: q[(j - r0) nlon + i] is the neighbour's (j, i)
static double *peer_ptr(const gmd_model *m, int side, const double *mine) {
  if (!mine || !m->peer_slab[side]) return nullptr;
  const size_t nlon = (size_t)m->geo.nlon;
  const size_t slot = (size_t)((mine - (size_t)GHOST * nlon) - m->slab) / m->fld_elems;
  double *base = m->peer_slab[side] + slot * m->peer_fld[side] + (size_t)GHOST * nlon;  // the neighbour's row r0'
  return base + (ptrdiff_t)(m->geo.r0 - m->peer_r0[side]) * (ptrdiff_t)nlon;
}
static int halo_sides(const gmd_model *m) {
  return (m->cfg.rank > 0 ? 1 : 0) | (m->cfg.rank + 1 < m->cfg.nranks ? 2 : 0);
}
// store halo rows of up to three fields into the neighbours' ghost rows and release halo epoch ++xk
static int halo_wait(gmd_model *m);
static int halo_push(gmd_model *m, double *const f[3], const int ns[3], const int nn[3]) {
  if (m->need_fence) {
    m->need_fence = false;
    double *const none[3] = {nullptr, nullptr, nullptr};
    const int z[3] = {0, 0, 0};
    int r;
    if ((r = halo_push(m, none, z, z))) return r;   // signal only
    if ((r = halo_wait(m))) return r;
  }
  const int ts = tseq(m, "k_halo_push");
  m->xk++;
  m->launches++;
  if (m->dry) return 0;
  PushArgs a;
  memset(&a, 0, sizeof a);
  a.tseq = ts;
  a.g = m->geo;
  int rows = 0;
  for (int k = 0; k < 3; k++) {
    a.src[k] = f[k];
    a.dstS[k] = peer_ptr(m, 0, f[k]);
    a.dstN[k] = peer_ptr(m, 1, f[k]);
    a.ns[k] = ns[k];
    a.nn[k] = nn[k];
    if (f[k]) rows += (a.dstS[k] ? nn[k] : 0) + (a.dstN[k] ? ns[k] : 0);
  }
  a.page = m->page;
  a.sigS = (m->cfg.rank > 0) ? m->peer_page[m->cfg.rank - 1] + SP_SIG + 1 : nullptr;
  a.sigN = (m->cfg.rank + 1 < m->cfg.nranks) ? m->peer_page[m->cfg.rank + 1] + SP_SIG : nullptr;
  a.k = m->xk;
  const int units = rows * (m->geo.nlon / 2);
  const int nb = std::max(1, std::min(64, (units + 255) / 256));
  k_halo_push<<<nb, 256, 0, m->stream>>>(a);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) return fail(GMD_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return 0;
}
// before a launch on the main stream that reads ghost rows and does not wait by itself
static int halo_wait(gmd_model *m) {
  if (!m->p2p || m->xwaited == m->xk) return 0;
  m->xwaited = m->xk;
  const int ts = tseq(m, "k_halo_wait");
  m->launches++;
  if (m->dry) return 0;
  k_halo_wait<<<1, 32, 0, m->stream>>>(m->page, m->xk, halo_sides(m), ts);
  return 0;
}
// close a unit of work (k_unit_end): all incoming halo rows have landed, epoch bases advance
static int unit_end(gmd_model *m) {
  if (!m->p2p || (m->xk == 0 && m->rk == 0)) return 0;
  const int ts = tseq(m, "k_unit_end");
  if (!m->dry) k_unit_end<<<1, 32, 0, m->stream>>>(m->page, m->xk, m->rk, halo_sides(m), ts);
  m->launches++;
  m->xk = m->rk = m->xwaited = 0;
  return 0;
}
static RedArgs red_args(gmd_model *m) {
  RedArgs r;
  memset(&r, 0, sizeof r);
  r.nranks = 1;
  if (m->p2p) {
    r.page = m->page;
    r.rank = m->cfg.rank;
    r.nranks = m->cfg.nranks;
    r.k = ++m->rk;
  }
  return r;
}

// ghost rows of a state: a sweep needs U, V, gd row r0-1 from the south and U, V row r1, gd rows r1, r1+1 from the
// north; a wide-halo predict_correct starts from HALO_S / HALO_N rows.  Every state exchange moves the
// wide set (my top HALO_S rows go north, my bottom HALO_N rows go south).
static int exchange_state(gmd_model *m, const State &s, bool with_gd) {
  if (m->cfg.nranks == 1) return 0;
  if (m->p2p) {
    double *const f[3] = {s.U, s.V, with_gd ? s.gd : nullptr};
    const int ns[3] = {HALO_S, HALO_S, HALO_S}, nn[3] = {HALO_N, HALO_N, HALO_N};
    return halo_push(m, f, ns, nn);
  }
  if (m->dry) return 0;
  if (!m->comm) return fail(GMD_ERR_COMM, "neither gmd_peer_connect nor gmd_comm_init has been called on this rank");
  NK(g_nccl.GroupStart());
  int r;
  if ((r = exchange_field(m, s.U, HALO_S, HALO_N))) return r;
  if ((r = exchange_field(m, s.V, HALO_S, HALO_N))) return r;
  if (with_gd && (r = exchange_field(m, s.gd, HALO_S, HALO_N))) return r;
  NK(g_nccl.GroupEnd());
  return 0;
}
static int exchange_tend3(gmd_model *m, double *a, double *b, double *c, int ns, int nn) {
  if (m->cfg.nranks == 1) return 0;
  if (m->p2p) {
    double *const f[3] = {a, b, c};
    const int nsv[3] = {ns, ns, ns}, nnv[3] = {nn, nn, nn};
    return halo_push(m, f, nsv, nnv);
  }
  if (m->dry) return 0;
  if (!m->comm) return fail(GMD_ERR_COMM, "neither gmd_peer_connect nor gmd_comm_init has been called on this rank");
  NK(g_nccl.GroupStart());
  int r;
  if (a && (r = exchange_field(m, a, ns, nn))) return r;
  if (b && (r = exchange_field(m, b, ns, nn))) return r;
  if (c && (r = exchange_field(m, c, ns, nn))) return r;
  NK(g_nccl.GroupEnd());
  return 0;
}
static int allreduce2(gmd_model *m, double *d) {
  if (m->cfg.nranks == 1 || m->dry || m->p2p) return 0;  // peer path: done inside k_reduce_pairs
  if (!m->comm) return fail(GMD_ERR_COMM, "neither gmd_peer_connect nor gmd_comm_init has been called on this rank");
  NK(g_nccl.AllReduce(d, d, 2, NCCL_F64, NCCL_SUM, m->comm, m->stream));
  return 0;
}

// ---- two-stream boundary / interior split ----------------------------------------------------------------
static cudaEvent_t next_event(gmd_model *m) { return m->evpool[m->evnext++ % m->evpool.size()]; }
// main stream waits for the last interior launch (needed before anything that reads / overwrites interior rows)
static int join(gmd_model *m) {
  if (m->dry || !m->last_eI) return 0;
  CK(cudaStreamWaitEvent(m->stream, m->last_eI, 0));
  m->last_eI = nullptr;
  return 0;
}
// stream2 sees everything queued on the main stream so far (previous boundary launch, polar rows, halo exchange,
// reductions); the main stream sees the previous interior launch
// `light`: stream2 only needs the polar-side launch of the previous sweep (recorded in ev_polar_side), valid when
// nothing but that launch and its polar rows has been queued on the main stream since
static int split_begin(gmd_model *m, bool light = false) {
  if (m->dry) return 0;
  if (light && m->ev_polar_side) {
    CK(cudaStreamWaitEvent(m->stream2, m->ev_polar_side, 0));
  } else {
    cudaEvent_t e = next_event(m);
    CK(cudaEventRecord(e, m->stream));
    CK(cudaStreamWaitEvent(m->stream2, e, 0));
  }
  m->ev_polar_side = nullptr;
  return join(m);
}
static int split_end(gmd_model *m) {
  if (m->dry) return 0;
  cudaEvent_t e = next_event(m);
  CK(cudaEventRecord(e, m->stream2));
  m->last_eI = e;
  return 0;
}
static bool use_split(const gmd_model *m) { return m->split && (m->bs + m->bn > 0) && (m->geo.r0 + m->bs < m->geo.r1 - m->bn); }

static int ensure_weno(gmd_model *m) {
  if (m->w_fpu) return 0;
  int r;
  double **us[] = {&m->w_fpu, &m->w_fnu, &m->w_fu, &m->w_alon_u, &m->w_alat_u};
  double **vs[] = {&m->w_fpv, &m->w_fnv, &m->w_fv, &m->w_alon_v, &m->w_alat_v};
  for (auto p : us)
    if ((r = acquire(m, KIND_U, p))) return r;
  for (auto p : vs)
    if ((r = acquire(m, KIND_V, p))) return r;
  return 0;
}
static int ensure_uv(gmd_model *m) {
  if (m->w_u) return 0;
  int r;
  if ((r = acquire(m, KIND_U, &m->w_u))) return r;
  if ((r = acquire(m, KIND_V, &m->w_v))) return r;
  return 0;
}

// derived u, v of state s on rows [max(r0-1,0), min(r1+1,nlat))
static int derive_uv(gmd_model *m, const State &s) {
  int r;
  if ((r = ensure_uv(m))) return r;
  if ((r = join(m))) return r;
  if ((r = halo_wait(m))) return r;
  const int ja = std::max(m->geo.r0 - 1, 0), jb = std::min(m->geo.r1 + 1, m->geo.nlat);
  if (!m->dry) {
    if (m->ew_rows) k_derive2<<<dim3(m->ew2_bx, (unsigned)(jb - ja)), EW2, 0, m->stream>>>(m->geo, ja, s.U, s.V, s.gd, m->w_u, m->w_v, nullptr);
    else k_derive<<<m->ew_blocks, 256, 0, m->stream>>>(m->geo, ja, jb, s.U, s.V, s.gd, m->w_u, m->w_v, nullptr);
  }
  return post_launch(m);
}

// WENO advection terms of state E into w_alon_u .. (src/weno_mod.F90:69-233)
static int weno_terms(gmd_model *m, const State &E) {
  int r;
  if ((r = ensure_weno(m))) return r;
  if ((r = derive_uv(m, E))) return r;
  WenoArgs a;
  a.g = m->geo;
  a.t = m->tab;
  a.u = m->w_u; a.v = m->w_v; a.U = E.U; a.V = E.V;
  a.fpu = m->w_fpu; a.fnu = m->w_fnu; a.fpv = m->w_fpv; a.fnv = m->w_fnv;
  a.fu = m->w_fu; a.fv = m->w_fv;
  a.alon_u = m->w_alon_u; a.alat_u = m->w_alat_u; a.alon_v = m->w_alon_v; a.alat_v = m->w_alat_v;
  for (int dir = 0; dir < 2; dir++) {
    a.dir = dir;
    if (!m->dry) k_weno_split<<<m->ew_blocks, 256, 0, m->stream>>>(a);
    if ((r = post_launch(m))) return r;
    if (dir == 1) {
      // the meridional reconstruction reads the split fluxes on rows j-1 .. j+2 (src/weno_mod.F90:186-217): one row
      // from the south, two from the north
      if ((r = exchange_tend3(m, a.fpu, a.fnu, a.fpv, 1, 2))) return r;
      if ((r = exchange_tend3(m, a.fnv, nullptr, nullptr, 1, 2))) return r;
      if ((r = halo_wait(m))) return r;
    }
    if (!m->dry) k_weno_flux<<<m->ew_blocks, 256, 0, m->stream>>>(a);
    if ((r = post_launch(m))) return r;
    if (dir == 1) {
      // the flux difference reads the reconstructed flux of row j-1 (:219-233)
      if ((r = exchange_tend3(m, a.fu, a.fv, nullptr, 1, 0))) return r;
      if ((r = halo_wait(m))) return r;
    }
    if (!m->dry) k_weno_adv<<<m->ew_blocks, 256, 0, m->stream>>>(a);
    if ((r = post_launch(m))) return r;
  }
  return 0;
}

// polar-row kernel launch: items in the parameters, rows (x, w and, if it fits, the prefetched base row) in smem
static size_t polar_smem(const gmd_model *m, int *use_q) {
  const size_t row = (size_t)m->geo.nlon * sizeof(double);
  *use_q = (3 * row <= 200 * 1024) ? 1 : 0;
  return (*use_q ? 3 : 2) * row;
}
static void fill_items(PolarArgs &p, const std::vector<unsigned> &v) {
  for (size_t k = 0; k < v.size() && k < (size_t)MAX_ITEMS; k++) p.items[k] = v[k];
}
// launch with the programmatic-stream-serialization attribute (the kernel must call pdl_wait())
template <typename A>
static void launch_pdl(void (*fn)(const A), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const A &args) {
  cudaLaunchConfig_t lc = {};
  lc.gridDim = grid;
  lc.blockDim = block;
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  cudaLaunchKernelEx(&lc, fn, args);
}
// element groups per thread (1 or 2) when every item of the launch is one k_polar_lean handles, else 0
static int polar_lean_groups(const gmd_model *m, const PolarArgs &p, int nitems) {
  const char *ev = getenv("GMD_POLAR_LEAN");   // (read per launch: the tests switch it between two models of one process)
  const bool off = ev && atoi(ev) == 0;
  const int n = m->geo.nlon;
  if (off || GMD_STRICT || (n & 3)) return 0;
  const int G = (n / 4 + PT - 1) / PT;
  if (G > 2) return 0;
  for (int q = 0; q < nitems; q++) {
    const unsigned pk = p.items[q];
    const int K = (int)((pk >> 16) & 0x1ffu), kind = (int)((pk >> 28) & 7u);   // K = cutoff + 1
    if (kind == IT_POLE_S || kind == IT_POLE_N) continue;
    if ((pk & ITEM_REDUCE) || K < 1 || K > KF || 2 * K >= n) return 0;
  }
  return G;
}
template <int G>
static void (*pick_polar_lean(int mode))(const PolarArgs) {
  return (mode == MODE_S1) ? k_polar_lean<MODE_S1, G> : (mode == MODE_S2) ? k_polar_lean<MODE_S2, G>
       : (mode == MODE_S3A) ? k_polar_lean<MODE_S3A, G> : k_polar_lean<MODE_EVAL, G>;
}
static void launch_polar(gmd_model *m, int mode, int nitems, PolarArgs &p, cudaStream_t st, bool pdl = false) {
  p.basis = m->d_basis;
  p.rot = m->d_rot;
  p.pdl = pdl ? 1 : 0;
  size_t sh = polar_smem(m, &p.use_q);
  void (*fn)(const PolarArgs) = (mode == MODE_S1) ? k_polar<MODE_S1> : (mode == MODE_S2) ? k_polar<MODE_S2>
                              : (mode == MODE_S3A) ? k_polar<MODE_S3A> : k_polar<MODE_EVAL>;
  if (const int G = polar_lean_groups(m, p, nitems)) {   // rows in registers: no dynamic shared memory
    fn = (G == 1) ? pick_polar_lean<1>(mode) : pick_polar_lean<2>(mode);
    sh = 0;
  }
  if (pdl) launch_pdl<PolarArgs>(fn, dim3((unsigned)nitems), dim3(PT), sh, st, p);
  else fn<<<nitems, PT, sh, st>>>(p);
}

// arguments of the polar rows that belong to the sweep described by `a`
static void polar_args(gmd_model *m, const StageArgs &a, bool lazy, int li, int nst, double dt, PolarArgs *pp) {
  PolarArgs &p = *pp;
  memset(&p, 0, sizeof p);
  p.g = m->geo;
  p.t = m->tab;
  fill_items(p, m->items[li]);
  p.EU = lazy ? a.MU : a.EU; p.EV = lazy ? a.MV : a.EV; p.Egd = lazy ? a.Mgd : a.Egd; p.ghs = a.ghs;
  p.OU = a.OU; p.OV = a.OV; p.Ogd = a.Ogd;
  p.NU = a.NU; p.NV = a.NV; p.Ngd = a.Ngd;
  p.TU = a.TU; p.TV = a.TV; p.Tgd = a.Tgd;
  p.PU = a.PU; p.PV = a.PV; p.Pgd = a.Pgd;
  p.dt = dt;
  p.partials = m->d_partials + 2 * (size_t)nst;
  p.rescale = 1;
  p.radius = m->mesh.radius;
  p.dlat = m->mesh.dlat;
  p.fold = a.fold;
  p.fold_partials = m->d_partials;
  p.basis = m->d_basis;
  p.rot = m->d_rot;
  p.reduce_smooth = m->cfg.use_reduce_tend_smooth;
}

// one fused operator evaluation (+ update / store / dots) of state E
// A deferred update handed to the next MODE_S1 launch: E = base + beta ldt L, written out to M
struct LazyIn {
  const Tend *L;
  double ldt;
  State *M;
  int kind;   // 1: U, V and gd carry a tendency; 2: U, V only (previous pass was slow)
};

// es / en: the sweep also covers `es` rows south and `en` rows north of the band (wide-halo predict_correct; 0 on a
// side without a neighbour).  push: MODE_S3A of a wide-halo predict_correct on the peer path -- the band-edge rows
// of the tendency T are stored into the neighbours' ghost rows by the launch itself.
static int stage(gmd_model *m, int pass, int mode, const State &E, const State *O, double dt, State *N, Tend *T,
                 const Tend *P, const LazyIn *lz = nullptr, int es = 0, int en = 0, bool push = false) {
  int r;
  const int adv = m->cfg.uv_adv_scheme;
  if (pass != PASS_FAST && adv == ADV_WENO && (r = weno_terms(m, E))) return r;
  if ((r = halo_wait(m))) return r;   // rows pushed by a generic exchange (isp, diffusion, WENO, direct API calls)
  StageArgs a;
  memset(&a, 0, sizeof a);
  a.g = m->geo;
  a.t = m->tab;
  a.EU = E.U; a.EV = E.V; a.Egd = E.gd; a.ghs = m->ghs;
  if (O) { a.OU = O->U; a.OV = O->V; a.Ogd = O->gd; }
  if (N) { a.NU = N->U; a.NV = N->V; a.Ngd = N->gd; }
  if (T) { a.TU = T->U; a.TV = T->V; a.Tgd = T->gd; }
  if (P) { a.PU = P->U; a.PV = P->V; a.Pgd = P->gd; }
  a.AUlon = m->w_alon_u; a.AUlat = m->w_alat_u; a.AVlon = m->w_alon_v; a.AVlat = m->w_alat_v;
  a.dt = dt;
  a.beta_lon = m->cfg.uv_adv_upwind_lon_beta;
  a.beta_lat = m->cfg.uv_adv_upwind_lat_beta;
  a.partials = m->d_partials;
  a.flmask = 0x0fu | ((pass != PASS_SLOW || m->cfg.reduce_adv_lon) ? 0xf0u : 0u);
  // the inner products are finalised (and all-reduced over peer memory) by the last CTA of the S3a launches; over
  // NCCL the all-reduce is a library call, so the partials are reduced by a launch of their own
  // (worth it once a band is short -- measured: 3.6 % at 225 rows per rank, nothing at 900 and above, where the
  // extra ticket per CTA costs as much as the launch it saves)
  static const int fold_env = getenv("GMD_FOLD") ? atoi(getenv("GMD_FOLD")) : -1;
  const bool fold = (mode == MODE_S3A) && (m->cfg.nranks == 1 || m->p2p) && (fold_env >= 0 ? fold_env != 0 : m->nr < 600);
  stage_fn fn = pick_stage(pass, adv, mode);
  if (lz) {
    a.LU = lz->L->U; a.LV = lz->L->V; a.Lgd = lz->L->gd;
    a.MU = lz->M->U; a.MV = lz->M->V; a.Mgd = lz->M->gd;
    a.OU = lz->M->U; a.OV = lz->M->V; a.Ogd = lz->M->gd;   // for the polar-row kernel: old state == evaluated state
    a.lip = m->d_ip;
    a.ldt = lz->ldt;
    a.lqcon = m->cfg.qcon_modified;
    fn = pick_stage_lazy(pass, adv, lz->kind);
  }
  const int r0 = m->geo.r0, r1 = m->geo.r1;
  if (push && mode == MODE_S3A && m->p2p) {
    stage_fn pf = pick_stage(pass, adv, mode, true);
    if (pf) {
      double *fg = (pass == PASS_SLOW) ? nullptr : T->gd;
      fn = pf;
      a.hpS_U = peer_ptr(m, 0, T->U); a.hpS_V = peer_ptr(m, 0, T->V); a.hpS_G = peer_ptr(m, 0, fg);
      a.hpN_U = peer_ptr(m, 1, T->U); a.hpN_V = peer_ptr(m, 1, T->V); a.hpN_G = peer_ptr(m, 1, fg);
      a.push_s_end = r0 + HALO_N;
      a.push_n_begin = r1 - HALO_S;
    }
  }
  if ((r = allow_smem((const void *)fn))) return r;
  const int li = (pass == PASS_SLOW) ? 1 : 0;
  if (m->fz_active) {
    // fused predict_correct: k_pc (already launched, stream2) covers rows [fz_I0, fz_I1); this call runs the rows next to
    // the poles, widened towards the fused rows by what the later sweeps of this chain read: S1 (4 north of the southern
    // range, 2 south of the northern one), S2 (2, 1), S3a none
    const int xs = (mode == MODE_S1) ? 4 : (mode == MODE_S2 ? 2 : 0), xn = (mode == MODE_S1) ? 2 : (mode == MODE_S2 ? 1 : 0);
    const bool haveS = m->fz_I0 > r0, haveN = m->fz_I1 < r1;
    const int ncb = (haveS || haveN) ? m->fz_ncb : 0;
    const int nst = 2 * m->nbx * ncb + m->fz_nstrips * m->fz_nchunks;
    if (mode == MODE_S3A) a.fold = m->fz_fold;
    if (ncb) {
      StageArgs b = a;
      b.hpS_U = b.hpS_V = b.hpS_G = b.hpN_U = b.hpN_V = b.hpN_G = nullptr;
      b.rows_per_cta = m->rows_per_cta_b;
      b.rb[0] = r0; b.re[0] = haveS ? m->fz_I0 + xs : r0; b.pofs[0] = 0;
      b.rb[1] = haveN ? m->fz_I1 - xn : r1; b.re[1] = r1; b.pofs[1] = m->nbx * ncb;
      dim3 gb((unsigned)m->nbx, (unsigned)ncb, 2);
      static const char *const bnames[4] = {"k_stage.S1.polar_side", "k_stage.S2.polar_side", "k_stage.S3a.polar_side", "k_stage.eval.polar_side"};
      b.tseq = tseq(m, bnames[mode]);
      b.pdl = (m->pdl & 2) && (mode == MODE_S2 || mode == MODE_S3A) && m->n_items[li] ? 1 : 0;
      if (!m->dry) {
        if (b.pdl) launch_pdl<StageArgs>(fn, gb, dim3(BX), m->stage_smem_b, m->stream, b);
        else fn<<<gb, BX, m->stage_smem_b, m->stream>>>(b);
      }
      if ((r = post_launch(m))) return r;
    }
    if (m->n_items[li]) {
      PolarArgs p;
      polar_args(m, a, lz != nullptr, li, nst, dt, &p);
      p.tseq = tseq(m, "k_polar");
      if (!m->dry) launch_polar(m, mode, m->n_items[li], p, m->stream, ncb && (m->pdl & 1));
      if ((r = post_launch(m))) return r;
    }
    if (mode == MODE_S3A) {
      if (m->fz_ev) m->last_eI = m->fz_ev;   // whoever reads the new tendency next waits for k_pc
      m->fz_ev = nullptr;
      if (!a.fold.ticket) {
        if ((r = join(m))) return r;
        RedArgs ra = red_args(m);
        ra.tseq = tseq(m, "k_reduce_pairs.ip");
        if (!m->dry) k_reduce_pairs<<<1, 256, 0, m->stream>>>(m->d_partials, nst + m->n_items[li], m->d_ip, ra);
        if ((r = post_launch(m))) return r;
        if ((r = allreduce2(m, m->d_ip))) return r;
      }
    }
    return 0;
  }
  const bool poleS = (r0 == 0), poleN = (r1 == m->geo.nlat);
  const int R0 = r0 - es, R1 = r1 + en;   // rows of this sweep
  const bool split = use_split(m);
  // launch geometry
  const int I0 = split ? (poleS ? r0 + m->bs : R0) : R0, I1 = split ? (poleN ? r1 - m->bn : R1) : R1;
  const int rpc = (mode == MODE_S3A) ? m->rows_per_cta_s3a : m->rows_per_cta;
  const size_t smem_i = stage_smem_bytes(rpc);
  const int nci = (I1 - I0 + rpc - 1) / rpc;
  const int ncb = split ? m->nchunks_b : 0;
  const cap_fn cfn = (split && m->cap && m->n_items[li]) ? pick_cap(pass, adv, mode, lz ? lz->kind : 0) : nullptr;
  // row ranges next to a pole: the fused cap launch covers only those that exist on this band
  const int nz = cfn ? (poleS ? 1 : 0) + (poleN ? 1 : 0) : 2;
  const int nst = nz * m->nbx * ncb + m->nbx * nci;
  if (fold) {  // one partial pair and one ticket per CTA of the stage launch(es) and of the polar-row launch
    a.fold.ticket = reinterpret_cast<unsigned *>(m->d_ip + 7);
    a.fold.total = (unsigned)(nst + m->n_items[li]);
    a.fold.n = nst + m->n_items[li];
    a.fold.out = m->d_ip;
    a.fold.r = red_args(m);
  }
  const int edges = (es ? 1 : 0) | (en ? 2 : 0);   // sides on which the deferred update also writes the rows beyond the sweep
  if (split) {
    // rows next to a pole first on the main stream (then the polar rows), the other rows on stream2
    // S2 / S3a follow S1 / S2 of the same predict_correct directly: their interior launch reads nothing the polar rows
    // of the previous sweep produce (S1 of a deferred update reads the inner products: full dependency)
    if ((r = split_begin(m, mode == MODE_S2 || mode == MODE_S3A))) return r;
    StageArgs b = a;
    b.hpS_U = b.hpS_V = b.hpS_G = b.hpN_U = b.hpN_V = b.hpN_G = nullptr;   // band-edge rows are interior rows
    b.rows_per_cta = m->rows_per_cta_b;
    b.rb[0] = r0; b.re[0] = poleS ? r0 + m->bs : r0; b.pofs[0] = 0;
    b.rb[1] = poleN ? r1 - m->bn : r1; b.re[1] = r1; b.pofs[1] = m->nbx * ncb;
    dim3 gb((unsigned)m->nbx, (unsigned)ncb, 2);
    static const char *const bnames[4] = {"k_stage.S1.polar_side", "k_stage.S2.polar_side", "k_stage.S3a.polar_side", "k_stage.eval.polar_side"};
    static const char *const cnames[4] = {"k_cap.S1", "k_cap.S2", "k_cap.S3a", "k_cap.eval"};
    if (cfn) {   // the sweep over these rows and their polar rows in one launch
      b.tseq = tseq(m, cnames[mode]);
      PolarArgs p;
      polar_args(m, a, lz != nullptr, li, nst, dt, &p);
      p.tseq = b.tseq;
      CapArgs c;
      c.bar = m->d_bar;
      c.n_march = nz * m->nbx * ncb;
      c.gx = m->nbx;
      c.gy = ncb;
      c.z0 = poleS ? 0 : 1;
      c.nitems = m->n_items[li];
      b.pofs[c.z0] = 0;                 // partial slots of the ranges the launch covers, compact
      b.pofs[1 - c.z0] = (nz == 2) ? m->nbx * ncb : 0;
      const int want_ctas = std::min(m->cap_ctas, std::max((c.n_march + CL - 1) / CL, c.nitems) * CL);
      if (!m->dry && (r = launch_cap(m, cfn, want_ctas, b, p, c))) return r;
    } else {
      b.tseq = tseq(m, bnames[mode]);
      // the polar chain (this launch -> polar rows -> the next sweep's launch) with programmatic dependent launches:
      // S2 / S3a follow the polar rows of the previous sweep with nothing but stream waits in between
      b.pdl = (m->pdl & 2) && (mode == MODE_S2 || mode == MODE_S3A) && m->n_items[li] ? 1 : 0;
      if (!m->dry) {
        if (b.pdl) launch_pdl<StageArgs>(fn, gb, dim3(BX), m->stage_smem_b, m->stream, b);
        else fn<<<gb, BX, m->stage_smem_b, m->stream>>>(b);
      }
    }
    if ((r = post_launch(m))) return r;
    if (!m->dry) {   // what the NEXT sweep's interior launch has to wait for on the main stream: this launch, not the
      cudaEvent_t e = next_event(m);   // polar rows that follow it (its rows are at least two rows away from them)
      CK(cudaEventRecord(e, m->stream));
      m->ev_polar_side = e;
    }
    a.rows_per_cta = rpc;
    a.rb[0] = I0; a.re[0] = I1; a.pofs[0] = nz * m->nbx * ncb;
    a.medge[0] = edges & ((poleS ? 0 : 1) | (poleN ? 0 : 2));
    dim3 gi((unsigned)m->nbx, (unsigned)nci, 1);
    static const char *const inames[4] = {"k_stage.S1.interior", "k_stage.S2.interior", "k_stage.S3a.interior", "k_stage.eval.interior"};
    a.tseq = tseq(m, inames[mode]);
    if (!m->dry) fn<<<gi, BX, smem_i, m->stream2>>>(a);
    if ((r = post_launch(m))) return r;
    if ((r = split_end(m))) return r;
  } else {
    if ((r = join(m))) return r;
    a.rows_per_cta = rpc;
    a.rb[0] = R0; a.re[0] = R1; a.pofs[0] = 0;
    a.medge[0] = edges;
    dim3 grid((unsigned)m->nbx, (unsigned)nci, 1);
    static const char *const wnames[4] = {"k_stage.S1", "k_stage.S2", "k_stage.S3a", "k_stage.eval"};
    a.tseq = tseq(m, wnames[mode]);
    // a band without polar rows runs a plain chain of stage launches: programmatic dependent launch hides the launch
    // latency between them (GMD_PDL bit 2)
    a.pdl = ((m->pdl & 4) && m->cfg.nranks > 1 && !m->n_items[li]) ? 1 : 0;
    if (!m->dry) {
      if (a.pdl) launch_pdl<StageArgs>(fn, grid, dim3(BX), smem_i, m->stream, a);
      else fn<<<grid, BX, smem_i, m->stream>>>(a);
    }
    if ((r = post_launch(m))) return r;
  }

  if (m->n_items[li] && !cfn) {
    PolarArgs p;
    polar_args(m, a, lz != nullptr, li, nst, dt, &p);
    p.tseq = tseq(m, "k_polar");
    if (!m->dry) launch_polar(m, mode, m->n_items[li], p, m->stream, split && (m->pdl & 1));
    if ((r = post_launch(m))) return r;
  }
  if (mode == MODE_S3A && !fold) {
    if ((r = join(m))) return r;
    {
      RedArgs ra = red_args(m);
      ra.tseq = tseq(m, "k_reduce_pairs.ip");
      if (!m->dry) k_reduce_pairs<<<1, 256, 0, m->stream>>>(m->d_partials, nst + m->n_items[li], m->d_ip, ra);
    }
    if ((r = post_launch(m))) return r;
    if ((r = allreduce2(m, m->d_ip))) return r;
  }
  return 0;
}

// es / en: also on `es` ghost rows south and `en` ghost rows north of the band (wide-halo predict_correct: the
// tendency's ghost rows have arrived with the inner-product all-reduce)
static int update(gmd_model *m, const State &O, const Tend &T, double dt, int beta_mode, double dt0, bool with_gd,
                  State *N, int es = 0, int en = 0) {
  UpdateArgs a;
  memset(&a, 0, sizeof a);
  a.g = m->geo;
  a.OU = O.U; a.OV = O.V; a.Ogd = O.gd;
  a.TU = T.U; a.TV = T.V; a.Tgd = T.gd;
  a.NU = N->U; a.NV = N->V; a.Ngd = N->gd;
  a.dt = dt;
  a.ip = m->d_ip;
  a.qcon = m->cfg.qcon_modified;
  a.beta_mode = beta_mode;
  a.dt0 = dt0;
  a.beta_out = m->d_beta;
  a.with_gd = with_gd ? 1 : 0;
  int r;
  const int r0 = m->geo.r0 - es, r1 = m->geo.r1 + en;
  if ((r = halo_wait(m))) return r;
  if ((r = join(m))) return r;
  a.rb[0] = r0; a.re[0] = r1; a.rb[1] = a.re[1] = 0;
  a.tseq = tseq(m, "k_update");
  if (!m->dry) k_update<<<m->ew_blocks, 256, 0, m->stream>>>(a);
  return post_launch(m);
}

// The fused sweep of a predict_correct over rows [fz_I0, fz_I1) (k_pc, gmd_pc.cuh): tend(new) of L(L-updated states)
// into Tn, the inner-product partials, the deferred update materialised into lz->M.  With rows next to a pole on this
// band the launch goes to stream2 and the polar-side chain (the three stage() calls that follow, m->fz_active) runs
// beside it on the main stream; otherwise it is the only launch of the predict_correct and stays on the main stream.
static int launch_pc(gmd_model *m, int pass, const State &E, const LazyIn *lz, double dt, Tend *Tn, bool push) {
  int r;
  const int adv = m->cfg.uv_adv_scheme;
  const int r0 = m->geo.r0, r1 = m->geo.r1;
  const bool push_ok = push && m->p2p;
  stage_fn fn = pick_pc(pass, adv, lz ? lz->kind : 0, push_ok);
  if (!fn) return fail(GMD_ERR_STATE, "internal: no fused predict_correct kernel for this configuration");
  if ((r = halo_wait(m))) return r;
  StageArgs a;
  memset(&a, 0, sizeof a);
  a.g = m->geo;
  a.t = m->tab;
  a.EU = E.U; a.EV = E.V; a.Egd = E.gd; a.ghs = m->ghs;
  a.TU = Tn->U; a.TV = Tn->V; a.Tgd = Tn->gd;
  a.dt = dt;
  a.beta_lon = m->cfg.uv_adv_upwind_lon_beta;
  a.beta_lat = m->cfg.uv_adv_upwind_lat_beta;
  a.partials = m->d_partials;
  if (lz) {
    a.LU = lz->L->U; a.LV = lz->L->V; a.Lgd = lz->L->gd;
    a.MU = lz->M->U; a.MV = lz->M->V; a.Mgd = lz->M->gd;
    a.lip = m->d_ip;
    a.ldt = lz->ldt;
    a.lqcon = m->cfg.qcon_modified;
  }
  if (push_ok) {
    double *fg = (pass == PASS_SLOW) ? nullptr : Tn->gd;
    a.hpS_U = peer_ptr(m, 0, Tn->U); a.hpS_V = peer_ptr(m, 0, Tn->V); a.hpS_G = peer_ptr(m, 0, fg);
    a.hpN_U = peer_ptr(m, 1, Tn->U); a.hpN_V = peer_ptr(m, 1, Tn->V); a.hpN_G = peer_ptr(m, 1, fg);
    a.push_s_end = r0 + HALO_N;
    a.push_n_begin = r1 - HALO_S;
  }
  const bool polar_side = (m->fz_I0 > r0) || (m->fz_I1 < r1);
  a.rows_per_cta = m->fz_rpc;
  a.rb[0] = m->fz_I0; a.re[0] = m->fz_I1;
  a.pofs[0] = polar_side ? 2 * m->nbx * m->fz_ncb : 0;
  // band edges with a neighbour: the deferred update is also materialised on the ghost rows the sweep reads
  a.medge[0] = ((m->wide && m->cfg.rank > 0 && m->fz_I0 == r0) ? 1 : 0) | ((m->wide && m->cfg.rank + 1 < m->cfg.nranks && m->fz_I1 == r1) ? 2 : 0);
  a.fold = m->fz_fold;
  dim3 grid((unsigned)m->fz_nstrips, (unsigned)m->fz_nchunks, 1);
  a.tseq = tseq(m, "k_pc");
  if ((r = join(m))) return r;
  m->fz_ev = nullptr;
  if (polar_side) {
    if (!m->dry) {
      cudaEvent_t e = next_event(m);
      CK(cudaEventRecord(e, m->stream));
      CK(cudaStreamWaitEvent(m->stream2, e, 0));
      fn<<<grid, PC_BX, PC_SMEM_BYTES, m->stream2>>>(a);
      cudaEvent_t f = next_event(m);
      CK(cudaEventRecord(f, m->stream2));
      m->fz_ev = f;
    }
    m->ev_polar_side = nullptr;
  } else if (!m->dry) {
    fn<<<grid, PC_BX, PC_SMEM_BYTES, m->stream>>>(a);
  }
  return post_launch(m);
}

// A state handed from one predict_correct to the next.  `deferred`: the value is base + beta dts tendNew with beta
// from the inner products still on the device -- the last update_state of predict_correct (src/dycore_mod.F90:
// 786-790) has not been run; the next predict_correct folds it into its first operator sweep (k_stage LAZY), which
// saves one full sweep over the state (9 words per column) and one launch + halo exchange per predict_correct.
struct Carry {
  State base;
  bool deferred = false;
  double dts = 0.0;
  bool with_gd = false;   // the tendency's gd takes part (the pass that produced it was not slow)
  const Tend *L = nullptr;  // deferred: the new tendency of the predict_correct that produced this carry
};

// predict_correct(dt, in -> *out, pass), src/dycore_mod.F90:754-792.  `in.base` is kept (the caller releases it);
// with `defer_out` the result is returned deferred (out->base = the old state of THIS call, retained once more).
//
// Latitude bands (m->wide): the call starts from a state that is valid on HALO_S ghost rows south and HALO_N north of
// the band; sweep 1 covers (2, 4) rows beyond the band, sweep 2 (1, 2), sweep 3 the band itself -- each sweep's
// stencil reaches 1 row south and 2 rows north -- and the only rows exchanged are those of the new tendency, whose
// arrival is signalled by the inner-product all-reduce every rank waits for anyway.  The last update_state (or the
// next call's deferred one) then runs on the ghost rows as well, which restores the invariant.
static int predict_correct(gmd_model *m, double dts, const Carry &in, int pass, bool defer_out, Carry *out) {
  int r;
  const bool slow = (pass == PASS_SLOW);
  const double dt = dts * 0.5;
  const bool wide = m->wide;
  const int hs = (wide && m->cfg.rank > 0) ? 1 : 0, hn = (wide && m->cfg.rank + 1 < m->cfg.nranks) ? 1 : 0;
  Tend *Tn = &m->tendNew;
  if (m->cfg.nranks > 1 || m->fused) {   // see gmd_model::tn_idx
    if (m->tn_idx) Tn = &m->tendNew2;
    m->tn_idx ^= 1;
  }
  State O = in.base, M, A, B;
  bool ownO = false;
  if (in.deferred) {
    // old state of this call = in.base + beta in.dts tend(new) of the previous call, materialised by the first sweep
    if ((r = new_state(m, &M, in.with_gd ? nullptr : in.base.gd))) return r;
    O = M;
    ownO = true;
  }
  if ((r = new_state(m, &A, slow ? O.gd : nullptr))) return r;
  if ((r = new_state(m, &B, slow ? O.gd : nullptr))) return r;
  const bool fz = m->fused;
  if (fz) {
    // the three sweeps of rows [fz_I0, fz_I1) as one wavefront kernel; the stage() calls below then only run the rows
    // next to the poles (m->fz_active)
    const int li = slow ? 1 : 0;
    const bool polar_side = (m->fz_I0 > m->geo.r0) || (m->fz_I1 < m->geo.r1);
    const int nst = (polar_side ? 2 * m->nbx * m->fz_ncb : 0) + m->fz_nstrips * m->fz_nchunks;
    static const int fold_env = getenv("GMD_FOLD") ? atoi(getenv("GMD_FOLD")) : -1;
    memset(&m->fz_fold, 0, sizeof m->fz_fold);
    if ((m->cfg.nranks == 1 || m->p2p) && fold_env != 0) {
      m->fz_fold.ticket = reinterpret_cast<unsigned *>(m->d_ip + 7);
      m->fz_fold.total = (unsigned)(nst + m->n_items[li]);
      m->fz_fold.n = nst + m->n_items[li];
      m->fz_fold.out = m->d_ip;
      m->fz_fold.r = red_args(m);
    }
    LazyIn lz = {in.L, in.dts, &M, in.with_gd ? 1 : 2};
    if ((r = launch_pc(m, pass, in.deferred ? in.base : O, in.deferred ? &lz : nullptr, dt, Tn, wide && m->fuse_push))) return r;
    m->fz_active = true;
  }
  struct FzGuard {   // an error return below must not leave the model in polar-side-only mode
    gmd_model *m;
    ~FzGuard() { m->fz_active = false; }
  } fz_guard = {m};
  const int ws = fz ? 0 : 1;   // the fused kernel widens its own sweeps at the band edges
  // tend(old) = L(old); new = old + dt/2 tend(old)
  if (in.deferred) {
    LazyIn lz = {in.L, in.dts, &M, in.with_gd ? 1 : 2};
    if ((r = stage(m, pass, MODE_S1, in.base, nullptr, dt, &A, &m->tendOld, nullptr, &lz, 2 * hs * ws, 4 * hn * ws))) return r;
  } else {
    if ((r = stage(m, pass, MODE_S1, O, &O, dt, &A, &m->tendOld, nullptr, nullptr, 2 * hs * ws, 4 * hn * ws))) return r;
  }
  if (!wide && (r = exchange_state(m, A, !slow))) return r;
  // tend(old) = L(new); new = old + dt/2 tend(old)
  if ((r = stage(m, pass, MODE_S2, A, &O, dt, &B, &m->tendOld, nullptr, nullptr, hs * ws, 2 * hn * ws))) return r;
  if (!wide && (r = exchange_state(m, B, !slow))) return r;
  // tend(new) = L(new); ip1 = <tend(old), tend(new)>, ip2 = <tend(new), tend(new)>
  if ((r = stage(m, pass, MODE_S3A, B, nullptr, 0.0, nullptr, Tn, &m->tendOld, nullptr, 0, 0, !fz && wide && m->fuse_push))) return r;
  m->fz_active = false;
  if (wide && m->p2p) m->need_fence = true;
  release_state(m, &B);
  if (wide && !m->fuse_push) {   // NCCL, or a peer run whose band-edge rows are filtered rows
    if ((r = join(m))) return r;
    if ((r = exchange_tend3(m, Tn->U, Tn->V, slow ? nullptr : Tn->gd, HALO_S, HALO_N))) return r;
  }
  if (defer_out) {
    // new = old + dt beta tend(new) is left to the next call: it needs the ghost rows of tend(new)
    const State tv = {Tn->U, Tn->V, Tn->gd};
    if (!wide && (r = exchange_state(m, tv, !slow))) return r;
    release_state(m, &A);
    out->base = O;
    if (!ownO) {  // the caller still owns in.base: take our own references
      retain(m, O.U);
      retain(m, O.V);
      retain(m, O.gd);
    }
    out->deferred = true;
    out->dts = dts;
    out->with_gd = !slow;
    out->L = Tn;
    return 0;
  }
  // new = old + dt beta tend(new)
  if ((r = update(m, O, *Tn, dts, 1, 0.0, !slow, &A, HALO_S * hs, HALO_N * hn))) return r;
  if (!wide && (r = exchange_state(m, A, !slow))) return r;
  if (ownO) release_state(m, &M);
  out->base = A;
  out->deferred = false;
  return 0;
}

static int runge_kutta(gmd_model *m, double dts, const State &in, int pass, State *out);
// csp2_splitting, src/dycore_mod.F90:671-687
static int csp2(gmd_model *m, const State &in, State *out) {
  int r;
  const double dtm = m->cfg.time_step_size;
  const double fast_dt = dtm / m->cfg.subcycles;
  static const bool no_lazy = getenv("GMD_NO_LAZY") != nullptr;
  const bool rk = (m->cfg.time_scheme == GMD_TIME_RUNGE_KUTTA);
  const bool lazy = !no_lazy && m->cfg.uv_adv_scheme != ADV_WENO && !rk;
  // the integrator slot of src/dycore_mod.F90:43-53
  auto integrator = [&](double dts, const Carry &cin, int pass, bool defer, Carry *cout) -> int {
    if (!rk) return predict_correct(m, dts, cin, pass, defer, cout);
    cout->deferred = false;
    return runge_kutta(m, dts, cin.base, pass, &cout->base);
  };
  Carry c0, c1;
  c0.base = in;
  if ((r = integrator(0.5 * dtm, c0, PASS_SLOW, lazy, &c1))) return r;
  for (int k = 0; k < m->cfg.subcycles; k++) {
    Carry c2;
    if ((r = integrator(fast_dt, c1, PASS_FAST, lazy, &c2))) return r;
    release_state(m, &c1.base);
    c1 = c2;
  }
  Carry c3;
  if ((r = integrator(0.5 * dtm, c1, PASS_SLOW, false, &c3))) return r;
  release_state(m, &c1.base);
  *out = c3.base;
  return 0;
}

static int axpby(gmd_model *m, double alpha, const Tend &x, double beta, Tend &y) {
  if (int rj = join(m)) return rj;
  const size_t total = (size_t)m->nr * m->geo.nlon;
  if (!m->dry) k_axpby3<<<m->ew_blocks, 256, 0, m->stream>>>(total, alpha, x.U, x.V, x.gd, beta, y.U, y.V, y.gd);
  return post_launch(m);
}
static int dot(gmd_model *m, const double *aU, const double *aV, const double *aG, const double *bU, const double *bV,
               const double *bG, int slot, int accumulate = 0) {
  if (int rj = join(m)) return rj;
  if (!m->dry) k_dot<<<m->ew_blocks, 256, 0, m->stream>>>(m->geo, m->tab, aU, aV, aG, bU, bV, bG, 1, m->d_partials, slot, accumulate);
  return post_launch(m);
}

// runge_kutta(dt, in -> *out, pass): the SPECIFIED extension of DESIGN.md section 8 (time_scheme = 'runge_kutta',
// params_mod.F90:40-44; the reference commit has no such integrator).  Explicit RK in increment form with the energy
// fix of predict_correct: out = in + beta dt K, K = sum b_i L(phi_i), beta = (sum of the stage tendency products that
// equal -3/dt <K, in> under the stage antisymmetries) / (3 <K, K>) -- see oracle/gmd_oracle.c runge_kutta.  Operator
// evaluations are the fused stage kernel in its store-the-tendency mode, the rest is the tendency algebra of isp.
static int runge_kutta(gmd_model *m, double dts, const State &in, int pass, State *out) {
  int r;
  const bool slow = (pass == PASS_SLOW);
  if (!m->tendA.U && (r = new_tend(m, &m->tendA))) return r;
  Tend *ka = &m->tendOld, *kb = &m->tendA;
  Tend &K = m->tendNew;
  const size_t bytes = (size_t)m->nr * m->geo.nlon * sizeof(double);
  State A, B;
  if ((r = new_state(m, &A, slow ? in.gd : nullptr))) return r;
  if ((r = new_state(m, &B, slow ? in.gd : nullptr))) return r;
  auto eval = [&](const State &s, Tend *k) -> int {   // k = L(s); a slow pass leaves dgd = 0 (src/dycore_mod.F90:297)
    int q;
    if (slow) {
      if ((q = join(m))) return q;
      if (!m->dry) CK(cudaMemsetAsync(k->gd, 0, bytes, m->stream));
    }
    return stage(m, pass, MODE_EVAL, s, nullptr, 0, nullptr, k, nullptr);
  };
  auto advance = [&](const Tend &t, double dt, State *to) -> int {
    int q;
    if ((q = update(m, in, t, dt, 0, 0, !slow, to))) return q;
    return exchange_state(m, *to, !slow);
  };
  auto tdot = [&](const Tend &x, const Tend &y, int slot, int acc) -> int {
    return dot(m, x.U, x.V, x.gd, y.U, y.V, y.gd, slot, acc);
  };
  if ((r = eval(in, ka))) return r;                          // k1
  if ((r = axpby(m, 1.0, *ka, 0.0, K))) return r;
  if (m->cfg.time_order == 4) {
    if ((r = advance(*ka, dts * 0.5, &A))) return r;
    if ((r = eval(A, kb))) return r;                         // k2
    if ((r = tdot(*ka, *kb, 0, 0))) return r;
    if ((r = axpby(m, 2.0, *kb, 1.0, K))) return r;
    if ((r = advance(*kb, dts * 0.5, &B))) return r;
    if ((r = eval(B, ka))) return r;                         // k3
    if ((r = tdot(*kb, *ka, 0, 1))) return r;
    if ((r = axpby(m, 2.0, *ka, 1.0, K))) return r;
    if ((r = advance(*ka, dts, &A))) return r;
    if ((r = eval(A, kb))) return r;                         // k4
    if ((r = tdot(*ka, *kb, 0, 1))) return r;
    if ((r = axpby(m, 1.0 / 6.0, *kb, 1.0 / 6.0, K))) return r;
  } else {
    if ((r = advance(*ka, dts, &A))) return r;
    if ((r = eval(A, ka))) return r;                         // k2
    if ((r = tdot(K, *ka, 0, 0))) return r;                  // <k1, k2>
    if ((r = axpby(m, 1.0, *ka, 1.0, K))) return r;
    if ((r = advance(K, dts * 0.25, &B))) return r;
    if ((r = eval(B, ka))) return r;                         // k3
    if ((r = tdot(K, *ka, 0, 1))) return r;                  // + <k1 + k2, k3>
    if ((r = axpby(m, 2.0 / 3.0, *ka, 1.0 / 6.0, K))) return r;
  }
  if ((r = tdot(K, K, 1, 0))) return r;
  {
    RedArgs ra = red_args(m);
    ra.tseq = tseq(m, "k_reduce_pairs.rk");
    if (!m->dry) k_reduce_pairs<<<1, 256, 0, m->stream>>>(m->d_partials, m->ew_blocks, m->d_ip, ra);
  }
  if ((r = post_launch(m))) return r;
  if ((r = allreduce2(m, m->d_ip))) return r;
  if ((r = update(m, in, K, dts, 3, dts, !slow, &A))) return r;
  if ((r = exchange_state(m, A, !slow))) return r;
  release_state(m, &B);
  *out = A;
  return 0;
}

// isp_splitting, src/dycore_mod.F90:689-752 (tend algebra on du, dv, dgd only)
static int isp(gmd_model *m, const State &F, State *out) {
  int r;
  const double dtm = m->cfg.time_step_size;
  const int S = m->cfg.subcycles;
  const double fast_dt = dtm / S, half_dt = dtm * 0.5;
  if (!m->tendA.U) {
    if ((r = new_tend(m, &m->tendA))) return r;
    if ((r = new_tend(m, &m->tendB))) return r;
  }
  Tend &slow = m->tendA, &acc = m->tendB, &T = m->tendOld, &T2 = m->tendNew;
  const size_t bytes = (size_t)m->nr * m->geo.nlon * sizeof(double);
  if ((r = join(m))) return r;
  // space_operators(slow) leaves dgd = 0 (:297)
  if (!m->dry) CK(cudaMemsetAsync(slow.gd, 0, bytes, m->stream));
  if ((r = stage(m, PASS_SLOW, MODE_EVAL, F, nullptr, 0, nullptr, &slow, nullptr))) return r;
  if ((r = axpby(m, 0.0, slow, 0.0, acc))) return r;  // acc = 0
  State P = F, P1, P2;
  bool ownP = false;
  for (int k = 0; k < S; k++) {
    if ((r = new_state(m, &P1, nullptr))) return r;
    if ((r = new_state(m, &P2, nullptr))) return r;
    if ((r = stage(m, PASS_FAST, MODE_EVAL, P, nullptr, 0, nullptr, &T, nullptr))) return r;
    if ((r = axpby(m, 1.0, slow, 1.0, T))) return r;
    if ((r = update(m, P, T, fast_dt * 0.5, 0, 0, true, &P1))) return r;
    if ((r = exchange_state(m, P1, true))) return r;
    if ((r = stage(m, PASS_FAST, MODE_EVAL, P1, nullptr, 0, nullptr, &T, nullptr))) return r;
    if ((r = axpby(m, 1.0, slow, 1.0, T))) return r;
    if ((r = update(m, P, T, fast_dt * 0.5, 0, 0, true, &P2))) return r;
    if ((r = exchange_state(m, P2, true))) return r;
    if ((r = stage(m, PASS_FAST, MODE_EVAL, P2, nullptr, 0, nullptr, &T2, nullptr))) return r;
    if ((r = axpby(m, 1.0, T2, 1.0, acc))) return r;
    if ((r = axpby(m, 1.0, slow, 1.0, T2))) return r;
    if ((r = update(m, P, T2, fast_dt, 0, 0, true, &P1))) return r;
    if ((r = exchange_state(m, P1, true))) return r;
    release_state(m, &P2);
    if (ownP) release_state(m, &P);
    P = P1;
    ownP = true;
  }
  if ((r = axpby(m, 0.0, slow, 2.0 / S, acc))) return r;  // acc *= 2/S
  State Q1, Q2;
  if ((r = new_state(m, &Q1, nullptr))) return r;
  if ((r = new_state(m, &Q2, nullptr))) return r;
  if ((r = join(m))) return r;
  if ((r = join(m))) return r;
  if (!m->dry) CK(cudaMemsetAsync(T.gd, 0, bytes, m->stream));
  if ((r = stage(m, PASS_SLOW, MODE_EVAL, P, nullptr, 0, nullptr, &T, nullptr))) return r;
  if ((r = axpby(m, -1.0, slow, 1.0, T))) return r;
  if ((r = update(m, P, T, half_dt, 0, 0, true, &Q1))) return r;
  if ((r = exchange_state(m, Q1, true))) return r;
  if ((r = join(m))) return r;
  if (!m->dry) CK(cudaMemsetAsync(T.gd, 0, bytes, m->stream));
  if ((r = stage(m, PASS_SLOW, MODE_EVAL, Q1, nullptr, 0, nullptr, &T, nullptr))) return r;
  if ((r = axpby(m, -1.0, slow, 1.0, T))) return r;
  if ((r = update(m, P, T, half_dt, 0, 0, true, &Q2))) return r;
  if ((r = exchange_state(m, Q2, true))) return r;
  if ((r = join(m))) return r;
  if (!m->dry) CK(cudaMemsetAsync(T2.gd, 0, bytes, m->stream));
  if ((r = stage(m, PASS_SLOW, MODE_EVAL, Q2, nullptr, 0, nullptr, &T2, nullptr))) return r;
  if ((r = axpby(m, 1.0, slow, 1.0, T2))) return r;
  if ((r = axpby(m, 1.0, acc, 1.0, T2))) return r;
  // ip1 = <R, F> (tend-state product, src/types_mod.F90:373-397), ip2 = <R, R>
  if ((r = dot(m, T2.U, T2.V, T2.gd, F.U, F.V, F.gd, 0))) return r;
  if ((r = dot(m, T2.U, T2.V, T2.gd, T2.U, T2.V, T2.gd, 1))) return r;
  {
    RedArgs ra = red_args(m);
    ra.tseq = tseq(m, "k_reduce_pairs.isp");
    if (!m->dry) k_reduce_pairs<<<1, 256, 0, m->stream>>>(m->d_partials, m->ew_blocks, m->d_ip, ra);
  }
  if ((r = post_launch(m))) return r;
  if ((r = allreduce2(m, m->d_ip))) return r;
  if ((r = update(m, F, T2, half_dt, 2, dtm, true, &Q1))) return r;
  if ((r = exchange_state(m, Q1, true))) return r;
  release_state(m, &Q2);
  if (ownP) release_state(m, &P);
  *out = Q1;
  return 0;
}

static int polar_filter_only(gmd_model *m, double *ud, double *vd, double *gdd) {
  if (!m->n_items[2]) return 0;
  PolarArgs p;
  memset(&p, 0, sizeof p);
  p.g = m->geo;
  p.t = m->tab;
  fill_items(p, m->items[2]);
  p.TU = ud; p.TV = vd; p.Tgd = gdd;
  p.rescale = 0;
  p.partials = m->d_partials;
  if (!m->dry) launch_polar(m, MODE_EVAL, m->n_items[2], p, m->stream);
  return post_launch(m);
}

// ordinary_diffusion(dt, state), src/diffusion_mod.F90:74-217.  *out is a fresh state.
static int diffusion(gmd_model *m, double dt, const State &in, State *out) {
  int r;
  const int norder = m->cfg.diffusion_order / 2;
  if (!m->d_ud) {
    if ((r = acquire(m, KIND_U, &m->d_ud))) return r;
    if ((r = acquire(m, KIND_V, &m->d_vd))) return r;
    if ((r = acquire(m, KIND_G, &m->d_gdd))) return r;
    if ((r = acquire(m, KIND_U, &m->d_ud2))) return r;
    if ((r = acquire(m, KIND_V, &m->d_vd2))) return r;
    if ((r = acquire(m, KIND_G, &m->d_gdd2))) return r;
  }
  if ((r = derive_uv(m, in))) return r;
  const bool south = (m->geo.r0 == 0), north = (m->geo.r1 == m->geo.nlat);
  const double *qu = m->w_u, *qv = m->w_v, *qg = in.gd;
  double *ou = m->d_ud, *ov = m->d_vd, *og = m->d_gdd;
  for (int order = 1; order <= norder; order++) {
    if ((r = halo_wait(m))) return r;
    if (!m->dry) {
      if (m->ew_rows) k_laplace2<<<dim3(m->ew2_bx, (unsigned)m->nr), EW2, 0, m->stream>>>(m->geo, m->tab, qu, qv, qg, ou, ov, og);
      else k_laplace<<<m->ew_blocks, 256, 0, m->stream>>>(m->geo, m->tab, qu, qv, qg, ou, ov, og);
    }
    if ((r = post_launch(m))) return r;
    if (south || north) {
      if (!m->dry) k_lap_pole<<<2, PT_EW, 0, m->stream>>>(m->geo, m->tab, qg, og, south ? 1 : 0, north ? 1 : 0);
      if ((r = post_launch(m))) return r;
    }
    if (order != norder) {
      if ((r = exchange_tend3(m, ou, ov, og, 1, 1))) return r;
      qu = ou; qv = ov; qg = og;
      ou = m->d_ud2; ov = m->d_vd2; og = m->d_gdd2;
    }
  }
  if ((r = polar_filter_only(m, ou, ov, og))) return r;
  if ((r = exchange_tend3(m, nullptr, nullptr, og, 1, 1))) return r;  // gdd(j+1) enters sqrt(gd') of row j+1
  const double sign = ((norder + 1) % 2 == 0) ? 1.0 : -1.0;
  const double sdc = sign * dt * m->cfg.diffusion_coef;
  State N;
  if ((r = new_state(m, &N, nullptr))) return r;
  if ((r = halo_wait(m))) return r;
  if (!m->dry) {
    if (m->ew_rows) k_diff_update2<<<dim3(m->ew2_bx, (unsigned)m->nr), EW2, 0, m->stream>>>(m->geo, m->w_u, m->w_v, in.gd, ou, ov, og, sdc, N.U, N.V, N.gd);
    else k_diff_update<<<m->ew_blocks, 256, 0, m->stream>>>(m->geo, m->w_u, m->w_v, in.gd, ou, ov, og, sdc, N.U, N.V, N.gd);
  }
  if ((r = post_launch(m))) return r;
  if ((r = exchange_state(m, N, true))) return r;
  *out = N;
  return 0;
}

// diag_run totals of state s into the ring slot of the device step counter (advanced first if `advance`)
static int diag(gmd_model *m, const State &s, int advance) {
  int r;
  if ((r = join(m))) return r;
  {
    const int ts = tseq(m, "k_diag");
    if (!m->dry) k_diag<<<m->ew_blocks, 256, 0, m->stream>>>(m->geo, m->tab, s.U, s.V, s.gd, m->ghs, m->mesh.dlon, m->mesh.dlat,
                                               m->d_partials, ts);
  }
  if ((r = post_launch(m))) return r;
  {
    RedArgs ra = red_args(m);
    ra.tseq = tseq(m, "k_reduce_pairs.diag");
    if (!m->dry) k_reduce_pairs<<<1, 256, 0, m->stream>>>(m->d_partials, m->ew_blocks, m->d_sums, ra);
  }
  if ((r = post_launch(m))) return r;
  if ((r = allreduce2(m, m->d_sums))) return r;
  {
    const int ts = tseq(m, "k_diag_store");
    if (!m->dry) k_diag_store<<<1, 32, 0, m->stream>>>(m->d_sums, m->d_beta, m->mesh.radius, m->d_ring, m->d_ctr, advance, gmd_model::RING, ts);
  }
  if ((r = post_launch(m))) return r;
  return unit_end(m);
}

// time_integrate (src/dycore_mod.F90:654-669) + time_advance + diag_run for ONE step; consumes m->cur
static int one_step_body(gmd_model *m);
static int one_step(gmd_model *m) {
  for (int q = 0; q < 3; q++) std::sort(m->free_[q].begin(), m->free_[q].end());   // canonical pool order
  m->in_step = true;
  m->step_launch0 = m->launches;
  const int r = one_step_body(m);
  m->in_step = false;
  return r;
}
static int one_step_body(gmd_model *m) {
  int r;
  State next;
  switch (m->cfg.split_scheme) {
    case GMD_SPLIT_CSP2: r = csp2(m, m->cur, &next); break;
    case GMD_SPLIT_ISP: r = isp(m, m->cur, &next); break;
    default: {
      if (m->cfg.time_scheme == GMD_TIME_RUNGE_KUTTA) {
        r = runge_kutta(m, m->cfg.time_step_size, m->cur, PASS_ALL, &next);
        break;
      }
      Carry ci, co;
      ci.base = m->cur;
      r = predict_correct(m, m->cfg.time_step_size, ci, PASS_ALL, false, &co);
      next = co.base;
    }
  }
  if (r) return r;
  release_state(m, &m->cur);
  m->cur = next;
  if (m->cfg.use_diffusion) {
    State d;
    if ((r = diffusion(m, m->cfg.time_step_size, m->cur, &d))) return r;
    release_state(m, &m->cur);
    m->cur = d;
  }
  m->step++;
  // canonical pool order: the next step's buffer choice depends on `cur` only (graph keys stay few)
  for (int q = 0; q < 3; q++) std::sort(m->free_[q].begin(), m->free_[q].end());
  return diag(m, m->cur, 1);
}

// ---------------------------------------------------------------------------------------------------------
// host <-> device field transfer
// ---------------------------------------------------------------------------------------------------------
// Host arrays are pageable and borrowed, so every field goes through a pinned double buffer on its own copy
// stream ("lane"): a host thread per field packs chunk k+1 while the DMA engine moves chunk k, and the fields of
// one set/get call travel concurrently.  This is what the end-to-end number (bench.py "e2e") pays per call.
struct XferJob {
  bool up;                 // host -> device (else device -> host)
  const double *hsrc;      // up: global host array or NULL (zeros)
  double *hdst;            // down: global host array
  double *dev;             // device field pointer (row r0)
  int layout, nrows_valid;
  bool zero_poles;
  cudaError_t err;
};
static int lanes_init(gmd_model *m) {
  if (m->lanes_ready) return 0;
  const int nlon = m->geo.nlon;
  m->lane_rows = std::max(1, (int)((size_t)(4u << 20) / sizeof(double) / (size_t)nlon));
  m->lane_rows = std::min(m->lane_rows, m->nr + 2 * GHOST);
  for (auto &L : m->lanes) {
    CK(cudaStreamCreateWithFlags(&L.s, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
      CK(cudaMallocHost(&L.pin[k], (size_t)m->lane_rows * nlon * sizeof(double)));
      CK(cudaEventCreateWithFlags(&L.ev[k], cudaEventDisableTiming));
    }
  }
  m->lanes_ready = true;
  return 0;
}
static inline const double *host_row(const double *base, int layout, int nlon, int j) {
  return (layout == GMD_LAYOUT_REFERENCE) ? base + (size_t)(j + 2) * (nlon + 4) + 2 : base + (size_t)j * nlon;
}
static void lane_upload(gmd_model *m, gmd_model::Lane &L, XferJob &job) {
#define CKL(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { job.err = e_; return; } } while (0)
  CKL(cudaSetDevice(m->dev));
  const int nlon = m->geo.nlon, r0 = m->geo.r0;
  double *base = job.dev - (size_t)GHOST * nlon;
  // rows outside the globe (and a missing field) are zeros
  CKL(cudaMemsetAsync(base, 0, m->fld_elems * sizeof(double), L.s));
  if (job.hsrc) {
    const int jlo = std::max(0, r0 - GHOST), jhi = std::min(job.nrows_valid, m->geo.r1 + GHOST);
    int k = 0;
    for (int ja = jlo; ja < jhi; ja += m->lane_rows, k++) {
      const int jb = std::min(jhi, ja + m->lane_rows), b = k & 1;
      if (k >= 2) CKL(cudaEventSynchronize(L.ev[b]));
      if (job.layout == GMD_LAYOUT_COMPACT) {
        memcpy(L.pin[b], job.hsrc + (size_t)ja * nlon, (size_t)(jb - ja) * nlon * sizeof(double));
      } else {
        for (int j = ja; j < jb; j++)
          memcpy(L.pin[b] + (size_t)(j - ja) * nlon, host_row(job.hsrc, job.layout, nlon, j), (size_t)nlon * sizeof(double));
      }
      CKL(cudaMemcpyAsync(base + (size_t)(ja - (r0 - GHOST)) * nlon, L.pin[b], (size_t)(jb - ja) * nlon * sizeof(double),
                          cudaMemcpyHostToDevice, L.s));
      CKL(cudaEventRecord(L.ev[b], L.s));
    }
  }
  CKL(cudaStreamSynchronize(L.s));
}
static void lane_unpack(const gmd_model *m, const XferJob &job, const double *pin, int ja, int jb) {
  const int nlon = m->geo.nlon;
  for (int j = ja; j < jb; j++) {
    const double *row = pin + (size_t)(j - ja) * nlon;
    const bool zero = job.zero_poles && (j == 0 || j == m->geo.nlat - 1);
    if (job.layout == GMD_LAYOUT_REFERENCE) {
      double *d = job.hdst + (size_t)(j + 2) * (nlon + 4);
      if (zero) memset(d + 2, 0, (size_t)nlon * sizeof(double));
      else memcpy(d + 2, row, (size_t)nlon * sizeof(double));
      d[0] = d[nlon];      // parallel_fill_halo, src/parallel_mod.F90:466-524
      d[1] = d[nlon + 1];
      d[nlon + 2] = d[2];
      d[nlon + 3] = d[3];
    } else {
      double *d = job.hdst + (size_t)j * nlon;
      if (zero) memset(d, 0, (size_t)nlon * sizeof(double));
      else memcpy(d, row, (size_t)nlon * sizeof(double));
    }
  }
}
static void lane_download(gmd_model *m, gmd_model::Lane &L, XferJob &job) {
  CKL(cudaSetDevice(m->dev));
  const int nlon = m->geo.nlon, r0 = m->geo.r0;
  const int jlo = r0, jhi = std::min(job.nrows_valid, m->geo.r1);
  int k = 0, pja = 0, pjb = 0;
  for (int ja = jlo; ja < jhi; ja += m->lane_rows, k++) {
    const int jb = std::min(jhi, ja + m->lane_rows), b = k & 1;
    CKL(cudaMemcpyAsync(L.pin[b], job.dev + (size_t)(ja - r0) * nlon, (size_t)(jb - ja) * nlon * sizeof(double),
                        cudaMemcpyDeviceToHost, L.s));
    CKL(cudaEventRecord(L.ev[b], L.s));
    if (k >= 1) {  // unpack the previous chunk while this one is in flight
      CKL(cudaEventSynchronize(L.ev[b ^ 1]));
      lane_unpack(m, job, L.pin[b ^ 1], pja, pjb);
    }
    pja = ja;
    pjb = jb;
  }
  if (k >= 1) {
    CKL(cudaEventSynchronize(L.ev[(k - 1) & 1]));
    lane_unpack(m, job, L.pin[(k - 1) & 1], pja, pjb);
  }
#undef CKL
}
// run up to 4 transfers concurrently; the compute streams are drained first and the call returns when all are done
static int run_xfers(gmd_model *m, XferJob *jobs, int n) {
  int r;
  if ((r = lanes_init(m))) return r;
  if ((r = join(m))) return r;
  CK(cudaStreamSynchronize(m->stream));
  std::thread th[4];
  for (int k = 0; k < n; k++) {
    jobs[k].err = cudaSuccess;
    if (k == n - 1) {  // the calling thread takes the last one
      if (jobs[k].up) lane_upload(m, m->lanes[k], jobs[k]);
      else lane_download(m, m->lanes[k], jobs[k]);
    } else {
      th[k] = std::thread([m, jobs, k]() {
        if (jobs[k].up) lane_upload(m, m->lanes[k], jobs[k]);
        else lane_download(m, m->lanes[k], jobs[k]);
      });
    }
  }
  for (int k = 0; k + 1 < n; k++) th[k].join();
  for (int k = 0; k < n; k++)
    if (jobs[k].err != cudaSuccess)
      return fail(GMD_ERR_CUDA, "CUDA error %s in host<->device field transfer (%s)", cudaGetErrorName(jobs[k].err),
                  cudaGetErrorString(jobs[k].err));
  return 0;
}
static XferJob up_job(const double *src, int layout, int nrows_valid, double *dst) {
  XferJob j = {true, src, nullptr, dst, layout, nrows_valid, false, cudaSuccess};
  return j;
}
// up to 3 device fields -> global host arrays (NULL destinations are skipped)
static int download_fields(gmd_model *m, int n, double *const dev[], double *const host[], const int nrows_valid[],
                           const bool zero_poles[], int layout) {
  XferJob jobs[4];
  int nj = 0;
  for (int k = 0; k < n; k++)
    if (host[k]) jobs[nj++] = XferJob{false, nullptr, host[k], dev[k], layout, nrows_valid[k], zero_poles[k], cudaSuccess};
  return nj ? run_xfers(m, jobs, nj) : 0;
}

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

const char *gmd_last_error(void) { return g_err; }
int gmd_version(void) { return GMD_VERSION; }

void gmd_config_defaults(gmd_config *c) {
  memset(c, 0, sizeof *c);
  c->subcycles = 4;
  c->uv_adv_upwind_lon_beta = 0.0;
  c->uv_adv_upwind_lat_beta = 0.5;
  c->use_zonal_tend_filter = 1;
  c->diffusion_order = 2;
  c->split_scheme = GMD_SPLIT_CSP2;
  c->rank = 0;
  c->nranks = 1;
  c->device = -1;
  c->time_scheme = GMD_TIME_PREDICT_CORRECT;
  c->time_order = 3;
}

void gmd_destroy(gmd_model *m) {
  if (!m) return;
  cudaSetDevice(m->dev);
  if (m->stream2) cudaStreamSynchronize(m->stream2);
  if (m->stream) cudaStreamSynchronize(m->stream);
  for (auto &g : m->graphs) cudaGraphExecDestroy(g.exec);
  if (m->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comm);
  for (void *p : m->ipc_opened) cudaIpcCloseMemHandle(p);
  cudaFree(m->slab);
  cudaFree(m->page);
  for (double *p : m->tab_allocs) cudaFree(p);
  cudaFree(m->d_flags_alloc);
  cudaFree(m->d_basis);
  cudaFree(m->d_rot);
  cudaFree(m->d_bar);
  cudaFree(m->d_partials);
  cudaFree(m->d_ip);
  cudaFree(m->d_ring);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  for (auto e : m->evpool) cudaEventDestroy(e);
  for (auto &L : m->lanes) {
    for (int k = 0; k < 2; k++) {
      if (L.pin[k]) cudaFreeHost(L.pin[k]);
      if (L.ev[k]) cudaEventDestroy(L.ev[k]);
    }
    if (L.s) cudaStreamDestroy(L.s);
  }
  if (m->stream2) cudaStreamDestroy(m->stream2);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  delete m;
}

int gmd_create(const gmd_config *cfg, gmd_model **out) {
  if (!cfg || !out) return fail(GMD_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->num_lon < 4 || cfg->num_lat < 5) return fail(GMD_ERR_ARG, "grid too small: %d x %d", cfg->num_lon, cfg->num_lat);
  if (cfg->num_lon % 2) return fail(GMD_ERR_ARG, "num_lon must be even (16-byte column pairs); got %d", cfg->num_lon);
  if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks) return fail(GMD_ERR_ARG, "bad rank %d of %d", cfg->rank, cfg->nranks);
  if (cfg->uv_adv_scheme < 0 || cfg->uv_adv_scheme > 2)
    return fail(GMD_ERR_ARG, "Unknown uv_adv_scheme %d!", cfg->uv_adv_scheme);  // dycore_mod.F90:104-106
  if (cfg->subcycles < 1) return fail(GMD_ERR_ARG, "subcycles must be >= 1");
  if (cfg->time_scheme != GMD_TIME_PREDICT_CORRECT && cfg->time_scheme != GMD_TIME_RUNGE_KUTTA)
    return fail(GMD_ERR_ARG, "Unknown time_scheme %d!", cfg->time_scheme);  // dycore_mod.F90:78-83
  if (cfg->time_scheme == GMD_TIME_RUNGE_KUTTA && cfg->time_order != 3 && cfg->time_order != 4)
    return fail(GMD_ERR_ARG, "runge_kutta: time_order must be 3 or 4, got %d", cfg->time_order);
  if (cfg->use_diffusion && cfg->diffusion_order != 2 && cfg->diffusion_order != 4)
    return fail(GMD_ERR_ARG, "diffusion_order must be 2 or 4");
  {  // FFTPACK accepts any n, but the reference's grids are 2^a 3^b 5^c; the projector needs no factorisation
    for (int k = 0; k < 20; k++) {
      const int c = cfg->zonal_tend_filter_cutoff_wavenumber[k];
      if (c < 0 || c > 254) return fail(GMD_ERR_ARG, "zonal_tend_filter_cutoff_wavenumber(%d)=%d outside 0..254", k + 1, c);
    }
  }
  if (cfg->num_lat / cfg->nranks < 4) return fail(GMD_ERR_ARG, "fewer than 4 latitude rows per rank");
  if (cfg->polar_band_rows < 0 || (cfg->polar_band_rows > 0 && cfg->polar_band_rows < 4))
    return fail(GMD_ERR_ARG, "polar_band_rows must be 0 or >= 4");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(GMD_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  gmd_model *m = new gmd_model();
  m->cfg = *cfg;
  if (cfg->device >= 0) m->dev = cfg->device;
  else cudaGetDevice(&m->dev);
#define CKD(call)                                                                                         \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) {                                                                              \
      fail(GMD_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__,         \
           cudaGetErrorString(e_));                                                                       \
      gmd_destroy(m);                                                                                     \
      return GMD_ERR_CUDA;                                                                                \
    }                                                                                                     \
  } while (0)
  CKD(cudaSetDevice(m->dev));
  CKD(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
  m->stream = m->own_stream;
  CKD(cudaEventCreate(&m->ev0));
  CKD(cudaEventCreate(&m->ev1));
  const int nlon = cfg->num_lon, nlat = cfg->num_lat;
  m->mesh.init(nlon, nlat, /*reset_poles=*/true);
  m->mesh.filter_init(cfg->use_zonal_tend_filter != 0, cfg->zonal_tend_filter_cutoff_wavenumber);
  {   // moving reduced tendency (specified extension)
    const int bad = m->mesh.reduce_init(cfg->use_zonal_reduce != 0, cfg->zonal_reduce_factors);
    if (bad) {
      fail(GMD_ERR_ARG, "zonal_reduce_factors(%d)=%d must be >= 0 and divide num_lon=%d", bad, cfg->zonal_reduce_factors[bad - 1], nlon);
      gmd_destroy(m);
      return GMD_ERR_ARG;
    }
    for (int j = 0; j < nlat; j++)
      if ((m->mesh.red_full[(size_t)j] > 1 && m->mesh.flag_full[(size_t)j]) || (m->mesh.red_half[(size_t)j] > 1 && m->mesh.flag_half[(size_t)j])) {
        fail(GMD_ERR_ARG, "row %d is both a zonal filter row and a reduced row", j + 1);
        gmd_destroy(m);
        return GMD_ERR_ARG;
      }
    if (cfg->use_zonal_reduce && nlon > PQ * PT) {
      fail(GMD_ERR_ARG, "use_zonal_reduce: num_lon must be <= %d", PQ * PT);
      gmd_destroy(m);
      return GMD_ERR_ARG;
    }
  }
  // latitude bands: rows split as evenly as possible, or (polar_band_rows) shorter first and last bands
  m->geo.nlon = nlon;
  m->geo.nlat = nlat;
  band_rows(nlat, cfg->nranks, cfg->polar_band_rows, cfg->rank, &m->geo.r0, &m->geo.r1);
  m->nr = m->geo.r1 - m->geo.r0;
  m->fld_elems = (size_t)(m->nr + 2 * GHOST) * nlon;
  {
    // field pool: csp2 needs 22 fields (state 3 + ghs + 2 tendencies + 2 stage states + 1 substep state), isp 6
    // more, WENO 12, diffusion 8, the get_* calls 2
    m->slab_cap = 56;
    if (const char *ev = getenv("GMD_POOL_FIELDS")) m->slab_cap = std::max(24, atoi(ev));
    const size_t bytes = (size_t)m->slab_cap * m->fld_elems * sizeof(double);
    CKD(cudaMalloc(&m->slab, bytes));
    CKD(cudaMemset(m->slab, 0, bytes));
    CKD(cudaMalloc(&m->page, SP_WORDS * sizeof(u64)));
    CKD(cudaMemset(m->page, 0, SP_WORDS * sizeof(u64)));
    m->peer_page[cfg->rank < MAXR ? cfg->rank : 0] = m->page;
    if (const char *ev = getenv("GMD_PEER_TIMEOUT_S")) {
      const unsigned long long ns = (unsigned long long)std::max(1.0, atof(ev)) * 1000000000ull;
      CKD(cudaMemcpyToSymbol(g_spin_timeout_ns, &ns, sizeof ns));
    }
  }
  int r = build_tables(m);
  if (r) { gmd_destroy(m); return r; }
  // stage grid: ~6 CTAs per SM
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, m->dev);
  {
    // boundary / interior split.  Boundary rows are those whose evaluation reads rows produced by the polar-row
    // kernel (filtered rows, pole caps) or received from a neighbour band.
    int K = 0;
    if (cfg->use_zonal_tend_filter)
      for (int k = 0; k < 20; k++)
        if (cfg->zonal_tend_filter_cutoff_wavenumber[k]) K = k + 1;
    if (cfg->use_zonal_reduce)
      for (int k = 0; k < 20; k++)
        if (cfg->zonal_reduce_factors[k] > 1) K = std::max(K, k + 1);
    m->bs = (m->geo.r0 == 0) ? K + 2 : 0;
    m->bn = (m->geo.r1 == nlat) ? K + 3 : 0;
    // wide-halo predict_correct: every band edge needs HALO_N plain rows (no filtered row, no pole row) on both sides
    // -- the rows a neighbour recomputes / receives.  Decided from the configuration alone, so every rank agrees.
    m->wide = cfg->nranks > 1 && cfg->uv_adv_scheme != GMD_ADV_WENO && cfg->time_scheme == GMD_TIME_PREDICT_CORRECT;
    for (int q = 1; q < cfg->nranks && m->wide; q++) {
      int e0, e1;
      band_rows(nlat, cfg->nranks, cfg->polar_band_rows, q, &e0, &e1);   // edge between band q-1 and band q: row e0
      for (int j = e0 - HALO_N; j < e0 + HALO_N; j++)
        if (j < 1 || j > nlat - 2 || m->mesh.flag_full[(size_t)j] || m->mesh.flag_half[(size_t)j] ||
            m->mesh.red_full[(size_t)j] > 1 || m->mesh.red_half[(size_t)j] > 1) m->wide = false;
      if (e1 - e0 < HALO_N) m->wide = false;
    }
    if (getenv("GMD_NO_WIDE")) m->wide = false;
    // the bands that hold polar rows split off the rows next to their pole, so that the polar-row kernel that follows
    // them overlaps the rest of the sweep
    // fused polar cap: every item must be one the thread-owned projector handles (cutoff < KF, nlon % 4 == 0, at most
    // PQ element groups per thread)
    m->cap = (m->n_items[0] + m->n_items[1] > 0) && (nlon % 4 == 0) && ((nlon / 4 + PT - 1) / PT <= PQ) &&
             cfg->uv_adv_scheme != GMD_ADV_WENO && getenv("GMD_CAP") != nullptr;   // opt-in: measured slower
    for (int li = 0; li < 2 && m->cap; li++)
      for (unsigned pk : m->items[li]) {
        const int cutoff = (int)((pk >> 16) & 0x1ffu) - 1, kind = (int)((pk >> 28) & 7u);
        const int Kk = cutoff + 1;
        if (kind != IT_POLE_S && kind != IT_POLE_N && !(Kk >= 1 && Kk <= KF && 2 * Kk < nlon)) m->cap = false;
        if (pk & ITEM_REDUCE) m->cap = false;
      }
    // the three-sweep path of a single band splits as well: the rows next to the poles first, the polar rows behind them,
    // the other rows beside both on stream2 (measured on B200, ms per step without / with: 360x181 0.416 / 0.321,
    // 1440x721 0.710 / 0.685, 3600x1801 4.18 / 4.11 -- round 1 saw no gain because every launch then paid a full gap)
    // (same schemes as the wide-halo bands: with WENO the advection sweeps between two stages read every row)
    const bool split_n1 = cfg->nranks == 1 && cfg->uv_adv_scheme != GMD_ADV_WENO && cfg->time_scheme == GMD_TIME_PREDICT_CORRECT;
    m->split = (m->wide || split_n1) && (m->bs + m->bn > 0);
    if (const char *ev = getenv("GMD_NO_SPLIT")) m->split = (atoi(ev) == 0) && (cfg->nranks == 1 || m->wide);
    if (m->bs + m->bn >= m->nr) m->split = false;
    // measured on two 226-row polar bands (profiles/r2_i_*): 1 -> 0.915 ms per step, 0 -> 0.962, 3 -> 1.015
    m->pdl = 1;
    if (const char *ev = getenv("GMD_PDL")) m->pdl = atoi(ev);
    const int nstrips = (nlon + WOUT - 1) / WOUT;
    m->nbx = (nstrips + SW - 1) / SW;
    // interior (or whole band): one wave of CTAs, as many row chunks as the resident-CTA slots allow
    int per_sm = 0;
    const int pass0 = (cfg->split_scheme == GMD_SPLIT_CSP2 || cfg->split_scheme == GMD_SPLIT_ISP) ? PASS_FAST : PASS_ALL;
    {   // every stage-kernel variant this configuration can launch (not during a graph capture later on)
      const int adv = cfg->uv_adv_scheme;
      bool bad = false;
      for (int pass = 0; pass < 3 && !bad; pass++) {
        for (int mode = 0; mode < 4; mode++) bad = bad || allow_smem((const void *)pick_stage(pass, adv, mode));
        bad = bad || allow_smem((const void *)pick_stage(pass, adv, MODE_S3A, true));
        if (adv != ADV_WENO)
          for (int lz = 1; lz <= 2; lz++) bad = bad || allow_smem((const void *)pick_stage_lazy(pass, adv, lz));
        if (m->cap)
          for (int mode = 0; mode < 4; mode++)
            for (int lz = 0; lz <= (mode == MODE_S1 ? 2 : 0); lz++) bad = bad || allow_smem((const void *)pick_cap(pass, adv, mode, lz));
      }
      if (bad) { gmd_destroy(m); return GMD_ERR_CUDA; }
    }
    CKD(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pick_stage(pass0, cfg->uv_adv_scheme, MODE_S2), BX,
                                                      stage_smem_bytes(64)));
    per_sm = std::max(per_sm, 1);
    if (const char *ev = getenv("GMD_CTAS_PER_SM")) per_sm = std::max(1, atoi(ev));
    const int slots = nsm * per_sm;
    // row records + packet ring of four resident CTAs must fit the SM's shared memory (taller bands: several waves)
    const int max_rows = GMD_RING ? 96 : 48 * 1024 / (RC_N * (int)sizeof(double)) - 2;
    const int rows_i = m->split ? m->nr - m->bs - m->bn : m->nr;
    int want = std::max(1, slots / m->nbx);
    m->rows_per_cta = std::min(max_rows, std::max(8, (rows_i + want - 1) / want));
    if (const char *ev = getenv("GMD_ROWS_PER_CTA")) m->rows_per_cta = std::min(max_rows, std::max(1, atoi(ev)));
    m->nchunks = (m->nr + m->rows_per_cta - 1) / m->rows_per_cta;
    m->nchunks_i = (rows_i + m->rows_per_cta - 1) / m->rows_per_cta;
    m->stage_smem = stage_smem_bytes(m->rows_per_cta);
    {   // S3a geometry from the occupancy of the S3a kernel itself
      int ps = 0;
      CKD(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, pick_stage(pass0, cfg->uv_adv_scheme, MODE_S3A), BX,
                                                        stage_smem_bytes(64)));
      ps = std::max(ps, 1);
      if (const char *ev = getenv("GMD_CTAS_PER_SM")) ps = std::max(1, atoi(ev));
      const int want3 = std::max(1, nsm * ps / m->nbx);
      m->rows_per_cta_s3a = std::min(max_rows, std::max(8, (rows_i + want3 - 1) / want3));
      if (const char *ev = getenv("GMD_ROWS_PER_CTA")) m->rows_per_cta_s3a = std::min(max_rows, std::max(1, atoi(ev)));
    }
    // boundary chunks: long while the interior launch hides them, short once the boundary -> polar rows -> next
    // boundary chain is what a phase waits for (measured at 900 and 225 rows per rank)
    m->rows_per_cta_b = (m->nr >= 600) ? 6 : 2;   // 226-row polar bands: 1 -> 0.967, 2 -> 0.904, 3 -> 0.916, 4 -> 0.913 ms per step
    if (const char *ev = getenv("GMD_ROWS_PER_CTA_B")) m->rows_per_cta_b = std::max(1, atoi(ev));
    m->nchunks_b = (std::max(m->bs, m->bn) + m->rows_per_cta_b - 1) / m->rows_per_cta_b;
    m->stage_smem_b = stage_smem_bytes(m->rows_per_cta_b);
    if (m->cap) {
      // the launch must be co-resident (grid barrier): as many clusters as there are items, within the device's limit
      int lo = 0, hi = 0;
      CKD(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      m->prio_hi = hi;
      CKD(cudaMalloc(&m->d_bar, 2 * sizeof(unsigned)));
      CKD(cudaMemset(m->d_bar, 0, 2 * sizeof(unsigned)));
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(CL * 64, 1, 1);
      lc.blockDim = dim3(BX, 1, 1);
      lc.dynamicSmemBytes = m->stage_smem_b;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at;
      lc.numAttrs = 1;
      int ncl = 0;
      cudaError_t ce = cudaOccupancyMaxActiveClusters(&ncl, pick_cap(pass0, cfg->uv_adv_scheme, MODE_S3A, 0), &lc);
      if (ce != cudaSuccess) { cudaGetLastError(); ncl = 0; }
      int lim = (ncl * 3 / 4) * CL;   // leave room: other kernels of the step hold slots while the cap CTAs arrive
      if (const char *ev = getenv("GMD_CAP_CTAS")) lim = std::min(ncl * CL, std::max(CL, atoi(ev) / CL * CL));
      const int n_march = (2 * m->nbx * m->nchunks_b + CL - 1) / CL * CL;
      m->cap_ctas = lim;
      if (n_march > lim) m->cap = false, m->split = m->split && m->wide;
    }
    // fused predict_correct (k_pc): product build, predict_correct with centred / upwind advection, one band or wide-halo
    // bands.  The fused rows keep (2 south, 4 north) plain rows between themselves and every row that takes part in a
    // full-row operation (filter rows, reduced rows, pole rows): S1 of the wavefront covers that many rows more.
    m->fused = !GMD_STRICT && cfg->uv_adv_scheme != GMD_ADV_WENO && cfg->time_scheme == GMD_TIME_PREDICT_CORRECT &&
               (cfg->nranks == 1 || m->wide);
    // GMD_FUSED=1 / 0 forces the fused kernel on (where it applies) / off; unset: on where it pays (below)
    const char *fused_env = getenv("GMD_FUSED");
    if (fused_env && atoi(fused_env) == 0) m->fused = false;
    if (m->fused) {
      auto flagged = [&](int j) {
        return j <= 0 || j >= nlat - 1 || m->mesh.flag_full[(size_t)j] || m->mesh.flag_half[(size_t)j] ||
               m->mesh.red_full[(size_t)j] > 1 || m->mesh.red_half[(size_t)j] > 1;
      };
      int I0 = m->geo.r0, I1 = m->geo.r1;
      for (int j = m->geo.r0; j < m->geo.r1; j++)
        if (flagged(j)) {
          if (j < nlat / 2) I0 = std::max(I0, j + 3);
          else I1 = std::min(I1, j - 4);
        }
      // a flagged row of the other hemisphere's cap inside [I0, I1) (a band that spans both caps with rows flagged
      // far from the poles) cannot happen: caps are contiguous from their pole
      for (int j = std::max(I0 - 2, 0); j < std::min(I1 + 4, nlat) && m->fused; j++)
        if (flagged(j)) m->fused = false;
      if (I1 - I0 < 6) m->fused = false;
      // a band of a multi-band run that holds polar rows keeps the three-sweep path: its step time is the chain
      // sweep -> polar rows -> sweep ..., and the polar-row CTAs (one SM each) cannot start while k_pc owns the SMs
      // Where it pays.  A band without polar rows: always (one launch per predict_correct instead of three).  A band
      // with polar rows: the polar-row CTAs (one SM each) cannot start while k_pc owns the SMs, so the chain sweep -> polar
      // rows -> sweep ... of the rows next to the poles finishes AFTER k_pc instead of beside the interior sweeps; that
      // costs 20-50 us per predict_correct and is only won back on a big band.  Measured on B200, ms per step, three
      // sweeps -> fused: 3600x1801 4.18 -> 3.44, 7200x3601 16.0 -> 12.9, two 900-row bands of 3600 2.39 -> 2.22, but
      // 1440x721 0.71 -> 0.90 and 400-row polar bands of 3600 (4 bands) 1.34 -> 1.49.
      if ((I0 > m->geo.r0 || I1 < m->geo.r1) && !(fused_env && atoi(fused_env) != 0)) {
        if ((double)(I1 - I0) * (double)nlon < 2.0e6) m->fused = false;
      }
      m->fz_I0 = I0;
      m->fz_I1 = I1;
    }
    if (m->fused) {
      const int adv = cfg->uv_adv_scheme;
      stage_fn f0 = pick_pc(pass0, adv, 1, false);
      int ps = 0;
      CKD(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, f0, PC_BX, PC_SMEM_BYTES));
      ps = std::max(ps, 1);
      if (const char *ev = getenv("GMD_PC_CTAS_PER_SM")) ps = std::max(1, atoi(ev));
      m->fz_nstrips = (nlon + WOUT3 - 1) / WOUT3;
      const int rows = m->fz_I1 - m->fz_I0;
      const int want = std::max(1, nsm * ps / m->fz_nstrips);
      m->fz_rpc = std::max(4, (rows + want - 1) / want);
      if (const char *ev = getenv("GMD_PC_ROWS")) m->fz_rpc = std::max(1, atoi(ev));
      m->fz_nchunks = (rows + m->fz_rpc - 1) / m->fz_rpc;
      const int tall = std::max(m->fz_I0 > m->geo.r0 ? m->fz_I0 - m->geo.r0 + 4 : 0, m->fz_I1 < m->geo.r1 ? m->geo.r1 - m->fz_I1 + 2 : 0);
      m->fz_ncb = (tall + m->rows_per_cta_b - 1) / m->rows_per_cta_b;
    }
    CKD(cudaStreamCreateWithFlags(&m->stream2, cudaStreamNonBlocking));
    m->evpool.resize(64);
    for (auto &e : m->evpool) CKD(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const size_t total = (size_t)m->nr * nlon;
  m->ew_blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)nsm * 8);
  // ordinary_diffusion and the derived u, v: the row-pair kernels (k_derive2 / k_laplace2 / k_diff_update2, one row per
  // blockIdx.y) on a single band; GMD_EW_ROWS=0/1 forces the element-indexed / the row-pair form
  m->ew2_bx = (unsigned)((nlon / 2 + EW2 - 1) / EW2);
  m->ew_rows = (cfg->nranks == 1) && (m->nr + 2 <= 65535);
  if (const char *ev = getenv("GMD_EW_ROWS")) m->ew_rows = (atoi(ev) != 0) && (m->nr + 2 <= 65535);
  {
    const int rmin = std::max(1, std::min(m->rows_per_cta, m->rows_per_cta_s3a));
    const int cmax = (m->nr + rmin - 1) / rmin + 2;
    m->n_partials = std::max(m->nbx * (2 * cmax + 2 * m->nchunks_b) + m->n_items[0] + m->n_items[1], m->ew_blocks) + 16;
    if (m->fused) m->n_partials = std::max(m->n_partials, 2 * m->nbx * m->fz_ncb + m->fz_nstrips * m->fz_nchunks + m->n_items[0] + m->n_items[1] + 16);
  }
  CKD(cudaMalloc(&m->d_partials, (size_t)m->n_partials * 2 * sizeof(double)));
  CKD(cudaMalloc(&m->d_ip, 8 * sizeof(double)));
  CKD(cudaMemset(m->d_ip, 0, 8 * sizeof(double)));
  m->d_sums = m->d_ip + 2;
  m->d_beta = m->d_ip + 4;
  m->d_ctr = (int *)(m->d_ip + 6);
  {
    const double one = 1.0;
    CKD(cudaMemcpy(m->d_beta, &one, sizeof one, cudaMemcpyHostToDevice));
  }
  CKD(cudaMalloc(&m->d_ring, (size_t)gmd_model::RING * 3 * sizeof(double)));
  CKD(cudaMemset(m->d_ring, 0, (size_t)gmd_model::RING * 3 * sizeof(double)));
  CKD(cudaFuncSetAttribute(k_polar<MODE_S1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CKD(cudaFuncSetAttribute(k_polar<MODE_S2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CKD(cudaFuncSetAttribute(k_polar<MODE_S3A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CKD(cudaFuncSetAttribute(k_polar<MODE_EVAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  if (2 * (size_t)nlon * sizeof(double) > 200 * 1024) { fail(GMD_ERR_ARG, "num_lon too large for the polar-row kernel"); gmd_destroy(m); return GMD_ERR_ARG; }
  // persistent buffers
  if ((r = new_state(m, &m->cur, nullptr)) || (r = acquire(m, KIND_G, &m->ghs)) || (r = new_tend(m, &m->tendOld)) ||
      (r = new_tend(m, &m->tendNew)) || ((cfg->nranks > 1 || m->fused) && (r = new_tend(m, &m->tendNew2)))) {
    gmd_destroy(m);
    return r;
  }
  CKD(cudaStreamSynchronize(m->stream));
#undef CKD
  *out = m;
  return 0;
}

int gmd_comm_unique_id(void *id128) {
  int r = nccl_load();
  if (r) return r;
  NK(g_nccl.GetUniqueId(id128));
  return 0;
}

int gmd_comm_init(gmd_model *m, const void *id128) {
  if (!m || !id128) return fail(GMD_ERR_ARG, "null argument");
  if (m->cfg.nranks == 1) return 0;
  int r = nccl_load();
  if (r) return r;
  if ((r = set_dev(m))) return r;
  NcclId128 id;
  memcpy(id.b, id128, 128);
  NK(g_nccl.CommInitRank(&m->comm, m->cfg.nranks, id, m->cfg.rank));
  return 0;
}

struct PeerBlob {  // what gmd_peer_export hands to the host program (<= GMD_PEER_BLOB_BYTES)
  unsigned magic;
  int rank, nranks, r0, r1, dev, cap, pad;
  long long pid;
  unsigned long long fld_elems;
  void *slab, *page;
  cudaIpcMemHandle_t hslab, hpage;
};
static_assert(sizeof(PeerBlob) <= GMD_PEER_BLOB_BYTES, "blob size");
static const unsigned PEER_MAGIC = 0x676d6431u;

int gmd_peer_export(gmd_model *m, void *blob) {
  if (!m || !blob) return fail(GMD_ERR_ARG, "null argument");
  int r = set_dev(m);
  if (r) return r;
  PeerBlob b;
  memset(&b, 0, sizeof b);
  b.magic = PEER_MAGIC;
  b.rank = m->cfg.rank;
  b.nranks = m->cfg.nranks;
  b.r0 = m->geo.r0;
  b.r1 = m->geo.r1;
  b.dev = m->dev;
  b.cap = m->slab_cap;
  b.pid = (long long)getpid();
  b.fld_elems = m->fld_elems;
  b.slab = m->slab;
  b.page = m->page;
  CK(cudaIpcGetMemHandle(&b.hslab, m->slab));
  CK(cudaIpcGetMemHandle(&b.hpage, m->page));
  memset(blob, 0, GMD_PEER_BLOB_BYTES);
  memcpy(blob, &b, sizeof b);
  return 0;
}

// map one peer allocation: CUDA IPC across processes, the pointer itself (with peer access) inside one process
static int peer_map(gmd_model *m, const PeerBlob &b, bool want_slab, void **out) {
  if (b.pid == (long long)getpid()) {
    if (b.dev != m->dev) {
      cudaError_t e = cudaDeviceEnablePeerAccess(b.dev, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) return fail(GMD_ERR_COMM, "no peer access from device %d to %d: %s", m->dev, b.dev, cudaGetErrorString(e));
    }
    *out = want_slab ? b.slab : b.page;
    return 0;
  }
  cudaError_t e = cudaIpcOpenMemHandle(out, want_slab ? b.hslab : b.hpage, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess)
    return fail(GMD_ERR_COMM, "cudaIpcOpenMemHandle of rank %d's %s failed: %s", b.rank, want_slab ? "field slab" : "signal page",
                cudaGetErrorString(e));
  m->ipc_opened.push_back(*out);
  return 0;
}

int gmd_peer_connect(gmd_model *m, const void *blobs, int nblobs) {
  if (!m || !blobs) return fail(GMD_ERR_ARG, "null argument");
  const int np = m->cfg.nranks, rank = m->cfg.rank;
  if (np == 1) return 0;
  if (nblobs != np) return fail(GMD_ERR_ARG, "%d blobs for %d ranks", nblobs, np);
  if (np > MAXR) return fail(GMD_ERR_ARG, "the peer-memory path serves up to %d ranks (one node); use gmd_comm_init", MAXR);
  if (m->p2p) return fail(GMD_ERR_STATE, "gmd_peer_connect called twice");
  // No kernel reads a ghost row while a neighbour may be writing it: every wait for a neighbour (the inner-product
  // all-reduce of predict_correct, k_halo_wait for the generic exchanges) completes in a launch BEFORE the one that
  // reads the rows, so the read-only (ld.global.nc) path of the stage kernel only ever sees rows that are immutable
  // for its lifetime.
  int r = set_dev(m);
  if (r) return r;
  if ((r = join(m))) return r;
  CK(cudaStreamSynchronize(m->stream));
  std::vector<PeerBlob> B((size_t)np);
  for (int p = 0; p < np; p++) {
    memcpy(&B[(size_t)p], (const char *)blobs + (size_t)p * GMD_PEER_BLOB_BYTES, sizeof(PeerBlob));
    const PeerBlob &b = B[(size_t)p];
    if (b.magic != PEER_MAGIC || b.rank != p || b.nranks != np) return fail(GMD_ERR_ARG, "blob %d is not rank %d's export", p, p);
    if (b.cap != m->slab_cap) return fail(GMD_ERR_ARG, "rank %d has a field pool of %d slots, this rank %d", p, b.cap, m->slab_cap);
  }
  if (rank > 0 && B[(size_t)rank - 1].r1 != m->geo.r0) return fail(GMD_ERR_ARG, "rank %d's band does not end where this one starts", rank - 1);
  if (rank + 1 < np && B[(size_t)rank + 1].r0 != m->geo.r1) return fail(GMD_ERR_ARG, "rank %d's band does not start where this one ends", rank + 1);
  for (int p = 0; p < np; p++) {
    if (p == rank) {
      m->peer_page[p] = m->page;
      continue;
    }
    void *q = nullptr;
    if ((r = peer_map(m, B[(size_t)p], false, &q))) return r;
    m->peer_page[p] = (u64 *)q;
    if (p == rank - 1 || p == rank + 1) {
      const int side = (p == rank - 1) ? 0 : 1;
      if ((r = peer_map(m, B[(size_t)p], true, &q))) return r;
      m->peer_slab[side] = (double *)q;
      m->peer_r0[side] = B[(size_t)p].r0;
      m->peer_fld[side] = (size_t)B[(size_t)p].fld_elems;
    }
  }
  {
    u64 tab[MAXR] = {};
    for (int p = 0; p < np; p++) tab[p] = (u64)(uintptr_t)m->peer_page[p];
    CK(cudaMemcpy(m->page + SP_PEERTAB, tab, sizeof tab, cudaMemcpyHostToDevice));
  }
  m->p2p = true;
  m->xk = m->rk = m->xwaited = 0;
  // wide-halo predict_correct: the S3a launches store the band-edge rows of the new tendency themselves
  m->fuse_push = m->wide && !getenv("GMD_NO_FUSED_PUSH");
  for (auto &g : m->graphs) cudaGraphExecDestroy(g.exec);
  m->graphs.clear();
  return 0;
}

int gmd_peer_disconnect(gmd_model *m) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  int r = set_dev(m);
  if (r) return r;
  if ((r = join(m))) return r;
  CK(cudaStreamSynchronize(m->stream));
  for (void *p : m->ipc_opened) cudaIpcCloseMemHandle(p);
  m->ipc_opened.clear();
  for (int p = 0; p < MAXR; p++) m->peer_page[p] = nullptr;
  if (m->cfg.rank < MAXR) m->peer_page[m->cfg.rank] = m->page;
  m->peer_slab[0] = m->peer_slab[1] = nullptr;
  m->p2p = false;
  m->fuse_push = false;
  m->xk = m->rk = m->xwaited = 0;
  CK(cudaMemset(m->page, 0, SP_WORDS * sizeof(u64)));
  for (auto &g : m->graphs) cudaGraphExecDestroy(g.exec);
  m->graphs.clear();
  return 0;
}

int gmd_get_band(const gmd_model *m, int *b, int *e) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  if (b) *b = m->geo.r0;
  if (e) *e = m->geo.r1;
  return 0;
}

int gmd_get_fused_rows(const gmd_model *m, int *b, int *e) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  if (b) *b = m->fused ? m->fz_I0 : 0;
  if (e) *e = m->fused ? m->fz_I1 : 0;
  return 0;
}

int gmd_set_stream(gmd_model *m, void *s) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  int r = set_dev(m);
  if (r) return r;
  if ((r = join(m))) return r;
  CK(cudaStreamSynchronize(m->stream));
  m->stream = s ? (cudaStream_t)s : m->own_stream;
  for (auto &g : m->graphs) cudaGraphExecDestroy(g.exec);
  m->graphs.clear();
  return 0;
}

int gmd_set_graph_mode(gmd_model *m, int on) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  m->graph_mode = on != 0;
  return 0;
}

static int diag_now(gmd_model *m) { return diag(m, m->cur, 0); }

// fetch the ring slots of the last n steps into h (full ring layout); one slot costs one 24-byte copy
static int ring_fetch(gmd_model *m, int n, std::vector<double> &h) {
  h.resize((size_t)3 * gmd_model::RING);
  if (n <= 1) {
    const int slot = m->step % gmd_model::RING;
    CK(cudaMemcpyAsync(&h[3 * (size_t)slot], m->d_ring + 3 * (size_t)slot, 3 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  } else {
    CK(cudaMemcpyAsync(h.data(), m->d_ring, h.size() * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  }
  CK(cudaStreamSynchronize(m->stream));
  return 0;
}

static int check_nan_last(gmd_model *m, int n) {
  // diag_run NaN abort, src/diag_mod.F90:79-87
  n = std::min(n, (int)gmd_model::RING);
  std::vector<double> h;
  if (int r = ring_fetch(m, n, h)) return r;
  for (int k = 0; k < n; k++) {
    const int slot = (m->step - k) % gmd_model::RING;
    if (slot < 0) break;
    if (std::isnan(h[3 * (size_t)slot])) return fail(GMD_ERR_NAN, "Total mass is NaN!");
    if (std::isnan(h[3 * (size_t)slot + 1])) return fail(GMD_ERR_NAN, "Total energy is NaN!");
  }
  return 0;
}

int gmd_set_state(gmd_model *m, const double *u, const double *v, const double *gd, const double *ghs, int layout) {
  if (!m || !u || !v || !gd) return fail(GMD_ERR_ARG, "null argument");
  if (layout != GMD_LAYOUT_COMPACT && layout != GMD_LAYOUT_REFERENCE) return fail(GMD_ERR_ARG, "bad layout %d", layout);
  int r = set_dev(m);
  if (r) return r;
  const int nlon = m->geo.nlon, nlat = m->geo.nlat;
  // pole rows of u must be zero (du is never written there, so U(pole) = 0 is an invariant of every reference IC).
  // The Shamir-Paldor wave IC evaluates cos(+-pi/2)^(sigma - 1.5) there, ~1e-146 m/s: values below GMD_POLE_U_TINY
  // are taken as the zero they stand for (they cannot change any other binary64 result of the step).
  for (int pj = 0; pj < 2; pj++) {
    const int j = pj ? nlat - 1 : 0;
    const double *row = (layout == GMD_LAYOUT_REFERENCE) ? u + (size_t)(j + 2) * (nlon + 4) + 2 : u + (size_t)j * nlon;
    for (int i = 0; i < nlon; i++)
      if (!(std::fabs(row[i]) <= GMD_POLE_U_TINY))
        return fail(GMD_ERR_ARG, "u must be 0 on the pole rows (row %d, column %d is %g)", j, i, row[i]);
  }
  if ((r = ensure_uv(m))) return r;
  {
    XferJob jobs[4] = {up_job(u, layout, nlat, m->w_u), up_job(v, layout, nlat - 1, m->w_v),
                       up_job(gd, layout, nlat, m->cur.gd), up_job(ghs, layout, nlat, m->ghs)};
    if ((r = run_xfers(m, jobs, 4))) return r;
  }
  for (int pj = 0; pj < 2; pj++) {
    const int j = pj ? nlat - 1 : 0;
    if (j >= m->geo.r0 - GHOST && j < m->geo.r1 + GHOST)
      CK(cudaMemsetAsync(m->w_u + (ptrdiff_t)(j - m->geo.r0) * nlon, 0, (size_t)nlon * sizeof(double), m->stream));
  }
  // iap_transform on owned + ghost rows inside the globe (src/types_mod.F90:399-426)
  Geo g = m->geo;
  g.r0 = std::max(m->geo.r0 - GHOST, 0);
  g.r1 = std::min(m->geo.r1 + GHOST - 1, nlat);   // V of a row needs gd of the next one
  const ptrdiff_t sh = (ptrdiff_t)(g.r0 - m->geo.r0) * nlon;
  if (!m->dry) k_iap<<<m->ew_blocks, 256, 0, m->stream>>>(g, m->w_u + sh, m->w_v + sh, m->cur.gd + sh, m->cur.U + sh, m->cur.V + sh);
  if ((r = post_launch(m))) return r;
  m->have_state = true;
  if (m->run_inited) {
    if ((r = diag_now(m))) return r;
    return check_nan_last(m, 1);
  }
  CK(cudaStreamSynchronize(m->stream));
  return 0;
}

int gmd_run_init(gmd_model *m) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  if (!m->have_state) return fail(GMD_ERR_STATE, "gmd_set_state has not been called");
  int r = set_dev(m);
  if (r) return r;
  m->run_inited = true;
  if ((r = diag_now(m))) return r;
  return check_nan_last(m, 1);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// step execution: direct launches for the first steps (warms the buffer pool), then one CUDA graph per
// distinct `cur` buffer triple.  A replay advances the host-side pool bookkeeping with a dry pass of the
// same step logic (no launches), so direct and replayed steps can be mixed freely.
// ---------------------------------------------------------------------------------------------------------
static int enqueue_steps(gmd_model *m, int nsteps) {
  int r;
  for (int n = 0; n < nsteps; n++) {
    // NCCL send/recv/all-reduce nodes are capturable; keep multi-rank capture opt-out via GMD_GRAPH_MULTI=0
    static const bool graph_multi = !(getenv("GMD_GRAPH_MULTI") && atoi(getenv("GMD_GRAPH_MULTI")) == 0);
    const bool use_graph = m->graph_mode && (m->cfg.nranks == 1 || graph_multi) && m->step >= 2;
    if (!use_graph) {
      if ((r = one_step(m))) return r;
      continue;
    }
    // the buffers a step uses follow from `cur`, the tendency parity and the free lists (public entry points other than
    // gmd_step may have acquired / released buffers since the graph was captured): all of it is the key
    for (int q = 0; q < 3; q++) std::sort(m->free_[q].begin(), m->free_[q].end());
    std::vector<double *> key = {m->cur.U, m->cur.V, m->cur.gd, (double *)(uintptr_t)m->tn_idx, (double *)(uintptr_t)m->slab_next};
    for (int q = 0; q < 3; q++) {
      key.push_back(nullptr);
      key.insert(key.end(), m->free_[q].begin(), m->free_[q].end());
    }
    gmd_model::GraphEntry *hit = nullptr;
    for (auto &g : m->graphs)
      if (g.key == key) { hit = &g; break; }
    if (hit) {
      CK(cudaGraphLaunch(hit->exec, m->stream));
      m->dry = true;
      r = one_step(m);
      m->dry = false;
      if (r) return r;
      continue;
    }
    if (m->graphs.size() >= 16) {  // unexpected: pool state does not cycle; stay correct, lose the graph
      if ((r = one_step(m))) return r;
      continue;
    }
    CK(cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal));
    m->capturing = true;
    const long long l0 = m->launches;
    r = one_step(m);
    m->capturing = false;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(m->stream, &graph);
    if (r) {
      if (graph) cudaGraphDestroy(graph);
      return r;
    }
    if (ce != cudaSuccess) return fail(GMD_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    gmd_model::GraphEntry ge;
    ge.key = key;
    ge.launches = m->launches - l0;
    CK(cudaGraphInstantiate(&ge.exec, graph, 0));
    cudaGraphDestroy(graph);
    m->graphs.push_back(ge);
    CK(cudaGraphLaunch(ge.exec, m->stream));  // capture executed nothing
  }
  return 0;
}

extern "C" {

int gmd_step_async(gmd_model *m, int nsteps) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  if (!m->run_inited) return fail(GMD_ERR_STATE, "gmd_run_init has not been called");
  if (nsteps < 0) return fail(GMD_ERR_ARG, "nsteps < 0");
  int r = set_dev(m);
  if (r) return r;
  if (!m->span_open) {
    CK(cudaEventRecord(m->ev0, m->stream));
    m->span_open = true;
    m->pending_steps = 0;
  }
  if ((r = enqueue_steps(m, nsteps))) return r;
  m->pending_steps += nsteps;
  return 0;
}

int gmd_sync(gmd_model *m) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  int r = set_dev(m);
  if (r) return r;
  if ((r = join(m))) return r;
  if (m->span_open) {
    CK(cudaEventRecord(m->ev1, m->stream));
    CK(cudaEventSynchronize(m->ev1));
    CK(cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1));
    m->span_open = false;
  } else {
    CK(cudaStreamSynchronize(m->stream));
  }
  const int n = m->pending_steps;
  m->pending_steps = 0;
  if (m->p2p) {
    u64 err = 0;
    CK(cudaMemcpy(&err, m->page + SP_ERR, sizeof err, cudaMemcpyDeviceToHost));
    if (err) return fail(GMD_ERR_COMM, "a wait for a neighbour rank's halo rows or partial sums timed out");
  }
  return check_nan_last(m, std::max(n, 1));
}

int gmd_step(gmd_model *m, int nsteps) {
  int r = gmd_step_async(m, nsteps);
  if (r) return r;
  return gmd_sync(m);
}

int gmd_last_step_ms(gmd_model *m, float *ms) {
  if (!m || !ms) return fail(GMD_ERR_ARG, "null argument");
  *ms = m->last_ms;
  return 0;
}

long long gmd_kernel_launches(const gmd_model *m) { return m ? m->launches : 0; }
int gmd_get_step_count(const gmd_model *m) { return m ? m->step : 0; }

double gmd_algorithmic_bytes_per_column_step(const gmd_model *m) {
  if (!m) return 0.0;
  double b;
  const int S = m->cfg.subcycles;
  switch (m->cfg.split_scheme) {
    case GMD_SPLIT_CSP2: b = 312.0 * S + 216.0 * 2; break;
    case GMD_SPLIT_ISP: b = 312.0 * S + 216.0 * 2; break;  // same operator-evaluation count as csp2 (3S fast + 6 slow)
    default: b = 312.0;
  }
  if (m->cfg.use_diffusion) b += 48.0 * (m->cfg.diffusion_order / 2);
  return b;
}

int gmd_get_diag_series(gmd_model *m, int n, double *mass, double *energy, double *beta) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  if (n < 0 || n > gmd_model::RING || n > m->step + 1) return fail(GMD_ERR_ARG, "bad series length %d", n);
  int r = set_dev(m);
  if (r) return r;
  std::vector<double> h;
  if ((r = ring_fetch(m, n, h))) return r;
  for (int k = 0; k < n; k++) {
    const int slot = (m->step - (n - 1 - k)) % gmd_model::RING;
    if (mass) mass[k] = h[3 * (size_t)slot];
    if (energy) energy[k] = h[3 * (size_t)slot + 1];
    if (beta) beta[k] = h[3 * (size_t)slot + 2];
  }
  return 0;
}

int gmd_get_diag(gmd_model *m, double *mass, double *energy, double *beta) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  if (!m->run_inited) return fail(GMD_ERR_STATE, "gmd_run_init has not been called");
  return gmd_get_diag_series(m, 1, mass, energy, beta);
}

int gmd_get_state(gmd_model *m, double *u, double *v, double *gd, int layout) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  if (!m->have_state) return fail(GMD_ERR_STATE, "gmd_set_state has not been called");
  int r = set_dev(m);
  if (r) return r;
  if ((r = derive_uv(m, m->cur))) return r;
  const int nlat = m->geo.nlat;
  double *const dev[3] = {m->w_u, m->w_v, m->cur.gd};
  double *const host[3] = {u, v, gd};
  const int rows[3] = {nlat, nlat - 1, nlat};
  const bool zp[3] = {false, false, false};
  return download_fields(m, 3, dev, host, rows, zp, layout);
}

int gmd_get_iap_state(gmd_model *m, double *iu, double *iv, double *igd, int layout) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  if (!m->have_state) return fail(GMD_ERR_STATE, "gmd_set_state has not been called");
  int r = set_dev(m);
  if (r) return r;
  const int nlat = m->geo.nlat;
  double *tmp = nullptr;
  if (igd) {
    if ((r = acquire(m, KIND_G, &tmp))) return r;
    if ((r = join(m))) return r;
    k_derive<<<m->ew_blocks, 256, 0, m->stream>>>(m->geo, m->geo.r0, m->geo.r1, m->cur.U, m->cur.V, m->cur.gd, nullptr,
                                                 nullptr, tmp);
    if ((r = post_launch(m))) return r;
  }
  double *const dev[3] = {m->cur.U, m->cur.V, tmp};
  double *const host[3] = {iu, iv, igd};
  const int rows[3] = {nlat, nlat - 1, nlat};
  const bool zp[3] = {false, false, false};
  r = download_fields(m, 3, dev, host, rows, zp, layout);
  if (tmp) release(m, tmp);
  return r;
}

int gmd_get_vor_div(gmd_model *m, double *vor, double *div, int layout) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  if (!m->run_inited) return fail(GMD_ERR_STATE, "gmd_run_init has not been called");
  int r = set_dev(m);
  if (r) return r;
  if ((r = derive_uv(m, m->cur))) return r;
  double *dv = nullptr, *dd = nullptr;
  if ((r = acquire(m, KIND_V, &dv))) return r;
  if ((r = acquire(m, KIND_G, &dd))) return r;
  k_vor_div<<<m->ew_blocks, 256, 0, m->stream>>>(m->geo, m->tab, m->w_u, m->w_v, dv, dd);
  if ((r = post_launch(m))) return r;
  const int nlat = m->geo.nlat;
  {
    double *const dev[2] = {dv, dd};
    double *const host[2] = {vor, div};
    const int rows[2] = {nlat - 1, nlat};
    const bool zp[2] = {false, false};
    r = download_fields(m, 2, dev, host, rows, zp, layout);
  }
  release(m, dv);
  release(m, dd);
  return r;
}

int gmd_space_operators(gmd_model *m, int pass, double *du, double *dv, double *dgd, int layout) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  if (!m->run_inited) return fail(GMD_ERR_STATE, "gmd_run_init has not been called");
  if (pass < 0 || pass > 2) return fail(GMD_ERR_ARG, "bad pass %d", pass);
  int r = set_dev(m);
  if (r) return r;
  const size_t bytes = (size_t)m->nr * m->geo.nlon * sizeof(double);
  if ((r = join(m))) return r;
  if (pass == PASS_SLOW) CK(cudaMemsetAsync(m->tendNew.gd, 0, bytes, m->stream));  // dgd = 0, :297
  if ((r = stage(m, pass, MODE_EVAL, m->cur, nullptr, 0.0, nullptr, &m->tendNew, nullptr))) return r;
  const int nlat = m->geo.nlat;
  double *const dev[3] = {m->tendNew.U, m->tendNew.V, m->tendNew.gd};
  double *const host[3] = {du, dv, dgd};
  const int rows[3] = {nlat, nlat - 1, nlat};
  const bool zp[3] = {true, false, false};
  return download_fields(m, 3, dev, host, rows, zp, layout);
}

int gmd_predict_correct(gmd_model *m, double dt, int pass) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  if (!m->run_inited) return fail(GMD_ERR_STATE, "gmd_run_init has not been called");
  if (pass < 0 || pass > 2) return fail(GMD_ERR_ARG, "bad pass %d", pass);
  int r = set_dev(m);
  if (r) return r;
  Carry ci, co;
  ci.base = m->cur;
  if ((r = predict_correct(m, dt, ci, pass, false, &co))) return r;
  const State out = co.base;
  release_state(m, &m->cur);
  m->cur = out;
  if ((r = join(m))) return r;
  if ((r = unit_end(m))) return r;
  CK(cudaStreamSynchronize(m->stream));
  return 0;
}

int gmd_ordinary_diffusion(gmd_model *m, double dt) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  if (!m->run_inited) return fail(GMD_ERR_STATE, "gmd_run_init has not been called");
  int r = set_dev(m);
  if (r) return r;
  State out;
  if ((r = diffusion(m, dt, m->cur, &out))) return r;
  release_state(m, &m->cur);
  m->cur = out;
  if ((r = join(m))) return r;
  if ((r = unit_end(m))) return r;
  CK(cudaStreamSynchronize(m->stream));
  return 0;
}

int gmd_filter_row(gmd_model *m, int half, int row0, double *x) {
  if (!m || !x) return fail(GMD_ERR_ARG, "null argument");
  const int nlat = m->geo.nlat, nlon = m->geo.nlon;
  if (row0 < 0 || row0 >= (half ? nlat - 1 : nlat)) return fail(GMD_ERR_ARG, "row %d out of range", row0);
  int r = set_dev(m);
  if (r) return r;
  double *buf = nullptr;
  if ((r = acquire(m, KIND_G, &buf))) return r;
  const int cutoff = half ? m->mesh.cut_half[(size_t)row0] : m->mesh.cut_full[(size_t)row0];
  CK(cudaMemcpyAsync(buf, x, (size_t)nlon * sizeof(double), cudaMemcpyHostToDevice, m->stream));
  PolarArgs p;
  memset(&p, 0, sizeof p);
  p.g = m->geo;
  p.t = m->tab;
  p.items[0] = pack_item(IT_DGD, m->geo.r0, cutoff);
  p.Tgd = buf;
  p.rescale = 0;
  p.partials = m->d_partials;
  if (2 * (cutoff + 1) > m->ncoef_max && cutoff >= 0) {
    release(m, buf);
    return fail(GMD_ERR_STATE, "internal: basis too small");
  }
  launch_polar(m, MODE_EVAL, 1, p, m->stream);
  if ((r = post_launch(m))) return r;
  CK(cudaMemcpyAsync(x, buf, (size_t)nlon * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  release(m, buf);
  return 0;
}

int gmd_get_filter_rows(const gmd_model *m, int *ff, int *fc, int *hf, int *hc) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
  const int nlat = m->geo.nlat;
  for (int j = 0; j < nlat; j++) {
    if (ff) ff[j] = m->mesh.flag_full[(size_t)j];
    if (fc) fc[j] = m->mesh.cut_full[(size_t)j];
  }
  for (int j = 0; j < nlat - 1; j++) {
    if (hf) hf[j] = m->mesh.flag_half[(size_t)j];
    if (hc) hc[j] = m->mesh.cut_half[(size_t)j];
  }
  return 0;
}

int gmd_get_table(const gmd_model *m, int which, double *out) {
  if (!m || !out) return fail(GMD_ERR_ARG, "null argument");
  const HostMesh &M = m->mesh;
  const std::vector<double> *t;
  int n = M.nlat;
  switch (which) {
    case 0: t = &M.full_cos; break;
    case 1: t = &M.half_cos; n--; break;
    case 2: t = &M.full_f; break;
    case 3: t = &M.full_c; break;
    case 4: t = &M.full_dlon; break;
    case 5: t = &M.half_dlon; n--; break;
    case 6: t = &M.full_dlat; break;
    case 7: t = &M.half_dlat; n--; break;
    case 8: t = &M.full_lat; break;
    case 9: t = &M.half_lat; n--; break;
    default: return fail(GMD_ERR_ARG, "bad table %d", which);
  }
  for (int j = 0; j < n; j++) out[j] = (*t)[(size_t)(j + TPAD)];
  return 0;
}

}  // extern "C"

extern "C" {

int gmd_trace_begin(gmd_model *m) {
  if (!m) return fail(GMD_ERR_ARG, "null model");
#if GMD_TRACE
  int r = set_dev(m);
  if (r) return r;
  if ((r = join(m))) return r;
  CK(cudaStreamSynchronize(m->stream));
  const size_t n = (size_t)gmd_model::TRACE_STEPS * gmd_model::TRACE_PER_STEP;
  if (!m->d_trace) CK(cudaMalloc(&m->d_trace, n * 3 * sizeof(u64)));
  std::vector<u64> h(n * 3, 0);
  for (size_t k = 0; k < n; k++) h[3 * k] = ~0ull;
  CK(cudaMemcpy(m->d_trace, h.data(), h.size() * sizeof(u64), cudaMemcpyHostToDevice));
  TraceBuf tb = {m->d_trace, m->d_ctr, gmd_model::TRACE_STEPS, gmd_model::TRACE_PER_STEP};
  CK(cudaMemcpyToSymbol(g_trace, &tb, sizeof tb));
  m->trace_step0 = m->step;
  return 0;
#else
  return fail(GMD_ERR_STATE, "this build has no device timeline (load libgmd_trace.so)");
#endif
}

int gmd_trace_dump(gmd_model *m, const char *path) {
  if (!m || !path) return fail(GMD_ERR_ARG, "null argument");
#if GMD_TRACE
  if (!m->d_trace) return fail(GMD_ERR_STATE, "gmd_trace_begin has not been called");
  int r = set_dev(m);
  if (r) return r;
  if ((r = join(m))) return r;
  CK(cudaStreamSynchronize(m->stream));
  const size_t n = (size_t)gmd_model::TRACE_STEPS * gmd_model::TRACE_PER_STEP;
  std::vector<u64> h(n * 3);
  CK(cudaMemcpy(h.data(), m->d_trace, h.size() * sizeof(u64), cudaMemcpyDeviceToHost));
  FILE *f = fopen(path, "w");
  if (!f) return fail(GMD_ERR_ARG, "cannot write %s", path);
  fprintf(f, "{\"rank\": %d, \"nranks\": %d, \"band\": [%d, %d], \"steps\": [", m->cfg.rank, m->cfg.nranks, m->geo.r0, m->geo.r1);
  const int first = std::max(m->trace_step0, m->step - gmd_model::TRACE_STEPS);
  for (int st = first; st < m->step; st++) {
    fprintf(f, "%s{\"step\": %d, \"launches\": [", st == first ? "" : ", ", st);
    const u64 *rec = h.data() + (size_t)(st % gmd_model::TRACE_STEPS) * gmd_model::TRACE_PER_STEP * 3;
    bool any = false;
    for (int q = 0; q < gmd_model::TRACE_PER_STEP; q++) {
      if (rec[3 * q] == ~0ull) continue;
      const char *nm = (size_t)q < m->trace_names.size() ? m->trace_names[(size_t)q].c_str() : "?";
      fprintf(f, "%s{\"seq\": %d, \"kernel\": \"%s\", \"start_ns\": %llu, \"end_ns\": %llu, \"wait_ns\": %llu}", any ? ", " : "", q, nm,
              rec[3 * q], rec[3 * q + 1], rec[3 * q + 2]);
      any = true;
    }
    fprintf(f, "]}");
  }
  fprintf(f, "]");
  {   // phase stamps of CTA 0 of the most recent fused polar-cap launches
    std::vector<u64> cd(64 * 12);
    CK(cudaMemcpyFromSymbol(cd.data(), g_capdbg, cd.size() * sizeof(u64)));
    fprintf(f, ", \"cap_phases_ns\": [");
    bool any = false;
    for (int q = 0; q < 64; q++) {
      const u64 *c = cd.data() + 12 * q;
      if (!c[0] || !c[4]) continue;
      // sweep, grid barrier, first item, all items; inside the first item: loads + analysis, first cluster sum,
      // synthesis, second cluster sum, stores
      fprintf(f, "%s[%llu, %llu, %llu, %llu, %llu, %llu, %llu, %llu, %llu]", any ? ", " : "", c[1] - c[0], c[2] - c[1],
              c[3] > c[2] ? c[3] - c[2] : 0ull, c[4] - c[2], c[5] > c[2] ? c[5] - c[2] : 0ull, c[6] > c[5] ? c[6] - c[5] : 0ull,
              c[7] > c[6] ? c[7] - c[6] : 0ull, c[8] > c[7] ? c[8] - c[7] : 0ull, c[9] > c[8] ? c[9] - c[8] : 0ull);
      any = true;
    }
    fprintf(f, "]");
  }
  fprintf(f, "}\n");
  fclose(f);
  TraceBuf tb = {nullptr, nullptr, 0, 0};
  CK(cudaMemcpyToSymbol(g_trace, &tb, sizeof tb));
  return 0;
#else
  return fail(GMD_ERR_STATE, "this build has no device timeline (load libgmd_trace.so)");
#endif
}

}  // extern "C"

// time `reps` back-to-back launches of one stage-kernel variant (no polar rows), CUDA events on the stream
static int time_stage(gmd_model *m, int pass, int mode, int reps, float *ms_per_launch) {
  int r;
  State A, B;
  const bool slow = (pass == PASS_SLOW);
  if ((r = new_state(m, &A, slow ? m->cur.gd : nullptr))) return r;
  if ((r = new_state(m, &B, slow ? m->cur.gd : nullptr))) return r;
  const double dt = 0.5 * m->cfg.time_step_size / std::max(1, m->cfg.subcycles);
  // warm-up / valid operands: A = cur + dt L(cur), tendOld = L(A), tendNew = L(B)
  if ((r = stage(m, pass, MODE_S1, m->cur, &m->cur, dt, &A, &m->tendOld, nullptr))) return r;
  if ((r = stage(m, pass, MODE_S2, A, &m->cur, dt, &B, &m->tendOld, nullptr))) return r;
  if ((r = stage(m, pass, MODE_S3A, B, nullptr, 0.0, nullptr, &m->tendNew, &m->tendOld))) return r;
  StageArgs a;
  memset(&a, 0, sizeof a);
  a.g = m->geo; a.t = m->tab;
  a.EU = A.U; a.EV = A.V; a.Egd = A.gd; a.ghs = m->ghs;
  a.OU = m->cur.U; a.OV = m->cur.V; a.Ogd = m->cur.gd;
  a.NU = B.U; a.NV = B.V; a.Ngd = B.gd;
  if (mode == MODE_S3A) {
    a.TU = m->tendNew.U; a.TV = m->tendNew.V; a.Tgd = m->tendNew.gd;
    a.PU = m->tendOld.U; a.PV = m->tendOld.V; a.Pgd = m->tendOld.gd;
  } else {
    a.TU = m->tendOld.U; a.TV = m->tendOld.V; a.Tgd = m->tendOld.gd;
  }
  a.dt = dt;
  a.beta_lon = m->cfg.uv_adv_upwind_lon_beta; a.beta_lat = m->cfg.uv_adv_upwind_lat_beta;
  a.AUlon = m->w_alon_u; a.AUlat = m->w_alat_u; a.AVlon = m->w_alon_v; a.AVlat = m->w_alat_v;
  a.partials = m->d_partials;
  a.flmask = 0x0fu;
  const int rpc = (mode == MODE_S3A) ? m->rows_per_cta_s3a : m->rows_per_cta;
  a.rows_per_cta = rpc;
  a.rb[0] = m->geo.r0; a.re[0] = m->geo.r1; a.pofs[0] = 0;
  if ((r = join(m))) return r;
  dim3 grid((unsigned)m->nbx, (unsigned)((m->nr + rpc - 1) / rpc), 1);
  stage_fn fn = pick_stage(pass, m->cfg.uv_adv_scheme, mode == 5 ? MODE_S1 : mode);
  if (mode == 4) {  // MODE_S1 with the deferred update folded in: E = cur + beta dt tendNew -> A (= M), N = B
    a.LU = m->tendNew.U; a.LV = m->tendNew.V; a.Lgd = m->tendNew.gd;
    a.MU = A.U; a.MV = A.V; a.Mgd = A.gd;
    a.EU = m->cur.U; a.EV = m->cur.V; a.Egd = m->cur.gd;
    a.lip = m->d_ip; a.ldt = dt; a.lqcon = m->cfg.qcon_modified;
    fn = pick_stage_lazy(pass, m->cfg.uv_adv_scheme == ADV_WENO ? ADV_CENTER : m->cfg.uv_adv_scheme, slow ? 2 : 1);
  }
  unsigned threads = BX;
  size_t smem = stage_smem_bytes(rpc);
  if (mode == 5) {  // the fused predict_correct kernel with the deferred update, over the rows it covers in a step
    if (!m->fused) return fail(GMD_ERR_STATE, "the fused predict_correct kernel is not in use in this configuration");
    a.LU = m->tendNew.U; a.LV = m->tendNew.V; a.Lgd = m->tendNew.gd;
    a.MU = A.U; a.MV = A.V; a.Mgd = A.gd;
    a.EU = m->cur.U; a.EV = m->cur.V; a.Egd = m->cur.gd;
    a.TU = m->tendNew2.U; a.TV = m->tendNew2.V; a.Tgd = m->tendNew2.gd;
    a.lip = m->d_ip; a.ldt = dt; a.lqcon = m->cfg.qcon_modified;
    a.rows_per_cta = m->fz_rpc;
    a.rb[0] = m->fz_I0; a.re[0] = m->fz_I1;
    fn = pick_pc(pass, m->cfg.uv_adv_scheme, 1, false);
    if (!fn) return fail(GMD_ERR_STATE, "no fused kernel for this configuration");
    grid = dim3((unsigned)m->fz_nstrips, (unsigned)m->fz_nchunks, 1);
    threads = PC_BX;
    smem = PC_SMEM_BYTES;
  }
  if ((r = allow_smem((const void *)fn))) return r;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, m->stream));
  for (int k = 0; k < reps; k++) {
    fn<<<grid, threads, smem, m->stream>>>(a);
    m->launches++;
  }
  CK(cudaEventRecord(e1, m->stream));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  CK(cudaGetLastError());
  *ms_per_launch = ms / reps;
  if ((r = unit_end(m))) return r;
  release_state(m, &A);
  release_state(m, &B);
  return 0;
}

extern "C" {

int gmd_time_stage_kernel(gmd_model *m, int reps, float *ms_per_launch, double *alg_bytes) {
  if (!m || !ms_per_launch) return fail(GMD_ERR_ARG, "null argument");
  if (!m->run_inited) return fail(GMD_ERR_STATE, "gmd_run_init has not been called");
  if (reps < 1) return fail(GMD_ERR_ARG, "reps < 1");
  int r = set_dev(m);
  if (r) return r;
  const int pass = (m->cfg.split_scheme == GMD_SPLIT_CSP2 || m->cfg.split_scheme == GMD_SPLIT_ISP) ? PASS_FAST : PASS_ALL;
  if ((r = time_stage(m, pass, MODE_S2, reps, ms_per_launch))) return r;
  // S2 of the all/fast pass: reads U,V,gd,ghs of the evaluated state + U,V,gd of the base state, writes the
  // new U,V,gd and the three tendencies: 13 words per column
  if (alg_bytes) *alg_bytes = 13.0 * 8.0 * (double)m->nr * (double)m->geo.nlon;
  return 0;
}

int gmd_time_stage_variant(gmd_model *m, int pass, int mode, int reps, float *ms_per_launch, double *alg_bytes) {
  if (!m || !ms_per_launch) return fail(GMD_ERR_ARG, "null argument");
  if (!m->run_inited) return fail(GMD_ERR_STATE, "gmd_run_init has not been called");
  if (reps < 1 || pass < 0 || pass > 2 || mode < 0 || mode > 5) return fail(GMD_ERR_ARG, "bad argument");
  int r = set_dev(m);
  if (r) return r;
  if ((r = time_stage(m, pass, mode, reps, ms_per_launch))) return r;
  // words per column: fast/all S1 7, S2 13, S3a 10; slow S1 5, S2 9, S3a 7; EVAL = reads + 3 (2) writes;
  // mode 5, the fused kernel: what it actually has to move -- old state 3 + ghs 1 + deferred tendency 3 in, materialised
  // state 3 + new tendency 3 out (slow pass: no ghs, gd carries no new tendency) -- over the rows it covers
  static const double words[2][6] = {{7, 13, 10, 7, 13, 13}, {5, 9, 7, 5, 9, 11}};
  const double rows = (mode == 5) ? (double)(m->fz_I1 - m->fz_I0) : (double)m->nr;
  if (alg_bytes) *alg_bytes = words[pass == PASS_SLOW ? 1 : 0][mode] * 8.0 * rows * (double)m->geo.nlon;
  return 0;
}

}  // extern "C"

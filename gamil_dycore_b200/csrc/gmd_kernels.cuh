// gmd_kernels.cuh -- sm_100a kernels of the barotropic shallow-water step (fp64, HBM-bound stencils).
//
// Persistent device state is the MINIMAL one: U = sqrt(gd)-weighted u, V likewise, gd (the IAP variables of
// src/types_mod.F90:25-46) per time level, plus ghs.  u, v and sqrt(gd) are recomputed on chip inside the
// fused stage kernel; the reference's 13 tendency arrays (src/types_mod.F90:58-72) never exist in HBM.
//
// Field addressing: a field pointer p addresses the rank's band, p[(j - r0) * nlon + i] for global row j
// and column i (0-based); rows r0-GHOST .. r1-1+GHOST are allocated.  Rows outside the globe are zeros
// that are never written (the reference's permanent zero latitude halos).
//
// GMD_STRICT=1 (libgmd_strict.so, built with -fmad=false) keeps every division and the exact operand
// order of the reference so that element-wise results are bit-identical to the oracle away from the
// reduction rows; GMD_STRICT=0 (libgmd.so, the product) multiplies by host-precomputed reciprocals and
// lets ptxas contract a*b+c into DFMA.
#pragma once
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include <cstdio>
#include <stdint.h>

#ifndef GMD_STRICT
#define GMD_STRICT 0
#endif

namespace gmd {

constexpr int GHOST = 6;      // ghost rows on each side of a band
// Wide halos (latitude bands): at the start of a predict_correct a band holds its state on
// HALO_S rows beyond its southern and HALO_N rows beyond its northern edge; the three operator sweeps of the
// predict_correct then run on rows that shrink by (1 south, 2 north) per sweep -- the stencil's reach -- so that the
// bands exchange rows ONCE per predict_correct (the new tendency) instead of once per sweep.
constexpr int HALO_S = 3, HALO_N = 6;
static_assert(HALO_N <= GHOST && HALO_S <= GHOST, "ghost rows");
typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------------------------
// Device timeline (libgmd_trace.so only, -DGMD_TRACE=1): every launch of a model step records
// {first CTA start, last CTA end, longest in-kernel wait for a neighbour} from %globaltimer into
// rec[((step % steps) * per_step + seq) * 3 ..]; `seq` is the launch's position inside the step, baked into its
// arguments (so a captured graph replays into the slots of the current step), the step number is the device step
// counter.  The product build compiles all of this away.
// ---------------------------------------------------------------------------------------------------------
#ifndef GMD_TRACE
#define GMD_TRACE 0
#endif
#if GMD_TRACE
struct TraceBuf {
  u64 *rec;
  const int *ctr;
  int steps, per_step;
};
__device__ TraceBuf g_trace;
__device__ __forceinline__ u64 gtimer() {
  u64 t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ u64 *trace_slot(int seq, int back) {
  seq -= 1;   // 0 = no slot (zero-initialised argument structs)
  if (seq < 0 || g_trace.rec == nullptr || seq >= g_trace.per_step) return nullptr;
  const int st = (*g_trace.ctr - back) % g_trace.steps;
  return g_trace.rec + ((size_t)(st < 0 ? st + g_trace.steps : st) * g_trace.per_step + seq) * 3;
}
__device__ __forceinline__ void trace_in(int seq, int back = 0) {
  if (threadIdx.x == 0) {
    u64 *r = trace_slot(seq, back);
    if (r) atomicMin(r, gtimer());
  }
}
__device__ __forceinline__ void trace_out(int seq, int back = 0) {
  if (threadIdx.x == 0) {
    u64 *r = trace_slot(seq, back);
    if (r) atomicMax(r + 1, gtimer());
  }
}
__device__ __forceinline__ void trace_wait(int seq, u64 t0) {
  u64 *r = trace_slot(seq, 0);
  if (r) atomicMax(r + 2, gtimer() - t0);
}
#define GMD_TRACE_T0() const u64 trace_t0_ = gtimer()
#define GMD_TRACE_WAIT(seq) trace_wait(seq, trace_t0_)
#else
__device__ __forceinline__ void trace_in(int, int = 0) {}
__device__ __forceinline__ void trace_out(int, int = 0) {}
#define GMD_TRACE_T0() do { } while (0)
#define GMD_TRACE_WAIT(seq) do { } while (0)
#endif

enum { PASS_ALL = 0, PASS_FAST = 1, PASS_SLOW = 2 };
enum { ADV_CENTER = 0, ADV_UPWIND = 1, ADV_WENO = 2 };
// what the stage kernel does with the tendency it has just computed
enum { MODE_S1 = 0,    // new = base + dt * tend               (tend stored only on filtered rows)
       MODE_S2 = 1,    // new = base + dt * tend, tend stored
       MODE_S3A = 2,   // tend stored, <tend,prev> and <tend,tend> accumulated
       MODE_EVAL = 3 };// tend stored

// row flags (global row index)
enum { FL_DU = 1, FL_DGD = 2, FL_POLE = 4, FL_DV = 8 };
// the same bits shifted left by FL_REDUCE_SHIFT mark the rows of the moving reduced tendency (specified extension,
// DESIGN.md section 8); StageArgs::flmask decides per launch whether they count (the slow pass only with reduce_adv_lon)
constexpr int FL_REDUCE_SHIFT = 4;

struct Tab {  // device pointers, already offset by TPAD: valid for j in [-TPAD, nlat+TPAD)
  const double *cosf, *cosh, *ff, *fc;
  const double *fdlon, *hdlon, *fdlat, *hdlat;
  const double *q_fdlon, *q_hdlon, *q_fdlat, *q_hdlat;  // 0.25 / d
  const double *cor1, *cor2;                            // cosh[j-1]/cosf[j], cosh[j]/cosf[j]
  const double *r_fdlon, *hc_hdlat, *h_fdlon, *h_fdlat; // 1/fdlon, cosh/hdlat, 0.5/fdlon, 0.5/fdlat
  const double *r_fdlon2, *r_hdlon2, *r_fdlat2, *r_hdlat2; // unused in strict mode (diffusion)
  const unsigned char *flags;                           // FL_* per row
  const double *rowrec;                                 // packed per-row record for k_stage, RC_N doubles per row
};

// packed row record (one 128-byte line per latitude row; copied to shared memory by k_stage)
enum { RC_COSF = 0, RC_COSH, RC_FF, RC_FC, RC_Q_FDLON, RC_Q_FDLAT, RC_Q_HDLON, RC_Q_HDLAT, RC_COR1, RC_COR2,
       RC_PGFU,   // product: 1/fdlon      strict: fdlon
       RC_PGFV,   // product: cosh/hdlat   strict: hdlat
       RC_MLON,   // product: 0.5/fdlon    strict: fdlon
       RC_MLAT,   // product: 0.5/fdlat    strict: fdlat
       RC_FLAGS, RC_PAD, RC_N };

struct Geo {
  int nlon, nlat;
  int r0, r1;  // this rank owns full rows [r0, r1)
};

// arguments of the peer all-reduce of one partial pair (reduce_pairs_body below)
struct RedArgs {
  u64 *page;          // my signal page (page[SP_PEERTAB + p] = rank p's page as mapped here)
  int rank, nranks;
  unsigned k;         // reduction epoch = page[SP_RBASE] + k
  int tseq;           // timeline slot + 1 of the launch (GMD_TRACE builds), 0 = none
};
// the inner-product finalisation folded into the launches that produce the partials (fold_tail below)
struct Fold {
  unsigned *ticket;   // null: the partials are reduced by a k_reduce_pairs launch instead
  unsigned total;     // CTAs of all launches that contribute
  int n;              // partial pairs
  double *out;
  RedArgs r;
};

struct StageArgs {
  Geo g;
  Tab t;
  const double *EU, *EV, *Egd, *ghs;  // state the operators are evaluated on
  const double *OU, *OV, *Ogd;        // base state of the update
  double *NU, *NV, *Ngd;              // new state
  double *TU, *TV, *Tgd;              // tendency out
  const double *PU, *PV, *Pgd;        // previous tendency for <tend, prev>
  const double *AUlon, *AUlat, *AVlon, *AVlat;  // WENO advection terms, precomputed
  double dt, beta_lon, beta_lat;
  double *partials;  // pairs {ip1, ip2}, one per CTA
  int rows_per_cta;
  // row ranges of this launch, selected by blockIdx.z (boundary launches cover two disjoint ranges)
  int rb[2], re[2];
  int pofs[2];       // first partial slot of each range
  // peer halo (the PUSH instantiations of MODE_S3A): rows j < push_s_end of the new tendency are also stored
  // into the south neighbour's ghost rows, rows j >= push_n_begin into the north neighbour's (pointers already shifted
  // to this rank's row coordinates, null = no neighbour).  The epoch that tells the neighbours the rows have landed
  // is the one the inner-product all-reduce of the same predict_correct releases (reduce_pairs_body): every CTA
  // fences its peer stores at system scope before it takes its ticket / before the kernel ends.
  double *hpS_U, *hpS_V, *hpS_G, *hpN_U, *hpN_V, *hpN_G;
  int push_s_end, push_n_begin;
  // deferred update (LAZY variants of MODE_S1): the evaluated state is E = (EU,EV,Egd) + beta ldt (LU,LV,Lgd), the
  // last update_state of the previous predict_correct (src/dycore_mod.F90:786-790) folded into this sweep; E is
  // also written out to (MU,MV,Mgd) on the rows of this CTA; the first CTA of a range with medge[z] & 1 also writes
  // the row below the range, the last CTA of a range with medge[z] & 2 the row above it (and gd of the next one):
  // the rows of E the later sweeps of this predict_correct read beyond the rows evaluated here
  const double *LU, *LV, *Lgd;
  double *MU, *MV, *Mgd;
  const double *lip;
  double ldt;
  int lqcon;
  int medge[2];
  Fold fold;   // MODE_S3A
  int tseq;    // timeline slot + 1 of the launch (GMD_TRACE builds), 0 = none
  int pdl;     // launched with programmatic stream serialization: wait for the predecessor inside the kernel
  unsigned flmask;   // row flags honoured by this launch: 0x0f filter rows, 0xf0 reduced rows
};

// beta of predict_correct (src/dycore_mod.F90:784-785) from the device-resident inner products
__device__ __forceinline__ double beta_from_ip(const double *ip, int qcon) {
  const double ip1 = ip[0], ip2 = ip[1];
  return (qcon && ip1 != 0.0 && ip2 != 0.0) ? ip1 / ip2 : 1.0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// programmatic dependent launch (the polar chain of a short band): a kernel launched with the programmatic stream
// serialization attribute may start before its predecessor in the stream has finished; pdl_wait() returns once the
// predecessor has completed and its writes are visible (no-op for an ordinary launch)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------
// Peer-memory signalling (latitude bands on the GPUs of one NVLink / NVSwitch node, one process per GPU).
// Every rank owns a SIGNAL PAGE (device memory, CUDA-IPC mapped by all ranks), 8-byte words:
// ---------------------------------------------------------------------------------------------------------
constexpr int MAXR = 8;   // ranks of one node
enum { SP_SIG = 0,                       // [2] halo epoch released by the south / north neighbour
       SP_XBASE = 2,                     // local: halo epochs completed before the current unit of work
       SP_RBASE = 3,                     // local: reductions completed before the current unit of work
       SP_TICKET = 4,                    // local: CTA ticket of the halo-push kernel
       SP_ERR = 5,                       // local: set when a wait timed out (reported by gmd_sync)
       SP_RFLAG = 8,                     // [2][MAXR] reduction epoch released by every rank (parity double buffer)
       SP_RSLOT = 8 + 2 * MAXR,          // [2][MAXR][2] doubles: the two partial sums of every rank
       SP_PEERTAB = 8 + 2 * MAXR + 4 * MAXR,  // [MAXR] local: every rank's signal page as mapped here (pointers)
       SP_WORDS = 8 + 2 * MAXR + 4 * MAXR + MAXR };

__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p) {
  u64 v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(u64 *p, u64 v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spin until *p >= want; a peer that never arrives (a bug, a dead rank) must not hang the GPU: give up after
// g_spin_timeout_ns (GMD_PEER_TIMEOUT_S, default 120 s of %globaltimer -- long enough for a neighbour that writes a
// history file or sits under a profiler), once (later waits return at once), flag the page and carry on; the results
// from then on are invalid and gmd_sync / gmd_step return GMD_ERR_COMM
__device__ unsigned long long g_spin_timeout_ns = 120ull * 1000000000ull;
__device__ __forceinline__ void spin_until(const u64 *p, u64 want, u64 *page) {
  if (ld_acquire_sys(p) >= want) return;
  if (*reinterpret_cast<volatile u64 *>(page + SP_ERR)) return;
  u64 t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (ld_acquire_sys(p) < want) {
    u64 t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > g_spin_timeout_ns) {
      *reinterpret_cast<volatile u64 *>(page + SP_ERR) = 1;
      return;
    }
  }
}

// deterministic block sum (fixed tree); result valid in thread 0; red must hold >= 32 doubles
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < NT / 32) ? red[l] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

// sum `n` pairs of partials in index order (deterministic) -> out[0], out[1]; with nranks > 1 the pair is then
// all-reduced over peer memory by the same CTA: every rank stores its pair into every rank's signal page,
// releases a flag, acquires the flags of all ranks and sums the pairs in rank order (same bits on every rank).
template <int NT>
__device__ __forceinline__ void reduce_pairs_body(const double *partials, int n, double *out, u64 *page, int rank,
                                                  int nranks, unsigned k, double *red, double *gath, int tseq = 0) {
  double a0 = 0.0, a1 = 0.0;
  for (int q = threadIdx.x; q < n; q += NT) {
    a0 += __ldcg(partials + 2 * q);
    a1 += __ldcg(partials + 2 * q + 1);
  }
  const double r0 = block_sum<NT>(a0, red);
  const double r1 = block_sum<NT>(a1, red);
  if (nranks <= 1) {
    if (threadIdx.x == 0) {
      out[0] = r0;
      out[1] = r1;
    }
    return;
  }
  if (threadIdx.x == 0) {
    gath[2 * MAXR] = r0;
    gath[2 * MAXR + 1] = r1;
  }
  __syncthreads();
  if ((int)threadIdx.x < nranks) {
    const int p = threadIdx.x;
    const u64 ep = page[SP_RBASE] + k;
    const int par = (int)(ep & 1);
    u64 *pp = reinterpret_cast<u64 *>(page[SP_PEERTAB + p]);
    volatile double *slot = reinterpret_cast<volatile double *>(pp + SP_RSLOT) + (par * MAXR + rank) * 2;
    slot[0] = gath[2 * MAXR];
    slot[1] = gath[2 * MAXR + 1];
    __threadfence_system();
    st_release_sys(pp + SP_RFLAG + par * MAXR + rank, ep);
    GMD_TRACE_T0();
    spin_until(page + SP_RFLAG + par * MAXR + p, ep, page);
    GMD_TRACE_WAIT(tseq);
    const volatile double *mine = reinterpret_cast<const volatile double *>(page + SP_RSLOT) + (par * MAXR + p) * 2;
    gath[2 * p] = mine[0];
    gath[2 * p + 1] = mine[1];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (int p = 0; p < nranks; p++) {
      s0 += gath[2 * p];
      s1 += gath[2 * p + 1];
    }
    out[0] = s0;
    out[1] = s1;
  }
}
// The inner products of predict_correct folded into the launches that produce their partial sums: every CTA of the
// S3a launches (boundary, interior, polar rows) takes a ticket after storing its partial pair; the last one sums all
// pairs in index order (so the result does not depend on which CTA that is) and runs the peer all-reduce.  Not
// inlined and fed by scalars only: the row loop of k_stage keeps its registers.
template <int NT>
__device__ __noinline__ void fold_tail(unsigned *ticket, unsigned total, const double *partials, int n, double *out,
                                       u64 *page, int rank, int nranks, unsigned k, int tseq = 0) {
  __shared__ double red[32];
  __shared__ double gath[2 * MAXR + 2];
  __shared__ int last;
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(ticket, 1u) == total - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  reduce_pairs_body<NT>(partials, n, out, page, rank, nranks, k, red, gath, tseq);
  if (threadIdx.x == 0) *ticket = 0;
}

// ---------------------------------------------------------------------------------------------------------
// Fused stage kernel: one operator evaluation (space_operators, src/dycore_mod.F90:184-365, with the seven
// operators :367-598 fused) + update_state (:600-652) + inner_product (src/types_mod.F90:347-371) in ONE
// sweep over the minimal state.
//
// Mapping ("warp marching"): a warp owns a strip of 64 columns (2 per lane, 16-byte loads) of which the inner
// WOUT = 60 are outputs (2 halo columns each side), and marches south -> north over `rows_per_cta` rows.
// Each lane keeps the rolling row window of ITS two columns in registers -- sqrt(gd) rows j-1..j+2, u rows
// j..j+1, v / U / V rows j-1..j+1, gd+ghs rows j..j+1 -- and gets the longitude neighbours with warp
// shuffles.  There is no shared memory and no block barrier on the path: warps are fully independent, every
// field element is read from HBM once (plus 4 halo columns per 64 and 3 prologue rows per chunk).
// ---------------------------------------------------------------------------------------------------------
struct D2 {
  double x, y;
};
__device__ __forceinline__ D2 ld2(const double *p) {
  const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
  D2 r;
  r.x = v.x;
  r.y = v.y;
  return r;
}
__device__ __forceinline__ void st2(double *p, double x, double y) {
  *reinterpret_cast<double2 *>(p) = make_double2(x, y);
}
#ifndef GMD_FAST_DIV
#define GMD_FAST_DIV (!GMD_STRICT)
#endif
#ifndef GMD_UNROLL
#define GMD_UNROLL 2
#endif
#ifndef GMD_MINB
#define GMD_MINB 4   // 4 CTAs x 128 threads per SM => <= 128 registers (measured best, profiles/r1_c_stage_tuning.txt)
#endif
#ifndef GMD_LAZY_UNROLL
#define GMD_LAZY_UNROLL 1
#endif
#ifndef GMD_LAZY_MINB
#define GMD_LAZY_MINB GMD_MINB
#endif
#ifndef GMD_S3A_MINB
#define GMD_S3A_MINB GMD_MINB
#endif
#ifndef GMD_S3A_UNROLL
#define GMD_S3A_UNROLL GMD_UNROLL
#endif
#ifndef GMD_S2_UNROLL
#define GMD_S2_UNROLL 1
#endif
#ifndef GMD_ALL_MINB
#define GMD_ALL_MINB GMD_MINB   // unsplit pass (advection + fast terms in one sweep): the widest register window
#endif
#define GMD_PRAGMA_(x) _Pragma(#x)
#define GMD_UNROLL_PRAGMA(n) GMD_PRAGMA_(unroll n)
// 2 a / b and sqrt(x) for the on-chip recomputation of u, v, sqrt(gd).  Strict build: IEEE division / sqrt.
// Product build: MUFU seed + Newton steps without the special-operand slow path (operands are O(1e2)
// geopotential roots, never 0, inf or denormal on rows where the result is used); error <= 1 ulp.
__device__ __forceinline__ double two_a_over_b(double a, double b) {
#if GMD_FAST_DIV
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  const double a2 = a + a;
  double q = a2 * r;
  const double rem = fma(-b, q, a2);
  return fma(rem, r, q);
#else
  return a * 2.0 / b;
#endif
}
__device__ __forceinline__ double fast_sqrt(double x) {
#if GMD_FAST_DIV
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  // y ~ x^-1/2 (about 20 bits): two Newton steps on y, then s = x y with one correction
  double h = 0.5 * y;
  double e = fma(-x * y, h, 0.5);   // 0.5 - 0.5 x y^2
  y = fma(y, e, y);
  h = 0.5 * y;
  e = fma(-x * y, h, 0.5);
  y = fma(y, e, y);
  double sr = x * y;
  const double rem = fma(-sr, sr, x);
  sr = fma(rem, 0.5 * y, sr);
  return (x > 0.0) ? sr : 0.0;
#else
  return sqrt(x);
#endif
}
__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// tendencies of ONE column at row j; all operands are scalars already in registers
template <int PASS, int ADV>
__device__ __forceinline__ void tend_col(
    const double *__restrict__ rc, const double *__restrict__ rcn /* record of row j+1 */, const bool rowU, const bool rowV,
    const bool rowG, const double bl, const double bt,
    const double hc0, const double hc1, const double hc2,
    // row j
    const double uw, const double uc, const double ue, const double Uw, const double Uc, const double Ue,
    const double Vw, const double Vc, const double Ve, const double sw, const double sc, const double se,
    const double vc, const double ve,
    // row j-1
    const double Us, const double Vs, const double Vse, const double vs, const double vse, const double ss,
    // row j+1
    const double Un, const double Unw, const double un, const double unw, const double Vn, const double vn,
    const double sn,
    // (gd+ghs)(i+1,j) - (gd+ghs)(i,j) and (gd+ghs)(i,j+1) - (gd+ghs)(i,j)
    const double ghdx, const double ghdy,
    // precomputed WENO advection terms
    const double aulon, const double aulat, const double avlon, const double avlat,
    double &dU, double &dV, double &dG) {
  dU = 0.0;
  dV = 0.0;
  dG = 0.0;
  // ===================== du, full rows 1..nlat-2 (src/dycore_mod.F90:207-210) ===========================
  if (rowU) {
    if (PASS != PASS_FAST) {
      double alon, alat;
      if (ADV == ADV_WENO) {
        alon = aulon;
        alat = aulat;
      } else {
        const double u1 = uc + uw, u2 = uc + ue;
        const double v1 = (vs + vse) * hc0;
        const double v2 = (vc + ve) * hc1;
        if (ADV == ADV_CENTER) {
          alon = rc[RC_Q_FDLON] * (u2 * Ue - u1 * Uw);   // :378-384
          alat = rc[RC_Q_FDLAT] * (v2 * Un - v1 * Us);   // :433-439
        } else {
          alon = rc[RC_Q_FDLON] * (u2 * (Uc + Ue) - bl * fabs(u2) * (Ue - Uc) - u1 * (Uc + Uw) +
                                 bl * fabs(u1) * (Uc - Uw) - (u2 - u1) * Uc);   // :395-404
          alat = rc[RC_Q_FDLAT] * (v2 * (Uc + Un) - bt * fabs(v2) * (Un - Uc) - v1 * (Uc + Us) +
                                 bt * fabs(v1) * (Uc - Us) - (v2 - v1) * Uc);   // :450-459
        }
      }
      dU = -alon - alat;
    }
    if (PASS != PASS_SLOW) {
      const double fv = 0.25 * (rc[RC_FF] + rc[RC_FC] * uc) * (rc[RC_COR1] * (Vs + Vse) + rc[RC_COR2] * (Vc + Ve));  // :485-493
#if GMD_STRICT
      const double pgf = 0.5 * (sc + se) / rc[RC_PGFU] * ghdx;  // :514-519
#else
      const double pgf = 0.5 * (sc + se) * rc[RC_PGFU] * ghdx;
#endif
      dU = (PASS == PASS_ALL) ? (dU + fv - pgf) : (fv - pgf);
    }
  }
  // ===================== dv, half rows 0..nlat-2 ==========================================================
  if (rowV) {
    if (PASS != PASS_FAST) {
      double alon, alat;
      if (ADV == ADV_WENO) {
        alon = avlon;
        alat = avlat;
      } else {
        const double u1 = uw + unw;
        const double u2 = uc + un;
        const double vh = vc * hc1;
        const double v1 = vh + vs * hc0;
        const double v2 = vh + vn * hc2;
        if (ADV == ADV_CENTER) {
          alon = rc[RC_Q_HDLON] * (u2 * Ve - u1 * Vw);   // :386-392
          alat = rc[RC_Q_HDLAT] * (v2 * Vn - v1 * Vs);   // :441-447
        } else {
          alon = rc[RC_Q_HDLON] * (u2 * (Vc + Ve) - bl * fabs(u2) * (Ve - Vc) - u1 * (Vc + Vw) +
                                 bl * fabs(u1) * (Vc - Vw) - (u2 - u1) * Vc);   // :406-415
          alat = rc[RC_Q_HDLAT] * (v2 * (Vc + Vn) - bt * fabs(v2) * (Vn - Vc) - v1 * (Vc + Vs) +
                                 bt * fabs(v1) * (Vc - Vs) - (v2 - v1) * Vc);   // :461-470
        }
      }
      dV = -alon - alat;
    }
    if (PASS != PASS_SLOW) {
      const double f0 = rc[RC_FF], c0 = rc[RC_FC], f1 = rcn[RC_FF], c1 = rcn[RC_FC];
      const double fu = 0.25 * ((f0 + c0 * uc) * Uc + (f0 + c0 * uw) * Uw + (f1 + c1 * un) * Un +
                                (f1 + c1 * unw) * Unw);   // :495-503
#if GMD_STRICT
      const double pgf = 0.5 * (sc + sn) / rc[RC_PGFV] * hc1 * ghdy;  // :530-535
#else
      const double pgf = 0.5 * (sc + sn) * rc[RC_PGFV] * ghdy;
#endif
      dV = (PASS == PASS_ALL) ? (dV - fu - pgf) : (-fu - pgf);
    }
  }
  // ===================== dgd, full rows 1..nlat-2 (pole rows: k_polar) ====================================
  if (rowG) {
#if GMD_STRICT
    const double mlon = ((sc + se) * Uc - (sc + sw) * Uw) * 0.5 / rc[RC_MLON];                   // :546-552
    const double mlat = ((sc + sn) * Vc * hc1 - (sc + ss) * Vs * hc0) * 0.5 / rc[RC_MLAT];       // :564-570
#else
    const double mlon = ((sc + se) * Uc - (sc + sw) * Uw) * rc[RC_MLON];
    const double mlat = ((sc + sn) * Vc * hc1 - (sc + ss) * Vs * hc0) * rc[RC_MLAT];
#endif
    dG = -mlon - mlat;
  }
}

constexpr int WOUT = 60;  // output columns per warp strip (64 held)
constexpr int SW = 4;     // warps per CTA
constexpr int BX = SW * 32;

// LAZY: 0 = E is read as stored; 1 = E = base + beta ldt L (U, V and gd); 2 = the same for U, V only (the previous
// predict_correct was a slow pass: gd unchanged)
// PUSH (MODE_S3A of a multi-rank run): the band-edge rows of the tendency also go to the neighbours (a separate
// instantiation, so that the row loop of the single-GPU kernels keeps all of its registers)
// resident CTAs per SM the register allocation is capped for
template <int PASS, int MODE, int LAZY>
constexpr int stage_minb() {
  return PASS == 0 /* PASS_ALL */ ? GMD_ALL_MINB : (LAZY ? GMD_LAZY_MINB : (MODE == 2 /* MODE_S3A */ ? GMD_S3A_MINB : GMD_MINB));
}
// ---- row-packet ring (cp.async) ---------------------------------------------------------------------------------
// Everything one iteration of the row loop consumes from global memory -- gd(j+2), U(j+1), V(j+1), ghs(j+1) of the
// evaluated state, the deferred update's tendency rows, the base-state / previous-tendency rows of row j -- is one
// PACKET of up to 7 x 16 bytes per lane.  Packets are copied to a per-warp shared-memory ring with cp.async
// (LDGSTS.128, L2 only) RING_DEPTH - 1 iterations ahead and read back by the lane that copied them, so the only
// synchronisation is the lane's own cp.async.wait_group.  Against the register prefetch this replaces (one row
// ahead, 7 double2 of live registers) the loads are in flight 2 rows ahead and the row loop has ~28 registers more.
#ifndef GMD_RING
#define GMD_RING 0   // measured slower than the register prefetch (profiles/r2_h_stage_ring_tuning.txt)
#endif
#ifndef GMD_RING_DEPTH
#define GMD_RING_DEPTH 3
#endif
constexpr int RING_DEPTH = GMD_RING_DEPTH;
template <int PASS, int MODE, int LAZY>
struct Packet {
  static constexpr bool gh = (PASS != PASS_SLOW);
  static constexpr bool upd = (MODE == MODE_S1 || MODE == MODE_S2);
  static constexpr bool q = (upd && !LAZY) || (MODE == MODE_S3A);   // base state (update) or previous tendency (dots)
  static constexpr int GD2 = 0, U1 = 1, V1 = 2, HS = 3;
  static constexpr int nE = 3 + (gh ? 1 : 0);
  static constexpr int TG = nE, TU = nE + (LAZY == 1 ? 1 : 0), TV = TU + 1;
  static constexpr int nL = LAZY ? (LAZY == 1 ? 3 : 2) : 0;
  static constexpr int QU = nE + nL, QV = QU + 1, QG = QU + 2;
  static constexpr int nQ = q ? (gh ? 3 : 2) : 0;
  static constexpr int NF = nE + nL + nQ;
};
constexpr int RING_NF_MAX = 7;
constexpr size_t RING_BYTES = (size_t)SW * RING_DEPTH * RING_NF_MAX * 32 * 16;   // per CTA
__device__ __forceinline__ void cp16(unsigned dst, const double *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// The sweep of one CTA: (bx, by, bz) = strip group, row chunk and row range of the CTA, gx = strip groups per chunk
// row; red: 2 SW doubles, srow: (rows_per_cta + 2) row records of shared memory.  k_stage runs it on its own block
// index, k_cap (below) on a linear CTA index, followed by the polar rows of the same sweep.
template <int PASS, int ADV, int MODE, int LAZY, bool PUSH>
__device__ __forceinline__ void stage_body(const StageArgs &a, const int bx, const int by, const int bz, const int gx,
                                           double *red, double *srow, double *ringmem) {
  const int nlon = a.g.nlon, nlat = a.g.nlat, r0 = a.g.r0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int strip = bx * SW + warp;
  const int nstrips = (nlon + WOUT - 1) / WOUT;
  const int ja = a.rb[bz] + by * a.rows_per_cta;
  const int jb = min(ja + a.rows_per_cta, a.re[bz]);
  if (ja >= jb) {  // whole CTA: this range has fewer chunks than the launch has chunk rows
    if (MODE == MODE_S3A && threadIdx.x == 0) {
      const size_t b = (size_t)a.pofs[bz] + (size_t)by * gx + bx;
      a.partials[2 * b] = 0.0;
      a.partials[2 * b + 1] = 0.0;
    }
    if (MODE == MODE_S3A && a.fold.ticket)
      fold_tail<BX>(a.fold.ticket, a.fold.total, a.partials, a.fold.n, a.fold.out, a.fold.r.page, a.fold.r.rank, a.fold.r.nranks,
                    a.fold.r.k, a.tseq);
    return;
  }
  const bool need_gh = (PASS != PASS_SLOW);
  const bool upd = (MODE == MODE_S1 || MODE == MODE_S2);
  double ip1 = 0.0, ip2 = 0.0;
  {
    const int nrec = (jb - ja + 2) * RC_N;
    const double *__restrict__ src = a.t.rowrec + (ptrdiff_t)(ja - 1) * RC_N;
    for (int k = threadIdx.x; k < nrec; k += BX) srow[k] = __ldg(src + k);
  }
  __syncthreads();

  if (strip < nstrips) {
    int c0 = strip * WOUT - 2 + 2 * lane;  // even; (c0, c0+1) never straddles the seam because nlon is even
    c0 %= nlon;
    if (c0 < 0) c0 += nlon;
    const bool out = (lane >= 1) && (lane <= WOUT / 2) && (strip * WOUT + 2 * (lane - 1) < nlon);
    // element offset of (row j, column c0) is `off`, advanced by nlon per row; AT addresses row j + k
    const ptrdiff_t nl = nlon;
    ptrdiff_t off = (ptrdiff_t)(ja - r0) * nl + (ptrdiff_t)c0;
#define ATK(p, k) ((p) + (off + (ptrdiff_t)(k) * nl))
#define AT(p, jj) ATK(p, (jj) - j)
    const D2 zero2 = {0.0, 0.0};
    // ---- deferred update: E(row r) = base + bdt * L on the rows where update_state applies, stored to M on the rows
    //      this CTA is responsible for (its own rows; the band's ghost rows for the CTAs at the band edges)
    double bdt = 0.0;
    if (LAZY) bdt = a.ldt * beta_from_ip(a.lip, a.lqcon);
    const bool edgeS = LAZY && (ja == a.rb[bz]) && (a.medge[bz] & 1);
    const bool edgeN = LAZY && (jb == a.re[bz]) && (a.medge[bz] & 2);
    auto mineUV = [&](int r) { return (r >= ja && r < jb) || (edgeS && r == ja - 1) || (edgeN && r == jb); };
    auto mineG = [&](int r) { return (r >= ja && r < jb) || (edgeS && r == ja - 1) || (edgeN && (r == jb || r == jb + 1)); };
    auto combU = [&](D2 b, D2 t, int r) -> D2 {
      if (r >= 1 && r <= nlat - 2) { b.x = fma(bdt, t.x, b.x); b.y = fma(bdt, t.y, b.y); }
      return b;
    };
    auto combV = [&](D2 b, D2 t, int r) -> D2 {
      if (r >= 0 && r <= nlat - 2) { b.x = fma(bdt, t.x, b.x); b.y = fma(bdt, t.y, b.y); }
      return b;
    };
    auto combG = [&](D2 b, D2 t, int r) -> D2 {
      if (r >= 0 && r <= nlat - 1) { b.x = fma(bdt, t.x, b.x); b.y = fma(bdt, t.y, b.y); }
      return b;
    };
#if GMD_RING
    typedef Packet<PASS, MODE, LAZY> PK;
    // this lane's slot of (stage, field): lanes 16 bytes apart => conflict-free LDS.128 / LDGSTS.128
    D2 *const ring = reinterpret_cast<D2 *>(ringmem) + (size_t)warp * (RING_DEPTH * PK::NF * 32) + lane;
    const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring);
    // packet of iteration r (the loop below is at row j; AT addresses relative to it) into ring stage st
    auto issue_packet = [&](const int r, const int st, const int j) {
      const unsigned d = ring_sa + (unsigned)(st * PK::NF * 32 * 16);
      cp16(d + PK::GD2 * 512, AT(a.Egd, r + 2));
      cp16(d + PK::U1 * 512, AT(a.EU, r + 1));
      cp16(d + PK::V1 * 512, AT(a.EV, r + 1));
      if (PK::gh) cp16(d + PK::HS * 512, AT(a.ghs, r + 1));
      if (LAZY) {
        if (LAZY == 1) cp16(d + PK::TG * 512, AT(a.Lgd, r + 2));
        cp16(d + PK::TU * 512, AT(a.LU, r + 1));
        cp16(d + PK::TV * 512, AT(a.LV, r + 1));
      }
      if (PK::q && out) {
        const double *qU = PK::upd ? a.OU : a.PU, *qV = PK::upd ? a.OV : a.PV, *qG = PK::upd ? a.Ogd : a.Pgd;
        cp16(d + PK::QU * 512, AT(qU, r));
        cp16(d + PK::QV * 512, AT(qV, r));
        if (PK::gh) cp16(d + PK::QG * 512, AT(qG, r));
      }
    };
#pragma unroll
    for (int q = 0; q < RING_DEPTH - 1; q++) {   // the first packets travel with the prologue's own loads
      if (ja + q < jb) issue_packet(ja + q, q, ja);
      cp_commit();
    }
    int rst = 0;   // ring stage of the current iteration's packet
#endif
    // ---- prologue: rows ja-1, ja, ja+1 of sqrt(gd); rows ja-1, ja of U, V; gh row ja --------------------
    D2 sm_, s0, sp, sq, u0, up, vm, v0, vp, Um, U0, Up, Vm, V0, Vp;
    D2 g0 = zero2, gp = zero2;
    D2 graw = zero2;   // LAZY: gd of row j (the base of this row's update)
    D2 a2keep = zero2;
#if GMD_STRICT
    D2 h0 = zero2, hp = zero2;  // ghs rows j, j+1 (g0/gp then hold gd alone)
#endif
    {
      const int j = ja;
      D2 a0 = ld2(AT(a.Egd, ja - 1)), a1 = ld2(AT(a.Egd, ja)), a2 = ld2(AT(a.Egd, ja + 1));
      Um = ld2(AT(a.EU, ja - 1));
      U0 = ld2(AT(a.EU, ja));
      Vm = ld2(AT(a.EV, ja - 1));
      V0 = ld2(AT(a.EV, ja));
      if (LAZY) {
        const D2 tUm = ld2(AT(a.LU, ja - 1)), tU0 = ld2(AT(a.LU, ja)), tVm = ld2(AT(a.LV, ja - 1)), tV0 = ld2(AT(a.LV, ja));
        if (LAZY == 1) {
          const D2 t0 = ld2(AT(a.Lgd, ja - 1)), t1 = ld2(AT(a.Lgd, ja)), t2 = ld2(AT(a.Lgd, ja + 1));
          a0 = combG(a0, t0, ja - 1);
          a1 = combG(a1, t1, ja);
          a2 = combG(a2, t2, ja + 1);
          if (out) {
            if (mineG(ja - 1)) st2(AT(a.Mgd, ja - 1), a0.x, a0.y);
            st2(AT(a.Mgd, ja), a1.x, a1.y);
            if (mineG(ja + 1)) st2(AT(a.Mgd, ja + 1), a2.x, a2.y);
          }
        }
        Um = combU(Um, tUm, ja - 1);
        U0 = combU(U0, tU0, ja);
        Vm = combV(Vm, tVm, ja - 1);
        V0 = combV(V0, tV0, ja);
        if (out) {
          if (mineUV(ja - 1)) {
            st2(AT(a.MU, ja - 1), Um.x, Um.y);
            st2(AT(a.MV, ja - 1), Vm.x, Vm.y);
          }
          st2(AT(a.MU, ja), U0.x, U0.y);
          st2(AT(a.MV, ja), V0.x, V0.y);
        }
        graw = a1;
      }
      a2keep = a2;
      sm_.x = fast_sqrt(a0.x); sm_.y = fast_sqrt(a0.y);
      s0.x = fast_sqrt(a1.x); s0.y = fast_sqrt(a1.y);
      sp.x = fast_sqrt(a2.x); sp.y = fast_sqrt(a2.y);
      if (need_gh) {
        const D2 hs = ld2(AT(a.ghs, ja));
#if GMD_STRICT
        g0 = a1;
        h0 = hs;
#else
        g0.x = a1.x + hs.x;
        g0.y = a1.y + hs.y;
#endif
      }
      // u(ja), v(ja-1), v(ja): update_state :636-645 (u = 2U/(s_i + s_i+1), v = 2V/(s_j + s_j+1))
      const double s0e = shfl_dn1(s0.x);
      u0.x = two_a_over_b(U0.x, s0.x + s0.y);
      u0.y = two_a_over_b(U0.y, s0.y + s0e);
      const bool vmok = (ja - 1 >= 0), v0ok = (ja < nlat - 1);
      vm.x = vmok ? two_a_over_b(Vm.x, sm_.x + s0.x) : 0.0;
      vm.y = vmok ? two_a_over_b(Vm.y, sm_.y + s0.y) : 0.0;
      v0.x = v0ok ? two_a_over_b(V0.x, s0.x + sp.x) : 0.0;
      v0.y = v0ok ? two_a_over_b(V0.y, s0.y + sp.y) : 0.0;
    }
    // longitude-neighbour values carried from one row to the next
    double uw_a = shfl_up1(u0.y);    // u(i-1, j)   for column a
    double Uw_a = shfl_up1(U0.y);    // U(i-1, j)
    double Vse_b = shfl_dn1(Vm.x);   // V(i+1, j-1) for column b
    double vse_b = shfl_dn1(vm.x);   // v(i+1, j-1)
    double se_b = shfl_dn1(s0.x);    // s(i+1, j)
#if GMD_RING
    D2 n_gd1 = a2keep;   // gd(ja+1) (with the deferred update applied)
#else
    // first prefetch: gd(ja+2), U(ja+1), V(ja+1), gd(ja+1), ghs(ja+1)
    D2 n_gd2 = ld2(ATK(a.Egd, 2));
    D2 n_U = ld2(ATK(a.EU, 1));
    D2 n_V = ld2(ATK(a.EV, 1));
    D2 n_gd1 = zero2, n_hs = zero2;
    D2 n_tg = zero2, n_tU = zero2, n_tV = zero2;   // LAZY: the tendency rows travelling with n_gd2, n_U, n_V
    if (LAZY) {
      if (LAZY == 1) n_tg = ld2(ATK(a.Lgd, 2));
      n_tU = ld2(ATK(a.LU, 1));
      n_tV = ld2(ATK(a.LV, 1));
    }
    if (need_gh) {
      n_gd1 = LAZY ? a2keep : ld2(ATK(a.Egd, 1));
      n_hs = ld2(ATK(a.ghs, 1));
    }
#endif

    // measured on B200 (tools/tune_stage.py): the deferred-update variants and S2 are fastest without unrolling (no
    // spills at the 128-register cap), S1 / S3a with two rows per trip
    constexpr int kUnroll = LAZY ? GMD_LAZY_UNROLL : (MODE == MODE_S2 ? GMD_S2_UNROLL : (MODE == MODE_S3A ? GMD_S3A_UNROLL : GMD_UNROLL));
#pragma unroll kUnroll
    for (int j = ja; j < jb; j++) {
      // ---- issue every load of this iteration first: the next row of the evaluated state (consumed one
      //      iteration later) and the operands of this row's update (consumed at the end of this iteration) ---
#if GMD_RING
      {   // packet j + DEPTH - 1 goes into the stage read one iteration ago; then this iteration's packet must have landed
        const int ist = (rst == 0) ? RING_DEPTH - 1 : rst - 1;
        if (j + (RING_DEPTH - 1) < jb) issue_packet(j + (RING_DEPTH - 1), ist, j);
        cp_commit();
        cp_wait<RING_DEPTH - 1>();
      }
      const D2 *const pk = ring + rst * (PK::NF * 32);
      rst = (rst == RING_DEPTH - 1) ? 0 : rst + 1;
      D2 c_gd2 = pk[PK::GD2 * 32], c_U = pk[PK::U1 * 32], c_V = pk[PK::V1 * 32];
      const D2 c_gd1 = n_gd1;
      const D2 c_hs = PK::gh ? pk[PK::HS * 32] : zero2;
      D2 c_tg = zero2, c_tU = zero2, c_tV = zero2;
      if (LAZY) {
        if (LAZY == 1) c_tg = pk[PK::TG * 32];
        c_tU = pk[PK::TU * 32];
        c_tV = pk[PK::TV * 32];
      }
#else
      D2 c_gd2 = n_gd2, c_U = n_U, c_V = n_V;
      const D2 c_gd1 = n_gd1, c_hs = n_hs;
      const D2 c_tg = n_tg, c_tU = n_tU, c_tV = n_tV;
      if (j + 1 < jb) {
        n_gd2 = ld2(AT(a.Egd, j + 3));
        n_U = ld2(AT(a.EU, j + 2));
        n_V = ld2(AT(a.EV, j + 2));
        if (need_gh) n_hs = ld2(AT(a.ghs, j + 2));
        if (LAZY) {
          if (LAZY == 1) n_tg = ld2(AT(a.Lgd, j + 3));
          n_tU = ld2(AT(a.LU, j + 2));
          n_tV = ld2(AT(a.LV, j + 2));
        }
      }
#endif
      if (LAZY) {
        if (LAZY == 1) {
          c_gd2 = combG(c_gd2, c_tg, j + 2);
          if (out && mineG(j + 2)) st2(AT(a.Mgd, j + 2), c_gd2.x, c_gd2.y);
        }
        c_U = combU(c_U, c_tU, j + 1);
        c_V = combV(c_V, c_tV, j + 1);
        if (out && mineUV(j + 1)) {
          st2(AT(a.MU, j + 1), c_U.x, c_U.y);
          st2(AT(a.MV, j + 1), c_V.x, c_V.y);
        }
      }
      if (need_gh) n_gd1 = c_gd2;  // gd(j+2) is next iteration's gd(j'+1)
      const double *__restrict__ rc = srow + (j - ja + 1) * RC_N;
      const unsigned flraw = (unsigned)rc[RC_FLAGS] & a.flmask;
      const unsigned fl = flraw | (flraw >> FL_REDUCE_SHIFT);
      const bool rowU = (j >= 1 && j <= nlat - 2);
      const bool rowV = (j <= nlat - 2);
      const bool rowG = rowU && (PASS != PASS_SLOW);
      D2 oU = zero2, oV = zero2, oG = zero2, pU = zero2, pV = zero2, pG = zero2;
      D2 wul = zero2, wut = zero2, wvl = zero2, wvt = zero2;
      if (out) {
#if GMD_RING
        if (upd && !LAZY) {
          oU = pk[PK::QU * 32];
          oV = pk[PK::QV * 32];
          if (PK::gh) oG = pk[PK::QG * 32];
        }
        if (MODE == MODE_S3A) {
          pU = pk[PK::QU * 32];
          pV = pk[PK::QV * 32];
          if (PK::gh) pG = pk[PK::QG * 32];
        }
#else
        if (upd && !LAZY) {
          oU = ld2(AT(a.OU, j));
          if (rowV) oV = ld2(AT(a.OV, j));
          if (rowG) oG = ld2(AT(a.Ogd, j));
        }
        if (MODE == MODE_S3A) {
          if (rowU) pU = ld2(AT(a.PU, j));
          if (rowV) pV = ld2(AT(a.PV, j));
          if (rowG) pG = ld2(AT(a.Pgd, j));
        }
#endif
        if (upd && LAZY) {  // old state == evaluated state, already on chip
          oU = U0;
          oV = V0;
          oG = graw;
        }
        if (ADV == ADV_WENO && PASS != PASS_FAST) {
          if (rowU) {
            wul = ld2(AT(a.AUlon, j));
            wut = ld2(AT(a.AUlat, j));
          }
          if (rowV) {
            wvl = ld2(AT(a.AVlon, j));
            wvt = ld2(AT(a.AVlat, j));
          }
        }
      }
      // ---- advance the window: s(j+2), U(j+1), V(j+1), gh(j+1), u(j+1), v(j+1) --------------------------
      sq.x = fast_sqrt(c_gd2.x);
      sq.y = fast_sqrt(c_gd2.y);
      Up = c_U;
      Vp = c_V;
      if (need_gh) {
#if GMD_STRICT
        gp = c_gd1;
        hp = c_hs;
#else
        gp.x = c_gd1.x + c_hs.x;
        gp.y = c_gd1.y + c_hs.y;
#endif
      }
      const double spe_b = shfl_dn1(sp.x);  // s(i+1, j+1) for column b
      {
        const bool uok = (j + 1 < nlat), vok = (j + 1 < nlat - 1);
        up.x = uok ? two_a_over_b(Up.x, sp.x + sp.y) : 0.0;
        up.y = uok ? two_a_over_b(Up.y, sp.y + spe_b) : 0.0;
        vp.x = vok ? two_a_over_b(Vp.x, sp.x + sq.x) : 0.0;
        vp.y = vok ? two_a_over_b(Vp.y, sp.y + sq.y) : 0.0;
      }
      // ---- longitude neighbours of row j / j+1 ------------------------------------------------------------
      const double unw_a = shfl_up1(up.y);   // u(i-1, j+1) for a
      const double Unw_a = shfl_up1(Up.y);   // U(i-1, j+1)
      const double Vw_a = shfl_up1(V0.y);    // V(i-1, j)
      const double sw_a = shfl_up1(s0.y);    // s(i-1, j)
      const double ue_b = shfl_dn1(u0.x);    // u(i+1, j) for b
      const double Ue_b = shfl_dn1(U0.x);
      const double Ve_b = shfl_dn1(V0.x);
      const double ve_b = shfl_dn1(v0.x);
      double ghdx_a = 0.0, ghdx_b = 0.0, ghdy_a = 0.0, ghdy_b = 0.0;
      if (need_gh) {
#if GMD_STRICT
        const double ge_b = shfl_dn1(g0.x), he_b = shfl_dn1(h0.x);
        // gd(b) + ghs(b) - gd(a) - ghs(a), left to right (src/dycore_mod.F90:516-517,532-533)
        ghdx_a = ((g0.y + h0.y) - g0.x) - h0.x;
        ghdx_b = ((ge_b + he_b) - g0.y) - h0.y;
        ghdy_a = ((gp.x + hp.x) - g0.x) - h0.x;
        ghdy_b = ((gp.y + hp.y) - g0.y) - h0.y;
#else
        const double ge_b = shfl_dn1(g0.x);
        ghdx_a = g0.y - g0.x;
        ghdx_b = ge_b - g0.y;
        ghdy_a = gp.x - g0.x;
        ghdy_b = gp.y - g0.y;
#endif
      }
      const double hc0 = rc[RC_COSH - RC_N], hc1 = rc[RC_COSH], hc2 = rc[RC_COSH + RC_N];
      const double cfj = rc[RC_COSF];
      double dUa, dVa, dGa, dUb, dVb, dGb;
      // column a: west = lane-1's b (shuffled / carried), east = own b
      tend_col<PASS, ADV>(rc, rc + RC_N, rowU, rowV, rowG, a.beta_lon, a.beta_lat, hc0, hc1, hc2,
                          uw_a, u0.x, u0.y, Uw_a, U0.x, U0.y, Vw_a, V0.x, V0.y, sw_a, s0.x, s0.y, v0.x, v0.y,
                          Um.x, Vm.x, Vm.y, vm.x, vm.y, sm_.x,
                          Up.x, Unw_a, up.x, unw_a, Vp.x, vp.x, sp.x,
                          ghdx_a, ghdy_a, wul.x, wut.x, wvl.x, wvt.x, dUa, dVa, dGa);
      // column b: west = own a, east = lane+1's a (shuffled / carried)
      tend_col<PASS, ADV>(rc, rc + RC_N, rowU, rowV, rowG, a.beta_lon, a.beta_lat, hc0, hc1, hc2,
                          u0.x, u0.y, ue_b, U0.x, U0.y, Ue_b, V0.x, V0.y, Ve_b, s0.x, s0.y, se_b, v0.y, ve_b,
                          Um.y, Vm.y, Vse_b, vm.y, vse_b, sm_.y,
                          Up.y, Up.x, up.y, up.x, Vp.y, vp.y, sp.y,
                          ghdx_b, ghdy_b, wul.y, wut.y, wvl.y, wvt.y, dUb, dVb, dGb);
      if (out) {
        if (rowU) {
          if (fl & FL_DU) {
            st2(AT(a.TU, j), dUa, dUb);  // filtered + applied by k_polar
          } else {
            if (upd) st2(AT(a.NU, j), oU.x + a.dt * dUa, oU.y + a.dt * dUb);
            if (MODE != MODE_S1) st2(AT(a.TU, j), dUa, dUb);
            if (MODE == MODE_S3A) {
              ip1 = ip1 + dUa * pU.x * cfj;
              ip2 = ip2 + dUa * dUa * cfj;
              ip1 = ip1 + dUb * pU.y * cfj;
              ip2 = ip2 + dUb * dUb * cfj;
            }
          }
        } else if (upd) {
          // pole rows: du is never written there (stays 0), so U' = U (src/dycore_mod.F90:623-627)
          st2(AT(a.NU, j), oU.x, oU.y);
        }
        if (rowV) {
          if (fl & FL_DV) {
            st2(AT(a.TV, j), dVa, dVb);
          } else {
            if (upd) st2(AT(a.NV, j), oV.x + a.dt * dVa, oV.y + a.dt * dVb);
            if (MODE != MODE_S1) st2(AT(a.TV, j), dVa, dVb);
            if (MODE == MODE_S3A) {
              ip1 = ip1 + dVa * pV.x * hc1;
              ip2 = ip2 + dVa * dVa * hc1;
              ip1 = ip1 + dVb * pV.y * hc1;
              ip2 = ip2 + dVb * dVb * hc1;
            }
          }
        }
        if (rowG) {
          if (fl & FL_DGD) {
            st2(AT(a.Tgd, j), dGa, dGb);
          } else {
            if (upd) st2(AT(a.Ngd, j), oG.x + a.dt * dGa, oG.y + a.dt * dGb);
            if (MODE != MODE_S1) st2(AT(a.Tgd, j), dGa, dGb);
            if (MODE == MODE_S3A) {
              ip1 = ip1 + dGa * pG.x * cfj;
              ip2 = ip2 + dGa * dGa * cfj;
              ip1 = ip1 + dGb * pG.y * cfj;
              ip2 = ip2 + dGb * dGb * cfj;
            }
          }
        }
      }
      // ---- band-edge rows of the tendency straight into the neighbours' ghost rows (the lowest HALO_N rows go south,
      //      the highest HALO_S rows north); these rows are never filtered rows (checked by gmd_peer_connect)
      if (PUSH && MODE == MODE_S3A && out) {
        if (j < a.push_s_end && a.hpS_U != nullptr) {
          st2(AT(a.hpS_U, j), dUa, dUb);
          st2(AT(a.hpS_V, j), dVa, dVb);
          if (rowG) st2(AT(a.hpS_G, j), dGa, dGb);
        }
        if (j >= a.push_n_begin && a.hpN_U != nullptr) {
          st2(AT(a.hpN_U, j), dUa, dUb);
          st2(AT(a.hpN_V, j), dVa, dVb);
          if (rowG) st2(AT(a.hpN_G, j), dGa, dGb);
        }
      }
      // ---- rotate the window ------------------------------------------------------------------------------
      sm_ = s0; s0 = sp; sp = sq;
      u0 = up;
      vm = v0; v0 = vp;
      Um = U0; U0 = Up;
      Vm = V0; V0 = Vp;
      g0 = gp;
#if GMD_STRICT
      h0 = hp;
#endif
      if (LAZY) graw = c_gd1;
      uw_a = unw_a;
      Uw_a = Unw_a;
      Vse_b = Ve_b;
      vse_b = ve_b;
      se_b = spe_b;
      off += nl;
    }
#undef AT
#undef ATK
  }
  if (MODE == MODE_S3A) {
    const double r1 = warp_sum(ip1), r2 = warp_sum(ip2);
    if (lane == 0) {
      red[2 * warp] = r1;
      red[2 * warp + 1] = r2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double q1 = 0.0, q2 = 0.0;
      for (int w = 0; w < SW; w++) {
        q1 += red[2 * w];
        q2 += red[2 * w + 1];
      }
      const size_t b = (size_t)a.pofs[bz] + (size_t)by * gx + bx;
      a.partials[2 * b] = q1;
      a.partials[2 * b + 1] = q2;
      // this CTA's peer stores (all issued before the barrier above) must be visible to the neighbours before the ticket
      // below / the end of the kernel lets the reduction release them; CTAs that pushed nothing skip the (slow) fence
      if (PUSH && ((a.hpS_U != nullptr && ja < a.push_s_end) || (a.hpN_U != nullptr && jb > a.push_n_begin))) __threadfence_system();
    }
    if (a.fold.ticket)
      fold_tail<BX>(a.fold.ticket, a.fold.total, a.partials, a.fold.n, a.fold.out, a.fold.r.page, a.fold.r.rank, a.fold.r.nranks,
                    a.fold.r.k, a.tseq);
  }
}

template <int PASS, int ADV, int MODE, int LAZY = 0, bool PUSH = false>
__global__ void __launch_bounds__(BX, (stage_minb<PASS, MODE, LAZY>())) k_stage(const StageArgs a) {
  __shared__ double red[2 * SW];
  extern __shared__ __align__(16) double srow[];  // row records of rows ja-1 .. jb: [(rows_per_cta + 2)][RC_N], then the ring
  trace_in(a.tseq);
  if (a.pdl) {
    pdl_trigger();
    pdl_wait();
  }
  stage_body<PASS, ADV, MODE, LAZY, PUSH>(a, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.x, red, srow,
                                          srow + (size_t)(a.rows_per_cta + 2) * RC_N);
  trace_out(a.tseq);
}

// ---------------------------------------------------------------------------------------------------------
// Polar rows: the SMOOTHING blocks of space_operators (src/dycore_mod.F90:212-219,228-235,244-251) with
// filter_array_at_{full,half}_lat (src/filter_mod.F90:105-167), and the pole caps of
// meridional_mass_divergence_operator (src/dycore_mod.F90:572-596).  One CTA per (row, field) item.
//
// The FFTPACK forward -> 0/1 mask -> backward round trip keeps halfcomplex entries 1..2(c+1), i.e. it is the
// orthogonal projector onto {1, cos k x, sin k x (k<=c), cos (c+1) x}.  The projector is applied directly
// (2c+2 dot products with a host-precomputed basis + reconstruction): same result to rounding, no
// butterflies, and the row never leaves shared memory.
// ---------------------------------------------------------------------------------------------------------
constexpr int PT = 512;       // threads per polar-row CTA (one CTA per SM: <= 128 registers per thread)
constexpr int PT_EW = 256;    // threads of the small pole-cap kernels
constexpr int PB = 4;         // row elements per thread per batch (independent loads in flight)
constexpr int KF = 6;         // fast projector: wavenumbers 1..KF, i.e. cutoff <= KF-1 ...
constexpr int PQ = 16;        // ... on rows of up to PT*PQ elements
constexpr int MAX_ITEMS = 128;
enum { IT_DU = 0, IT_DV = 1, IT_DGD = 2, IT_POLE_S = 3, IT_POLE_N = 4 };

// one (row, field) work item, packed so that the whole list travels in the kernel parameters (no dependent
// global load at the head of the latency chain): bits 0-15 row, 16-24 cutoff+1, 28-30 kind
__host__ __device__ inline unsigned pack_item(int kind, int row, int cutoff) {
  return (unsigned)row | ((unsigned)(cutoff + 1) << 16) | ((unsigned)kind << 28);
}
// a row of the moving reduced tendency: bit 27 set, the cutoff field carries the reduction factor
constexpr unsigned ITEM_REDUCE = 1u << 27;
__host__ __device__ inline unsigned pack_reduce_item(int kind, int row, int factor) {
  return (unsigned)row | ((unsigned)factor << 16) | ITEM_REDUCE | ((unsigned)kind << 28);
}

struct PolarArgs {
  Geo g;
  Tab t;
  const double *basis;  // [2*cmax+3][nlon]: row 0 = 1, 2k-1 = cos(k x_i), 2k = sin(k x_i), x_i = 2 pi i / nlon
  const double *rot;    // [PQ][KF][2]: cos, sin of 2 pi k (q PT) / nlon  (k = 1..KF): basis(i0 + q PT) from basis(i0)
  const double *EU, *EV, *Egd, *ghs;
  const double *OU, *OV, *Ogd;
  double *NU, *NV, *Ngd;
  double *TU, *TV, *Tgd;
  const double *PU, *PV, *Pgd;
  double dt;
  double *partials;       // [nitems][2]
  int rescale;            // 1: tendency filter with inner-product rescale; 0: plain filter (diffusion)
  int use_q;              // dynamic shared memory holds a third row (prefetched base-state / previous-tendency row)
  double radius, dlat;
  Fold fold;              // MODE_S3A: see k_stage
  const double *fold_partials;   // start of the whole partials array (this kernel's own pairs start at `partials`)
  int tseq;               // timeline slot + 1 (GMD_TRACE builds), 0 = none
  int pdl;                // see StageArgs
  int reduce_smooth;      // use_reduce_tend_smooth: reduced rows are rescaled like filter rows (when `rescale` is set)
  unsigned items[MAX_ITEMS];
};

// General projector (any cutoff, any row length): project x[0..n) (shared) onto the kept modes, result overwrites
// x.  The 2c+2 dot products are spread over the CTA's warps (one (coefficient, segment) item per warp, fixed
// summation order => deterministic) with the basis streamed from L2.
__device__ inline void project_row(double *x, int n, int cutoff, const double *__restrict__ basis, double *coef,
                                   double *part) {
  const int K = cutoff + 1;  // highest (cosine-only) wavenumber kept
  if (cutoff < 0) {          // all-zero mask
    for (int i = threadIdx.x; i < n; i += PT) x[i] = 0.0;
    __syncthreads();
    return;
  }
  if (2 * K >= n) return;    // mask keeps every entry: the FFT round trip is the identity
  const int ncoef = 2 * K;   // halfcomplex entries 0 .. 2c+1
  const int nwarp = PT / 32, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nseg = (ncoef >= nwarp) ? 1 : nwarp / ncoef;
  const int seglen = (n + nseg - 1) / nseg;
  for (int item = warp; item < ncoef * nseg; item += nwarp) {
    const int m = item / nseg, q = item - m * nseg;
    const int lo = q * seglen, hi = min(n, lo + seglen);
    const double *__restrict__ b = basis + (size_t)m * n;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int i = lo + lane; i < hi; i += 128) {
      a0 += x[i] * __ldg(b + i);
      if (i + 32 < hi) a1 += x[i + 32] * __ldg(b + i + 32);
      if (i + 64 < hi) a2 += x[i + 64] * __ldg(b + i + 64);
      if (i + 96 < hi) a3 += x[i + 96] * __ldg(b + i + 96);
    }
    const double r = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) part[item] = r;
  }
  __syncthreads();
  if ((int)threadIdx.x < ncoef) {
    const int m = threadIdx.x;
    double r = 0.0;
    for (int q = 0; q < nseg; q++) r += part[m * nseg + q];
    const double sc = (m == 0) ? 1.0 / n : 2.0 / n;  // rfftf1.f:87-107 normalisation
    coef[m] = r * sc;
  }
  __syncthreads();
  for (int i0 = threadIdx.x; i0 < n; i0 += PT * PB) {
    double y[PB];
#pragma unroll
    for (int q = 0; q < PB; q++) y[q] = 0.0;
    for (int m = 0; m < ncoef; m++) {
      const double cm = coef[m];
      const double *__restrict__ b = basis + (size_t)m * n;
#pragma unroll
      for (int q = 0; q < PB; q++) {
        const int i = i0 + q * PT;
        if (i < n) y[q] += cm * __ldg(b + i);
      }
    }
#pragma unroll
    for (int q = 0; q < PB; q++) {
      const int i = i0 + q * PT;
      if (i < n) x[i] = y[q];
    }
  }
  __syncthreads();
}

// Moving reduced tendency of one row (specified extension): x'(i) = sum_{|d| < r} (r - |d|) x(i + d)
// / r^2, periodic -- the average over the r offsets of the zonally reduced grid of the cell means handed back to the
// fine cells.  Same summation order as the oracle (d ascending).  Rows of up to PQ * PT elements.
__device__ inline void reduce_row_dev(double *x, int n, int r) {
  double y[PQ];
#pragma unroll
  for (int c = 0; c < PQ; c++) {
    const int i = threadIdx.x + c * PT;
    y[c] = 0.0;
    if (i < n) {
      double acc = 0.0;
      for (int dd = -(r - 1); dd <= r - 1; dd++) {
        int ii = i + dd;
        if (ii < 0) ii += n;
        if (ii >= n) ii -= n;
        acc = acc + (double)(r - (dd < 0 ? -dd : dd)) * x[ii];
      }
      y[c] = acc / (double)(r * r);
    }
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < PQ; c++) {
    const int i = threadIdx.x + c * PT;
    if (i < n) x[i] = y[c];
  }
  __syncthreads();
}

// Latency is what this kernel is about (it sits between two stage launches, and at one CTA per SM it is a chain
// of dependent phases, not a throughput problem):
//  * the work item comes from the kernel parameters (no dependent global load at the head of the chain);
//  * EVERY global load of a row -- tendency, weight (+ghs), base-state / previous-tendency row, basis -- is issued
//    before the first use, so the CTA pays one memory round trip (2.5 us behind the stage kernel's write-back);
//  * for the cutoffs the reference ships (<= KF-1) the projector runs on THREAD-OWNED elements: thread t owns
//    i = b + q n/4 (q = 0..3) for b = t + g PT.  The basis at b + q n/4 is the basis at b turned by k q 90 degrees --
//    sign changes and swaps -- so the four elements are folded with additions first and one (cos, sin) pair per
//    wavenumber serves all four; the pair at b = t + g PT is the pair at t turned by a host-tabulated angle.  The
//    inner product s1 and all 2K+1 dot products come out of ONE interleaved block reduction, the reconstruction
//    needs no loads at all.
template <int NV>
__device__ __forceinline__ void warp_sum_n(double (&v)[NV]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int m = 0; m < NV; m++) v[m] += __shfl_xor_sync(0xffffffffu, v[m], o);
  }
}
// deterministic block sum of two values at once; results valid in thread 0; red must hold >= 64 doubles
template <int NT>
__device__ __forceinline__ void block_sum2(double &a, double &b, double *red) {
  double v[2] = {a, b};
  warp_sum_n<2>(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) {
    red[w] = v[0];
    red[32 + w] = v[1];
  }
  __syncthreads();
  if (w == 0) {
    v[0] = (l < NT / 32) ? red[l] : 0.0;
    v[1] = (l < NT / 32) ? red[32 + l] : 0.0;
    warp_sum_n<2>(v);
  }
  a = v[0];
  b = v[1];
}

template <int MODE>
__global__ void __launch_bounds__(PT) k_polar(const __grid_constant__ PolarArgs a) {
  extern __shared__ double psm[];  // x[n], w[n] (, q[n])
  __shared__ double red[64];
  __shared__ double coef[512];
  __shared__ double part[512];
  __shared__ double rot_s[PQ * KF * 2];
  __shared__ double bc[2];
  constexpr int NW = PT / 32;
  constexpr int NV = 2 + 2 * KF;  // s1, <x,1>, <x,cos k>, <x,sin k>
  static_assert(NV * NW <= 512 && (NV * NW) % 32 == 0, "second reduction stage runs on whole warps");
  const unsigned pk = a.items[blockIdx.x];
  const int j = (int)(pk & 0xffffu), cutoff = (int)((pk >> 16) & 0x1ffu) - 1, kind = (int)((pk >> 28) & 7u);
  const bool isred = (pk & ITEM_REDUCE) != 0;   // a reduced row: cutoff + 1 is the reduction factor
  const int rescale = isred ? (a.rescale && a.reduce_smooth) : a.rescale;
  const int n = a.g.nlon, r0 = a.g.r0;
  const int tid = threadIdx.x;
  trace_in(a.tseq);
  double *x = psm, *w = psm + n, *qs = psm + 2 * (size_t)n;
  const ptrdiff_t off = (ptrdiff_t)(j - r0) * (ptrdiff_t)n;
  const bool useq = (MODE != MODE_EVAL) && a.use_q;
  double ip1 = 0.0, ip2 = 0.0;
  if (a.pdl) {
    pdl_trigger();
    pdl_wait();
  }

  if (kind == IT_POLE_S || kind == IT_POLE_N) {
    // src/dycore_mod.F90:572-596: zonal sum of the single adjacent flux, broadcast along the pole row
    const double *__restrict__ g0 = a.Egd + off;
    const double *__restrict__ g1 = (kind == IT_POLE_S) ? a.Egd + off + n : a.Egd + off - n;
    const double *__restrict__ vv = (kind == IT_POLE_S) ? a.EV + off : a.EV + off - n;
    const double *__restrict__ Q = (MODE == MODE_S3A) ? a.Pgd : a.Ogd;
    double acc = 0.0;
    for (int i0 = tid; i0 < n; i0 += PT * 8) {
      double av[8], bv[8], cv[8], qv[8];
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int i = i0 + e * PT;
        const bool ok = i < n;
        av[e] = ok ? __ldg(g0 + i) : 0.0;
        bv[e] = ok ? __ldg(g1 + i) : 0.0;
        cv[e] = ok ? __ldg(vv + i) : 0.0;
        qv[e] = (ok && useq) ? __ldg(Q + off + i) : 0.0;
      }
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int i = i0 + e * PT;
        if (i < n) {
          const double f = (sqrt(av[e]) + sqrt(bv[e])) * cv[e];
          acc = (kind == IT_POLE_S) ? acc + f : acc - f;
          if (useq) qs[i] = qv[e];
        }
      }
    }
    const double r = block_sum<PT>(acc, red);
    if (tid == 0) bc[0] = -(r * 2.0 / n / a.radius / a.dlat);  // dgd = -mass_div_lon(=0) - mass_div_lat
    __syncthreads();
    const double dG = bc[0];
    const double cw = a.t.cosf[j];
    for (int i = tid; i < n; i += PT) {
      double o = 0.0;
      if (MODE != MODE_EVAL) o = useq ? qs[i] : __ldg(Q + off + i);
      if (MODE == MODE_S1 || MODE == MODE_S2) a.Ngd[off + i] = o + a.dt * dG;
      if (MODE != MODE_S1) a.Tgd[off + i] = dG;
      if (MODE == MODE_S3A) {
        ip1 = ip1 + dG * o * cw;
        ip2 = ip2 + dG * dG * cw;
      }
    }
  } else {
    double *T = (kind == IT_DU) ? a.TU : (kind == IT_DV) ? a.TV : a.Tgd;
    const double *__restrict__ W = (kind == IT_DU) ? a.EU : (kind == IT_DV) ? a.EV : a.Egd;
    const double *__restrict__ O = (kind == IT_DU) ? a.OU : (kind == IT_DV) ? a.OV : a.Ogd;
    double *N = (kind == IT_DU) ? a.NU : (kind == IT_DV) ? a.NV : a.Ngd;
    const double *__restrict__ P = (kind == IT_DU) ? a.PU : (kind == IT_DV) ? a.PV : a.Pgd;
    const double *__restrict__ Q = (MODE == MODE_S1 || MODE == MODE_S2) ? O : (MODE == MODE_S3A) ? P : nullptr;
    const int K = cutoff + 1;
    const int n4 = n >> 2;
    const int G = (n4 + PT - 1) / PT;  // element groups per thread on the fast path
    const bool fast = !isred && (K >= 1) && (K <= KF) && ((n & 3) == 0) && (G <= PQ) && (2 * K < n);
    // element e of a batch of 8: fast path: group g0 + e/4, quarter e%4; general path: i0 + e PT
    auto elem = [&](int i0, int e) -> int {
      if (fast) {
        const int b = tid + (i0 + (e >> 2)) * PT;
        return (b < n4) ? b + (e & 3) * n4 : n;
      }
      return min(i0 + e * PT, n);
    };
    double bcv[KF], bsv[KF];
#pragma unroll
    for (int k = 0; k < KF; k++) {
      bcv[k] = bsv[k] = 0.0;
      if (fast && k < K && tid < n4) {
        bcv[k] = __ldg(a.basis + (size_t)(2 * k + 1) * n + tid);
        bsv[k] = __ldg(a.basis + (size_t)(2 * k + 2) * n + tid);
      }
    }
    double rv = 0.0;
    if (fast && tid < PQ * KF * 2) rv = __ldg(a.rot + tid);
    double s1p = 0.0;
    const int nb = fast ? G : n;           // loop bound / step of the batch loop in its own units
    const int step = fast ? 2 : PT * 8;
    for (int i0 = fast ? 0 : tid; i0 < nb; i0 += step) {
      double xv[8], wv[8], gv[8], qv[8];
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int i = elem(i0, e);
        const bool ok = i < n;
        xv[e] = ok ? T[off + i] : 0.0;
        wv[e] = (ok && rescale) ? __ldg(W + off + i) : 0.0;
        gv[e] = (ok && rescale && kind == IT_DGD) ? __ldg(a.ghs + off + i) : 0.0;
        qv[e] = (ok && useq) ? __ldg(Q + off + i) : 0.0;
      }
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int i = elem(i0, e);
        if (i < n) {
          const double ww = wv[e] + gv[e];
          x[i] = xv[e];
          w[i] = ww;
          if (useq) qs[i] = qv[e];
          s1p = s1p + xv[e] * ww;
        }
      }
    }
    if (fast && tid < PQ * KF * 2) rot_s[tid] = rv;
    __syncthreads();
    bool do_filter = true;
    double s1 = 0.0, s2 = 1.0;
    if (fast) {
      // ---- s1 and the 2K+1 dot products in one pass over the thread's own elements ---------------------------
      double v[NV];
#pragma unroll
      for (int m = 0; m < NV; m++) v[m] = 0.0;
      v[0] = s1p;
      for (int g = 0; g < G; g++) {
        const int b = tid + g * PT;
        if (b < n4) {
          const double x0 = x[b], x1 = x[b + n4], x2 = x[b + 2 * n4], x3 = x[b + 3 * n4];
          const double e0 = x0 + x2, e1 = x1 + x3, d0 = x0 - x2, d1 = x1 - x3;
          const double a0 = e0 + e1, a2 = e0 - e1;
          v[1] += a0;
          const double *__restrict__ e = rot_s + g * (KF * 2);
#pragma unroll
          for (int k = 0; k < KF; k++) {
            if (k < K) {
              const double ec = e[2 * k], es = e[2 * k + 1];
              const double cr = bcv[k] * ec - bsv[k] * es;   // cos, sin of wavenumber k+1 at element b
              const double sr = bsv[k] * ec + bcv[k] * es;
              switch ((k + 1) & 3) {                         // quarter turns of the other three elements
                case 1: v[2 + 2 * k] += cr * d0 - sr * d1; v[3 + 2 * k] += sr * d0 + cr * d1; break;
                case 2: v[2 + 2 * k] += cr * a2;           v[3 + 2 * k] += sr * a2;           break;
                case 3: v[2 + 2 * k] += cr * d0 + sr * d1; v[3 + 2 * k] += sr * d0 - cr * d1; break;
                default: v[2 + 2 * k] += cr * a0;          v[3 + 2 * k] += sr * a0;           break;
              }
            }
          }
        }
      }
      const int warp = tid >> 5, lane = tid & 31;
      warp_sum_n<NV>(v);
      if (lane == 0) {
#pragma unroll
        for (int m = 0; m < NV; m++) part[m * NW + warp] = v[m];
      }
      __syncthreads();
      if (tid < NV * NW) {  // NW-lane segments, fixed tree
        double p = part[tid];
#pragma unroll
        for (int o = NW / 2; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        if ((tid & (NW - 1)) == 0) {
          const int m = tid / NW;
          if (m == 0) bc[0] = p;
          else coef[m - 1] = p * ((m == 1) ? 1.0 / n : 2.0 / n);  // rfftf1.f:87-107 normalisation
        }
      }
      __syncthreads();
      s1 = bc[0];
      if (rescale) do_filter = fabs(s1) > 1.0e-16;  // filter_inner_product_threshold, src/filter_mod.F90:31
      if (do_filter) {
        // ---- reconstruction from entries 0 .. 2K-1 (sin(K x) is dropped: quirk B3) ---------------------------
        const double c0 = coef[0];
        double cc[KF], cs[KF];
#pragma unroll
        for (int k = 0; k < KF; k++) {
          cc[k] = (k < K) ? coef[1 + 2 * k] : 0.0;
          cs[k] = (k < K - 1) ? coef[2 + 2 * k] : 0.0;
        }
        double s2p = 0.0;
        for (int g = 0; g < G; g++) {
          const int b = tid + g * PT;
          if (b < n4) {
            const double *__restrict__ e = rot_s + g * (KF * 2);
            double S0 = c0, Pa = 0.0, Ra = 0.0, P2 = 0.0;
#pragma unroll
            for (int k = 0; k < KF; k++) {
              if (k < K) {
                const double ec = e[2 * k], es = e[2 * k + 1];
                const double cr = bcv[k] * ec - bsv[k] * es;
                const double sr = bsv[k] * ec + bcv[k] * es;
                const double Pk = cc[k] * cr + cs[k] * sr;   // value at b; a quarter turn further: Rk, -Pk, -Rk
                const double Rk = cs[k] * cr - cc[k] * sr;
                switch ((k + 1) & 3) {
                  case 1: Pa += Pk; Ra += Rk; break;
                  case 2: P2 += Pk; break;
                  case 3: Pa += Pk; Ra -= Rk; break;
                  default: S0 += Pk; break;
                }
              }
            }
            const double y0 = (S0 + P2) + Pa, y2 = (S0 + P2) - Pa, y1 = (S0 - P2) + Ra, y3 = (S0 - P2) - Ra;
            x[b] = y0;
            x[b + n4] = y1;
            x[b + 2 * n4] = y2;
            x[b + 3 * n4] = y3;
            s2p = s2p + y0 * w[b] + y1 * w[b + n4] + y2 * w[b + 2 * n4] + y3 * w[b + 3 * n4];
          }
        }
        if (rescale) {
          const double r = block_sum<PT>(s2p, red);
          if (tid == 0) bc[1] = r;
          __syncthreads();
          s2 = bc[1];
        }
      }
    } else {
      if (rescale) {
        const double r = block_sum<PT>(s1p, red);
        if (tid == 0) bc[0] = r;
        __syncthreads();
        s1 = bc[0];
        do_filter = fabs(s1) > 1.0e-16;
      }
      if (do_filter) {
        if (isred) reduce_row_dev(x, n, cutoff + 1);
        else project_row(x, n, cutoff, a.basis, coef, part);
        if (rescale) {
          double s2p = 0.0;
          for (int i = tid; i < n; i += PT) s2p = s2p + x[i] * w[i];
          const double r = block_sum<PT>(s2p, red);
          if (tid == 0) bc[1] = r;
          __syncthreads();
          s2 = bc[1];
        }
      }
    }
    const double cw = (kind == IT_DV) ? a.t.cosh[j] : a.t.cosf[j];
    // src/dycore_mod.F90:218.  s2 == 0 exactly (a row whose filtered inner product cancels to the last bit; the
    // reference would divide by zero and abort with NaN, quirk B13): the filtered row is left unscaled.
    const bool scale = do_filter && rescale && (s2 != 0.0);
#if !GMD_STRICT
    const double ratio = scale ? s1 / s2 : 1.0;
#endif
    // own elements again on the fast path (x and qs of element i were written by this thread)
    for (int i0 = fast ? 0 : tid; i0 < nb; i0 += step) {
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int i = elem(i0, e);
        if (i < n) {
          double o = 0.0;
          if (MODE != MODE_EVAL) o = useq ? qs[i] : __ldg(Q + off + i);
          double d = x[i];
#if GMD_STRICT
          if (scale) d = d * s1 / s2;
#else
          d = d * ratio;
#endif
          if (MODE == MODE_S1 || MODE == MODE_S2) N[off + i] = o + a.dt * d;
          if (MODE != MODE_S1) T[off + i] = d;
          if (MODE == MODE_S3A) {
            ip1 = ip1 + d * o * cw;
            ip2 = ip2 + d * d * cw;
          }
        }
      }
    }
  }
  if (MODE == MODE_S3A) {
    block_sum2<PT>(ip1, ip2, red);
    if (tid == 0) {
      a.partials[2 * blockIdx.x] = ip1;
      a.partials[2 * blockIdx.x + 1] = ip2;
    }
    if (a.fold.ticket)
      fold_tail<PT>(a.fold.ticket, a.fold.total, a.fold_partials, a.fold.n, a.fold.out, a.fold.r.page, a.fold.r.rank,
                    a.fold.r.nranks, a.fold.r.k, a.tseq);
  }
  trace_out(a.tseq);
}

// ---------------------------------------------------------------------------------------------------------
// k_polar_lean<MODE, G>: the fast projector of k_polar with everything but the reductions in REGISTERS.
// k_polar is one kernel for every kind of item (any cutoff, any row length, reduced rows); that generality costs
// instructions -- 1860 per warp and item at 3600 columns, 17 % of them fp64 (profiles/r1_k_ncu_polar_summary.txt) -- and the
// kernel sits on the critical chain sweep -> polar rows -> sweep of every band that holds polar rows.  When every item
// of a launch is one the thread-owned projector handles (filter rows with cutoff < KF on a row of n % 4 == 0 elements,
// G = ceil(n / 4 / PT) <= 2 element groups per thread; pole caps) the host launches this variant instead: the G x 4
// elements a thread owns, their weights and base values stay in registers from the first load to the last store (no
// row in shared memory, no index arithmetic beyond b + q n/4), the loops over groups and quarters are unrolled at
// compile time.  Same element-to-thread mapping, same operation order, same reduction trees as k_polar: the results are
// bit-identical (tests/test_gpu_parity.py::test_polar_lean_is_bit_identical).
// ---------------------------------------------------------------------------------------------------------
template <int MODE, int G>
__global__ void __launch_bounds__(PT) k_polar_lean(const __grid_constant__ PolarArgs a) {
  __shared__ double red[64];
  __shared__ double coef[2 * KF + 2];
  __shared__ double part[512];
  __shared__ double bc[2];
  constexpr int NW = PT / 32;
  constexpr int NV = 2 + 2 * KF;
  static_assert(NV * NW <= 512 && (NV * NW) % 32 == 0, "second reduction stage runs on whole warps");
  const unsigned pk = a.items[blockIdx.x];
  const int j = (int)(pk & 0xffffu), cutoff = (int)((pk >> 16) & 0x1ffu) - 1, kind = (int)((pk >> 28) & 7u);
  const int rescale = a.rescale;
  const int n = a.g.nlon, r0 = a.g.r0;
  const int tid = threadIdx.x;
  trace_in(a.tseq);
  const ptrdiff_t off = (ptrdiff_t)(j - r0) * (ptrdiff_t)n;
  double ip1 = 0.0, ip2 = 0.0;
  if (a.pdl) {
    pdl_trigger();
    pdl_wait();
  }
  if (kind == IT_POLE_S || kind == IT_POLE_N) {
    // src/dycore_mod.F90:572-596 (as k_polar)
    const double *__restrict__ g0 = a.Egd + off;
    const double *__restrict__ g1 = (kind == IT_POLE_S) ? a.Egd + off + n : a.Egd + off - n;
    const double *__restrict__ vv = (kind == IT_POLE_S) ? a.EV + off : a.EV + off - n;
    const double *__restrict__ Q = (MODE == MODE_S3A) ? a.Pgd : a.Ogd;
    double acc = 0.0;
    for (int i0 = tid; i0 < n; i0 += PT * 8) {
      double av[8], bv[8], cv[8];
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int i = i0 + e * PT;
        const bool ok = i < n;
        av[e] = ok ? __ldg(g0 + i) : 0.0;
        bv[e] = ok ? __ldg(g1 + i) : 0.0;
        cv[e] = ok ? __ldg(vv + i) : 0.0;
      }
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int i = i0 + e * PT;
        if (i < n) {
          const double f = (sqrt(av[e]) + sqrt(bv[e])) * cv[e];
          acc = (kind == IT_POLE_S) ? acc + f : acc - f;
        }
      }
    }
    const double r = block_sum<PT>(acc, red);
    if (tid == 0) bc[0] = -(r * 2.0 / n / a.radius / a.dlat);
    __syncthreads();
    const double dG = bc[0];
    const double cw = a.t.cosf[j];
    for (int i = tid; i < n; i += PT) {
      double o = 0.0;
      if (MODE != MODE_EVAL) o = __ldg(Q + off + i);
      if (MODE == MODE_S1 || MODE == MODE_S2) a.Ngd[off + i] = o + a.dt * dG;
      if (MODE != MODE_S1) a.Tgd[off + i] = dG;
      if (MODE == MODE_S3A) {
        ip1 = ip1 + dG * o * cw;
        ip2 = ip2 + dG * dG * cw;
      }
    }
  } else {
    double *T = (kind == IT_DU) ? a.TU : (kind == IT_DV) ? a.TV : a.Tgd;
    const double *__restrict__ W = (kind == IT_DU) ? a.EU : (kind == IT_DV) ? a.EV : a.Egd;
    const double *__restrict__ O = (kind == IT_DU) ? a.OU : (kind == IT_DV) ? a.OV : a.Ogd;
    double *N = (kind == IT_DU) ? a.NU : (kind == IT_DV) ? a.NV : a.Ngd;
    const double *__restrict__ P = (kind == IT_DU) ? a.PU : (kind == IT_DV) ? a.PV : a.Pgd;
    const double *__restrict__ Q = (MODE == MODE_S1 || MODE == MODE_S2) ? O : (MODE == MODE_S3A) ? P : nullptr;
    const int K = cutoff + 1;
    const int n4 = n >> 2;
    // ---- every global load of the row first: element (g, q) is b + q n4, b = tid + g PT -----------------------
    double xv[G][4], wv[G][4], qv[G][4];
    bool okg[G];
#pragma unroll
    for (int g = 0; g < G; g++) {
      const int b = tid + g * PT;
      okg[g] = b < n4;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const ptrdiff_t i = off + b + q * n4;
        xv[g][q] = okg[g] ? T[i] : 0.0;
        double ww = (okg[g] && rescale) ? __ldg(W + i) : 0.0;
        if (kind == IT_DGD) ww += (okg[g] && rescale) ? __ldg(a.ghs + i) : 0.0;
        wv[g][q] = ww;
        qv[g][q] = (okg[g] && MODE != MODE_EVAL) ? __ldg(Q + i) : 0.0;
      }
    }
    double bcv[KF], bsv[KF];
#pragma unroll
    for (int k = 0; k < KF; k++) {
      bcv[k] = bsv[k] = 0.0;
      if (k < K && tid < n4) {
        bcv[k] = __ldg(a.basis + (size_t)(2 * k + 1) * n + tid);
        bsv[k] = __ldg(a.basis + (size_t)(2 * k + 2) * n + tid);
      }
    }
    // ---- s1 and the 2K+1 dot products in one pass over the thread's own elements --------------------------------
    double v[NV];
#pragma unroll
    for (int m = 0; m < NV; m++) v[m] = 0.0;
    {
      double s1p = 0.0;
#pragma unroll
      for (int g = 0; g < G; g++)
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (okg[g]) s1p = s1p + xv[g][q] * wv[g][q];
      v[0] = s1p;
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
      if (okg[g]) {
        const double x0 = xv[g][0], x1 = xv[g][1], x2 = xv[g][2], x3 = xv[g][3];
        const double e0 = x0 + x2, e1 = x1 + x3, d0 = x0 - x2, d1 = x1 - x3;
        const double a0 = e0 + e1, a2 = e0 - e1;
        v[1] += a0;
#pragma unroll
        for (int k = 0; k < KF; k++) {
          if (k < K) {
            const double ec = __ldg(a.rot + (g * KF + k) * 2), es = __ldg(a.rot + (g * KF + k) * 2 + 1);
            const double cr = bcv[k] * ec - bsv[k] * es;   // cos, sin of wavenumber k+1 at element b
            const double sr = bsv[k] * ec + bcv[k] * es;
            switch ((k + 1) & 3) {                         // quarter turns of the other three elements
              case 1: v[2 + 2 * k] += cr * d0 - sr * d1; v[3 + 2 * k] += sr * d0 + cr * d1; break;
              case 2: v[2 + 2 * k] += cr * a2;           v[3 + 2 * k] += sr * a2;           break;
              case 3: v[2 + 2 * k] += cr * d0 + sr * d1; v[3 + 2 * k] += sr * d0 - cr * d1; break;
              default: v[2 + 2 * k] += cr * a0;          v[3 + 2 * k] += sr * a0;           break;
            }
          }
        }
      }
    }
    const int warp = tid >> 5, lane = tid & 31;
    warp_sum_n<NV>(v);
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < NV; m++) part[m * NW + warp] = v[m];
    }
    __syncthreads();
    if (tid < NV * NW) {  // NW-lane segments, fixed tree
      double p = part[tid];
#pragma unroll
      for (int o = NW / 2; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      if ((tid & (NW - 1)) == 0) {
        const int m = tid / NW;
        if (m == 0) bc[0] = p;
        else coef[m - 1] = p * ((m == 1) ? 1.0 / n : 2.0 / n);  // rfftf1.f:87-107 normalisation
      }
    }
    __syncthreads();
    const double s1 = bc[0];
    bool do_filter = true;
    double s2 = 1.0;
    if (rescale) do_filter = fabs(s1) > 1.0e-16;  // filter_inner_product_threshold, src/filter_mod.F90:31
    if (do_filter) {
      // ---- reconstruction from entries 0 .. 2K-1 (sin(K x) is dropped: quirk B3), in place in the registers -----
      const double c0 = coef[0];
      double cc[KF], cs[KF];
#pragma unroll
      for (int k = 0; k < KF; k++) {
        cc[k] = (k < K) ? coef[1 + 2 * k] : 0.0;
        cs[k] = (k < K - 1) ? coef[2 + 2 * k] : 0.0;
      }
      double s2p = 0.0;
#pragma unroll
      for (int g = 0; g < G; g++) {
        if (okg[g]) {
          double S0 = c0, Pa = 0.0, Ra = 0.0, P2 = 0.0;
#pragma unroll
          for (int k = 0; k < KF; k++) {
            if (k < K) {
              const double ec = __ldg(a.rot + (g * KF + k) * 2), es = __ldg(a.rot + (g * KF + k) * 2 + 1);
              const double cr = bcv[k] * ec - bsv[k] * es;
              const double sr = bsv[k] * ec + bcv[k] * es;
              const double Pk = cc[k] * cr + cs[k] * sr;   // value at b; a quarter turn further: Rk, -Pk, -Rk
              const double Rk = cs[k] * cr - cc[k] * sr;
              switch ((k + 1) & 3) {
                case 1: Pa += Pk; Ra += Rk; break;
                case 2: P2 += Pk; break;
                case 3: Pa += Pk; Ra -= Rk; break;
                default: S0 += Pk; break;
              }
            }
          }
          const double y0 = (S0 + P2) + Pa, y2 = (S0 + P2) - Pa, y1 = (S0 - P2) + Ra, y3 = (S0 - P2) - Ra;
          xv[g][0] = y0;
          xv[g][1] = y1;
          xv[g][2] = y2;
          xv[g][3] = y3;
          s2p = s2p + y0 * wv[g][0] + y1 * wv[g][1] + y2 * wv[g][2] + y3 * wv[g][3];
        }
      }
      if (rescale) {
        const double r = block_sum<PT>(s2p, red);
        if (tid == 0) bc[1] = r;
        __syncthreads();
        s2 = bc[1];
      }
    }
    const double cw = (kind == IT_DV) ? a.t.cosh[j] : a.t.cosf[j];
    const bool scale = do_filter && rescale && (s2 != 0.0);   // s2 == 0: see k_polar
    const double ratio = scale ? s1 / s2 : 1.0;
#pragma unroll
    for (int g = 0; g < G; g++) {
      if (okg[g]) {
        const int b = tid + g * PT;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const ptrdiff_t i = off + b + q * n4;
          const double o = qv[g][q];
          const double d = xv[g][q] * ratio;
          if (MODE == MODE_S1 || MODE == MODE_S2) N[i] = o + a.dt * d;
          if (MODE != MODE_S1) T[i] = d;
          if (MODE == MODE_S3A) {
            ip1 = ip1 + d * o * cw;
            ip2 = ip2 + d * d * cw;
          }
        }
      }
    }
  }
  if (MODE == MODE_S3A) {
    block_sum2<PT>(ip1, ip2, red);
    if (tid == 0) {
      a.partials[2 * blockIdx.x] = ip1;
      a.partials[2 * blockIdx.x + 1] = ip2;
    }
    if (a.fold.ticket)
      fold_tail<PT>(a.fold.ticket, a.fold.total, a.fold_partials, a.fold.n, a.fold.out, a.fold.r.page, a.fold.r.rank,
                    a.fold.r.nranks, a.fold.r.k, a.tseq);
  }
  trace_out(a.tseq);
}

// ---------------------------------------------------------------------------------------------------------
// Fused polar cap (k_cap): the sweep over the rows next to a pole AND the polar rows of the same sweep -- filter +
// rescale + update of the flagged rows, pole caps -- in ONE launch.  On a short latitude band the chain
//   cap sweep -> polar rows -> cap sweep of the next operator evaluation -> ...
// is what a polar rank waits for (36 links per model step with csp2 x 10): as two launches a link costs
// 7-10 us + 10-13 us + two launch gaps (profiles/r2_d_timeline_*), and the 512-thread, one-per-SM CTAs of k_polar cannot
// co-run with the interior sweep.  Here
//  * CTAs [0, n_march) run stage_body on the cap rows (same code as k_stage), then every CTA of the launch meets at
//    a grid barrier (arrive / depart counters in global memory, self-resetting, so a captured graph replays it);
//    the launch is sized to be co-resident and is given the highest launch priority;
//  * the polar items are then dealt to CLUSTERS of CL = 4 CTAs x 128 threads: the PT = 512 threads of one item are
//    the same thread-owned-element projector as k_polar's fast path (element b + q n/4 of thread b; one (cos, sin)
//    pair per wavenumber serves four elements), but every row element stays in REGISTERS -- no row in shared memory,
//    so a cap CTA occupies one ordinary stage-CTA slot -- and the block reductions become cluster reductions over
//    distributed shared memory (fixed order: deterministic, and the same bits in each CTA of the cluster).
// Only the cutoffs / row lengths of the fast projector are supported (the host falls back to k_stage + k_polar).
// ---------------------------------------------------------------------------------------------------------
constexpr int CL = 4;   // CTAs per cluster
static_assert(CL * BX == PT, "the rotation table of the fast projector is built for PT threads per item");
constexpr int GR = 2;   // element groups of a thread held in registers (rows of up to 4 GR PT elements; longer: reloaded)

struct CapArgs {
  unsigned *bar;     // {arrived, departed}
  int n_march;       // CTAs [0, n_march) run the sweep: linear index = ((bz - z0) * gy + by) * gx + bx
  int gx, gy, z0;    // z0: first row range of StageArgs this launch covers (a band with one pole has one range)
  int nitems;        // polar items, dealt round-robin to the launch's clusters
};
#if GMD_TRACE
// phase stamps of CTA 0 of the last k_cap launches (slot = timeline slot % 64): start, sweep done, barrier passed,
// first item done, end
__device__ u64 g_capdbg[64 * 12];
#define GMD_CAP_STAMP(k) do { if (blockIdx.x == 0 && threadIdx.x == 0 && a.tseq > 0) g_capdbg[(a.tseq % 64) * 12 + (k)] = gtimer(); } while (0)
#else
#define GMD_CAP_STAMP(k) do { } while (0)
#endif

__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_cluster(const double *p, unsigned rank) {
  const unsigned la = (unsigned)__cvta_generic_to_shared(p);
  unsigned ra;
  double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Sum NVAL values over the CL * BX threads of a cluster in a fixed order: warp shuffles, one shared-memory slot per
// warp, then every CTA adds the CL * SW slots of the cluster in (rank, warp) order.  `part` ([NVAL][SW]) must not be
// written again before the cluster has passed another barrier; the result is in out[0 .. NVAL) for every thread.
template <int NVAL>
__device__ __forceinline__ void cluster_sum(double (&v)[NVAL], double *part, double *out) {
  warp_sum_n<NVAL>(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int m = 0; m < NVAL; m++) part[m * SW + warp] = v[m];
  }
  cluster_sync_all();
  if ((int)threadIdx.x < NVAL) {
    double q[CL * SW];
#pragma unroll
    for (int r = 0; r < CL; r++)
#pragma unroll
      for (int w = 0; w < SW; w++) q[r * SW + w] = ld_cluster(part + threadIdx.x * SW + w, (unsigned)r);
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < CL * SW; k++) sum += q[k];
    out[threadIdx.x] = sum;
  }
  __syncthreads();
}

struct CapSmem {
  double partA[(2 + 2 * KF) * SW], outA[2 + 2 * KF];
  double partB[SW], outB[1];
  double partC[2 * SW], outC[2];
  double rot[PQ * KF * 2];
};

// one polar item worked on by the PT threads of a cluster; gtid = thread index inside the cluster
template <int MODE>
__device__ __noinline__ void cap_item(const PolarArgs &a, const int it, const int gtid, CapSmem &sm) {
  const unsigned pk = a.items[it];
  const int j = (int)(pk & 0xffffu), cutoff = (int)((pk >> 16) & 0x1ffu) - 1, kind = (int)(pk >> 28);
  const int n = a.g.nlon, r0 = a.g.r0;
  const ptrdiff_t off = (ptrdiff_t)(j - r0) * (ptrdiff_t)n;
  double ip1 = 0.0, ip2 = 0.0;

  if (kind == IT_POLE_S || kind == IT_POLE_N) {
    // src/dycore_mod.F90:572-596: zonal sum of the single adjacent flux, broadcast along the pole row
    const double *g0 = a.Egd + off;
    const double *g1 = (kind == IT_POLE_S) ? a.Egd + off + n : a.Egd + off - n;
    const double *vv = (kind == IT_POLE_S) ? a.EV + off : a.EV + off - n;
    const double *Q = (MODE == MODE_S3A) ? a.Pgd : a.Ogd;
    double acc[1] = {0.0};
    for (int i0 = gtid; i0 < n; i0 += PT * 8) {
      double av[8], bv[8], cv[8];
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int i = i0 + e * PT;
        const bool ok = i < n;
        av[e] = ok ? __ldcg(g0 + i) : 0.0;
        bv[e] = ok ? __ldcg(g1 + i) : 0.0;
        cv[e] = ok ? __ldcg(vv + i) : 0.0;
      }
#pragma unroll
      for (int e = 0; e < 8; e++) {
        if (i0 + e * PT < n) {
          const double f = (sqrt(av[e]) + sqrt(bv[e])) * cv[e];
          acc[0] = (kind == IT_POLE_S) ? acc[0] + f : acc[0] - f;
        }
      }
    }
    cluster_sum<1>(acc, sm.partB, sm.outB);
    const double dG = -(sm.outB[0] * 2.0 / n / a.radius / a.dlat);  // dgd = -mass_div_lon(=0) - mass_div_lat
    const double cw = a.t.cosf[j];
    for (int i = gtid; i < n; i += PT) {
      double o = 0.0;
      if (MODE != MODE_EVAL) o = __ldcg(Q + off + i);
      if (MODE == MODE_S1 || MODE == MODE_S2) a.Ngd[off + i] = o + a.dt * dG;
      if (MODE != MODE_S1) a.Tgd[off + i] = dG;
      if (MODE == MODE_S3A) {
        ip1 = ip1 + dG * o * cw;
        ip2 = ip2 + dG * dG * cw;
      }
    }
  } else {
    double *T = (kind == IT_DU) ? a.TU : (kind == IT_DV) ? a.TV : a.Tgd;
    const double *W = (kind == IT_DU) ? a.EU : (kind == IT_DV) ? a.EV : a.Egd;
    const double *O = (kind == IT_DU) ? a.OU : (kind == IT_DV) ? a.OV : a.Ogd;
    double *N = (kind == IT_DU) ? a.NU : (kind == IT_DV) ? a.NV : a.Ngd;
    const double *P = (kind == IT_DU) ? a.PU : (kind == IT_DV) ? a.PV : a.Pgd;
    const double *Q = (MODE == MODE_S1 || MODE == MODE_S2) ? O : (MODE == MODE_S3A) ? P : nullptr;
    const bool isG = (kind == IT_DGD);
    const int K = cutoff + 1;
    const int n4 = n >> 2;
    const int G = (n4 + PT - 1) / PT;
    constexpr int NV = 2 + 2 * KF;
    double bcv[KF], bsv[KF];
#pragma unroll
    for (int k = 0; k < KF; k++) {
      bcv[k] = bsv[k] = 0.0;
      if (k < K && gtid < n4) {
        bcv[k] = __ldg(a.basis + (size_t)(2 * k + 1) * n + gtid);
        bsv[k] = __ldg(a.basis + (size_t)(2 * k + 2) * n + gtid);
      }
    }
    // the (cos, sin) pairs of wavenumber k+1 at element b = gtid + g PT
    auto turn = [&](int g, int k, double &cr, double &sr) {
      const double ec = sm.rot[g * (KF * 2) + 2 * k], es = sm.rot[g * (KF * 2) + 2 * k + 1];
      cr = bcv[k] * ec - bsv[k] * es;
      sr = bsv[k] * ec + bcv[k] * es;
    };
    double v[NV];
#pragma unroll
    for (int m = 0; m < NV; m++) v[m] = 0.0;
    // s1 and the 2K+1 dot products of the four elements b + q n/4 of one group
    auto analyse = [&](int g, const double (&x)[4], const double (&w)[4]) {
#pragma unroll
      for (int q = 0; q < 4; q++) v[0] = v[0] + x[q] * w[q];
      const double e0 = x[0] + x[2], e1 = x[1] + x[3], d0 = x[0] - x[2], d1 = x[1] - x[3];
      const double a0 = e0 + e1, a2 = e0 - e1;
      v[1] += a0;
#pragma unroll
      for (int k = 0; k < KF; k++) {
        if (k < K) {
          double cr, sr;
          turn(g, k, cr, sr);
          switch ((k + 1) & 3) {  // quarter turns of the other three elements
            case 1: v[2 + 2 * k] += cr * d0 - sr * d1; v[3 + 2 * k] += sr * d0 + cr * d1; break;
            case 2: v[2 + 2 * k] += cr * a2;           v[3 + 2 * k] += sr * a2;           break;
            case 3: v[2 + 2 * k] += cr * d0 + sr * d1; v[3 + 2 * k] += sr * d0 - cr * d1; break;
            default: v[2 + 2 * k] += cr * a0;          v[3 + 2 * k] += sr * a0;           break;
          }
        }
      }
    };
    auto load4 = [&](const double *f, int b, double (&o)[4]) {
#pragma unroll
      for (int q = 0; q < 4; q++) o[q] = __ldcg(f + off + b + q * n4);
    };
    // ---- every load of the register groups first (one memory round trip), then the analysis ------------------
    double xk[GR][4], wk[GR][4], qk[GR][4];
    {
      double gk[GR][4];
#pragma unroll
      for (int g = 0; g < GR; g++) {
        const int b = gtid + g * PT;
        const bool ok = (g < G) && (b < n4);
#pragma unroll
        for (int q = 0; q < 4; q++) xk[g][q] = wk[g][q] = qk[g][q] = gk[g][q] = 0.0;
        if (ok) {
          load4(T, b, xk[g]);
          if (a.rescale) load4(W, b, wk[g]);
          if (a.rescale && isG) load4(a.ghs, b, gk[g]);
        }
      }
#pragma unroll
      for (int g = 0; g < GR; g++) {
#pragma unroll
        for (int q = 0; q < 4; q++) wk[g][q] = wk[g][q] + gk[g][q];
        if ((g < G) && (gtid + g * PT < n4)) analyse(g, xk[g], wk[g]);
      }
    }
    // the base-state / previous-tendency row travels while the cluster reduces
    if (MODE != MODE_EVAL) {
#pragma unroll
      for (int g = 0; g < GR; g++)
        if ((g < G) && (gtid + g * PT < n4)) load4(Q, gtid + g * PT, qk[g]);
    }
    for (int g = GR; g < G; g++) {   // rows longer than 4 GR PT elements
      const int b = gtid + g * PT;
      if (b < n4) {
        double x[4], w[4] = {0.0, 0.0, 0.0, 0.0}, gh[4] = {0.0, 0.0, 0.0, 0.0};
        load4(T, b, x);
        if (a.rescale) load4(W, b, w);
        if (a.rescale && isG) load4(a.ghs, b, gh);
#pragma unroll
        for (int q = 0; q < 4; q++) w[q] = w[q] + gh[q];
        analyse(g, x, w);
      }
    }
    GMD_CAP_STAMP(5);
    cluster_sum<NV>(v, sm.partA, sm.outA);
    GMD_CAP_STAMP(6);
    const double s1 = sm.outA[0];
    double s2 = 1.0;
    bool do_filter = true;
    if (a.rescale) do_filter = fabs(s1) > 1.0e-16;  // filter_inner_product_threshold, src/filter_mod.F90:31
    // ---- reconstruction from entries 0 .. 2K-1 (sin(K x) is dropped: quirk B3) -------------------------------
    const double c0 = sm.outA[1] * (1.0 / n);       // rfftf1.f:87-107 normalisation
    double cc[KF], cs[KF];
#pragma unroll
    for (int k = 0; k < KF; k++) {
      cc[k] = (k < K) ? sm.outA[2 + 2 * k] * (2.0 / n) : 0.0;
      cs[k] = (k < K - 1) ? sm.outA[3 + 2 * k] * (2.0 / n) : 0.0;
    }
    auto synth = [&](int g, double (&y)[4]) {
      double S0 = c0, Pa = 0.0, Ra = 0.0, P2 = 0.0;
#pragma unroll
      for (int k = 0; k < KF; k++) {
        if (k < K) {
          double cr, sr;
          turn(g, k, cr, sr);
          const double Pk = cc[k] * cr + cs[k] * sr;   // value at b; a quarter turn further: Rk, -Pk, -Rk
          const double Rk = cs[k] * cr - cc[k] * sr;
          switch ((k + 1) & 3) {
            case 1: Pa += Pk; Ra += Rk; break;
            case 2: P2 += Pk; break;
            case 3: Pa += Pk; Ra -= Rk; break;
            default: S0 += Pk; break;
          }
        }
      }
      y[0] = (S0 + P2) + Pa;
      y[2] = (S0 + P2) - Pa;
      y[1] = (S0 - P2) + Ra;
      y[3] = (S0 - P2) - Ra;
    };
    if (do_filter) {
      double s2p[1] = {0.0};
#pragma unroll
      for (int g = 0; g < GR; g++) {
        if ((g < G) && (gtid + g * PT < n4)) {
          synth(g, xk[g]);   // the filtered row replaces the tendency
#pragma unroll
          for (int q = 0; q < 4; q++) s2p[0] = s2p[0] + xk[g][q] * wk[g][q];
        }
      }
      for (int g = GR; g < G; g++) {
        const int b = gtid + g * PT;
        if (b < n4) {
          double y[4], w[4] = {0.0, 0.0, 0.0, 0.0}, gh[4] = {0.0, 0.0, 0.0, 0.0};
          if (a.rescale) load4(W, b, w);
          if (a.rescale && isG) load4(a.ghs, b, gh);
          synth(g, y);
#pragma unroll
          for (int q = 0; q < 4; q++) s2p[0] = s2p[0] + y[q] * (w[q] + gh[q]);
        }
      }
      GMD_CAP_STAMP(7);
      if (a.rescale) {
        cluster_sum<1>(s2p, sm.partB, sm.outB);
        s2 = sm.outB[0];
      }
      GMD_CAP_STAMP(8);
    }
    const double cw = (kind == IT_DV) ? a.t.cosh[j] : a.t.cosf[j];
    // src/dycore_mod.F90:218.  s2 == 0 exactly: the filtered row is left unscaled (see k_polar)
    const bool scale = do_filter && a.rescale && (s2 != 0.0);
#if !GMD_STRICT
    const double ratio = scale ? s1 / s2 : 1.0;
#endif
    auto finish = [&](int b, const double (&y)[4], const double (&o)[4]) {
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const ptrdiff_t e = off + b + q * n4;
        double d = y[q];
#if GMD_STRICT
        if (scale) d = d * s1 / s2;
#else
        d = d * ratio;
#endif
        if (MODE == MODE_S1 || MODE == MODE_S2) N[e] = o[q] + a.dt * d;
        if (MODE != MODE_S1) T[e] = d;
        if (MODE == MODE_S3A) {
          ip1 = ip1 + d * o[q] * cw;
          ip2 = ip2 + d * d * cw;
        }
      }
    };
#pragma unroll
    for (int g = 0; g < GR; g++) {
      const int b = gtid + g * PT;
      if ((g < G) && (b < n4)) finish(b, xk[g], qk[g]);
    }
    for (int g = GR; g < G; g++) {
      const int b = gtid + g * PT;
      if (b < n4) {
        double y[4], o[4] = {0.0, 0.0, 0.0, 0.0};
        if (MODE != MODE_EVAL) load4(Q, b, o);
        if (do_filter) synth(g, y);
        else load4(T, b, y);
        finish(b, y, o);
      }
    }
  }
  GMD_CAP_STAMP(9);
  if (MODE == MODE_S3A) {
    double w2[2] = {ip1, ip2};
    cluster_sum<2>(w2, sm.partC, sm.outC);
    if (gtid == 0) {
      a.partials[2 * it] = sm.outC[0];
      a.partials[2 * it + 1] = sm.outC[1];
    }
    if (a.fold.ticket && gtid < BX)   // the first CTA of the cluster
      fold_tail<BX>(a.fold.ticket, a.fold.total, a.fold_partials, a.fold.n, a.fold.out, a.fold.r.page, a.fold.r.rank,
                    a.fold.r.nranks, a.fold.r.k, a.tseq);
  }
}

template <int PASS, int ADV, int MODE, int LAZY>
__global__ void __launch_bounds__(BX, (stage_minb<PASS, MODE, LAZY>()))
k_cap(const StageArgs a, const __grid_constant__ PolarArgs p, const CapArgs c) {
  __shared__ double red[2 * SW];
  __shared__ CapSmem sm;
  extern __shared__ __align__(16) double srow[];
  trace_in(a.tseq);
  GMD_CAP_STAMP(0);
  const int tid = threadIdx.x, b = blockIdx.x;
  for (int k = tid; k < PQ * KF * 2; k += BX) sm.rot[k] = __ldg(p.rot + k);
  if (b < c.n_march) {
    const int bx = b % c.gx, by = (b / c.gx) % c.gy, bz = c.z0 + b / (c.gx * c.gy);
    stage_body<PASS, ADV, MODE, LAZY, false>(a, bx, by, bz, c.gx, red, srow, srow + (size_t)(a.rows_per_cta + 2) * RC_N);
  }
  GMD_CAP_STAMP(1);
  // ---- grid barrier: the polar items read rows written by any CTA of the sweep --------------------------------
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    atomicAdd(c.bar, 1u);
    while (ld_acquire_gpu_u32(c.bar) < gridDim.x) { }
    if (atomicAdd(c.bar + 1, 1u) == gridDim.x - 1) {  // everybody has seen the count: reset for the next launch
      c.bar[1] = 0u;
      c.bar[0] = 0u;
    }
  }
  __syncthreads();
  GMD_CAP_STAMP(2);
  const int gtid = (int)cluster_ctarank() * BX + tid;
  const int ncl = (int)gridDim.x / CL;
  for (int it = b / CL; it < c.nitems; it += ncl) {
    cap_item<MODE>(p, it, gtid, sm);
    cluster_sync_all();   // nobody overwrites its reduction slots or leaves while a neighbour still reads them
    if (it == b / CL) GMD_CAP_STAMP(3);
  }
  GMD_CAP_STAMP(4);
  trace_out(a.tseq);
}

__global__ void __launch_bounds__(256) k_reduce_pairs(const double *__restrict__ partials, int n, double *out,
                                                      const RedArgs r) {
  __shared__ double red[32];
  __shared__ double gath[2 * MAXR + 2];
  trace_in(r.tseq);
  reduce_pairs_body<256>(partials, n, out, r.page, r.rank, r.nranks, r.k, red, gath, r.tseq);
  trace_out(r.tseq);
}

// Halo rows of up to three fields stored straight into the neighbours' ghost rows, then one release per
// neighbour.  dstS / dstN are the neighbour's copies of the same buffers, already shifted so that the element
// offset of (row j, column i) is the one of `src` (gmd.cu: peer_ptr).
struct PushArgs {
  Geo g;
  const double *src[3];
  double *dstS[3], *dstN[3];
  int ns[3], nn[3];   // my top `ns` rows go north, my bottom `nn` rows go south
  u64 *page;          // my signal page
  u64 *sigS, *sigN;   // the word each neighbour waits on (its SP_SIG+1 / SP_SIG+0), or null
  unsigned k;         // halo epoch = page[SP_XBASE] + k
  int tseq;
};
__global__ void __launch_bounds__(256) k_halo_push(const PushArgs a) {
  trace_in(a.tseq);
  const int nlon = a.g.nlon, nr = a.g.r1 - a.g.r0;
  const int n2 = nlon >> 1;  // 16-byte units per row
  int first[7];              // prefix of rows: (f0 S, f0 N, f1 S, f1 N, f2 S, f2 N)
  first[0] = 0;
#pragma unroll
  for (int f = 0; f < 3; f++) {
    first[2 * f + 1] = first[2 * f] + ((a.src[f] && a.dstS[f]) ? a.nn[f] : 0);
    first[2 * f + 2] = first[2 * f + 1] + ((a.src[f] && a.dstN[f]) ? a.ns[f] : 0);
  }
  const int total = first[6] * n2;
  for (int idx = blockIdx.x * 256 + threadIdx.x; idx < total; idx += gridDim.x * 256) {
    const int row = idx / n2, c = idx - row * n2;
    int seg = 0;
#pragma unroll
    for (int q = 1; q < 6; q++) seg += (row >= first[q]) ? 1 : 0;
    const int f = seg >> 1, north = seg & 1, r = row - first[seg];
    const int lj = north ? nr - a.ns[f] + r : r;
    const ptrdiff_t o = (ptrdiff_t)lj * nlon + 2 * c;
    const double2 v = *reinterpret_cast<const double2 *>(a.src[f] + o);
    double *d = north ? a.dstN[f] : a.dstS[f];
    *reinterpret_cast<double2 *>(d + o) = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned *ticket = reinterpret_cast<unsigned *>(a.page + SP_TICKET);
    const unsigned t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {  // every CTA's rows are on their way: release both neighbours
      *ticket = 0;
      __threadfence_system();
      const u64 ep = a.page[SP_XBASE] + a.k;
      if (a.sigS) st_release_sys(a.sigS, ep);
      if (a.sigN) st_release_sys(a.sigN, ep);
    }
  }
  trace_out(a.tseq);
}
// consumers that are not the stage kernel: block the stream until halo epoch page[SP_XBASE] + k has arrived
__global__ void k_halo_wait(u64 *page, unsigned k, int sides, int tseq) {
  trace_in(tseq);
  if (threadIdx.x == 0) {
    const u64 want = page[SP_XBASE] + k;
    GMD_TRACE_T0();
    if (sides & 1) spin_until(page + SP_SIG, want, page);
    if (sides & 2) spin_until(page + SP_SIG + 1, want, page);
    GMD_TRACE_WAIT(tseq);
  }
  trace_out(tseq);
}
// end of a unit of work (one model step, one direct API call): all incoming halos of the unit have landed (so the
// host may touch the ghost rows, and a neighbour may free its slab), then the epoch bases advance.  The epochs
// inside a unit are base + k with k baked into the launches, which is what makes a captured step replayable.
__global__ void k_unit_end(u64 *page, unsigned nx, unsigned nred, int sides, int tseq) {
  trace_in(tseq, 1);   // runs after k_diag_store has advanced the step counter
  if (threadIdx.x == 0) {
    const u64 want = page[SP_XBASE] + nx;
    if (sides & 1) spin_until(page + SP_SIG, want, page);
    if (sides & 2) spin_until(page + SP_SIG + 1, want, page);
    page[SP_XBASE] = want;
    page[SP_RBASE] = page[SP_RBASE] + nred;
  }
  trace_out(tseq, 1);
}

// ---------------------------------------------------------------------------------------------------------
// Element-wise kernels
// ---------------------------------------------------------------------------------------------------------
struct UpdateArgs {
  Geo g;
  const double *OU, *OV, *Ogd;
  const double *TU, *TV, *Tgd;
  double *NU, *NV, *Ngd;
  double dt;
  const double *ip;   // device {ip1, ip2} or NULL
  int qcon, beta_mode;  // beta_mode 0: dt ; 1: dt*beta (predict_correct :786-790) ; 2: dt * (beta*4/dt0) (isp :745-750)
                        // 3: runge_kutta (specified extension): beta = ip1 / (3 ip2), ip1 = the stage tendency products
  double dt0;
  double *beta_out;   // device scalar, written by one thread
  int with_gd;        // 0: slow pass, gd is shared
  int rb[2], re[2];   // up to two row ranges handled by this launch (empty when rb >= re)
  int tseq;
};


// update_state on stored tendencies (src/dycore_mod.F90:600-652), U/V/gd only: new = old + dt' * tend
__global__ void __launch_bounds__(256) k_update(const UpdateArgs a) {
  trace_in(a.tseq);
  double dt = a.dt;
  if (a.beta_mode) {
    double beta = beta_from_ip(a.ip, a.qcon);
    if (a.beta_mode == 2) beta = beta * 4.0 / a.dt0;
    if (a.beta_mode == 3) {
      const double ip1 = a.ip[0], ip2 = a.ip[1];
      beta = (a.qcon && ip1 != 0.0 && ip2 != 0.0) ? ip1 / (3.0 * ip2) : 1.0;
    }
    dt = a.dt * beta;
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.beta_out) *a.beta_out = beta;
  }
  const int nlon = a.g.nlon, nlat = a.g.nlat;
#pragma unroll
  for (int z = 0; z < 2; z++) {
    if (a.rb[z] >= a.re[z]) continue;
    // rows may lie south of the band (ghost rows): count from the first row of the range, not from r0
    const ptrdiff_t k0 = (ptrdiff_t)(a.rb[z] - a.g.r0) * nlon, kn = (ptrdiff_t)(a.re[z] - a.rb[z]) * nlon;
    for (ptrdiff_t q = (ptrdiff_t)blockIdx.x * 256 + threadIdx.x; q < kn; q += (ptrdiff_t)gridDim.x * 256) {
      const ptrdiff_t k = k0 + q;
      const int j = a.rb[z] + (int)(q / (ptrdiff_t)nlon);
      if (j >= 1 && j <= nlat - 2) a.NU[k] = a.OU[k] + dt * a.TU[k];
      else a.NU[k] = a.OU[k];
      if (j <= nlat - 2) a.NV[k] = a.OV[k] + dt * a.TV[k];
      if (a.with_gd) a.Ngd[k] = a.Ogd[k] + dt * a.Tgd[k];
    }
  }
  trace_out(a.tseq);
}

// iap_transform (src/types_mod.F90:399-426): U = 0.5 (s_i + s_i+1) u ; V = 0.5 (s_j + s_j+1) v
__global__ void __launch_bounds__(256) k_iap(Geo g, const double *u, const double *v, const double *gd, double *U,
                                             double *V) {
  const int nlon = g.nlon;
  const size_t total = (size_t)(g.r1 - g.r0) * (size_t)nlon;
  for (ptrdiff_t k = (ptrdiff_t)blockIdx.x * 256 + threadIdx.x; k < (ptrdiff_t)total; k += (ptrdiff_t)gridDim.x * 256) {
    const int lj = (int)(k / (ptrdiff_t)nlon), i = (int)(k - (ptrdiff_t)lj * nlon), j = g.r0 + lj;
    const int ie = (i + 1 == nlon) ? 0 : i + 1;
    const double s = sqrt(gd[k]);
    const double se = sqrt(gd[(ptrdiff_t)lj * nlon + ie]);
    U[k] = 0.5 * (s + se) * u[k];
    if (j <= g.nlat - 2) {
      const double sn = sqrt(gd[k + nlon]);
      V[k] = 0.5 * (s + sn) * v[k];
    }
  }
}

// inverse: u = 2U/(s_i+s_i+1), v = 2V/(s_j+s_j+1) on rows [ja, jb) (may include ghost rows inside the globe)
__global__ void __launch_bounds__(256) k_derive(Geo g, int ja, int jb, const double *U, const double *V,
                                                const double *gd, double *u, double *v, double *s_out) {
  const int nlon = g.nlon;
  const size_t total = (size_t)(jb - ja) * (size_t)nlon;
  for (ptrdiff_t k = (ptrdiff_t)blockIdx.x * 256 + threadIdx.x; k < (ptrdiff_t)total; k += (ptrdiff_t)gridDim.x * 256) {
    const int lj = (int)(k / (ptrdiff_t)nlon), i = (int)(k - (ptrdiff_t)lj * nlon), j = ja + lj;
    const ptrdiff_t o = (ptrdiff_t)(j - g.r0) * nlon + i;
    const int ie = (i + 1 == nlon) ? 0 : i + 1;
    const double s = sqrt(gd[o]);
    const double se = sqrt(gd[(ptrdiff_t)(j - g.r0) * nlon + ie]);
    if (u) u[o] = U[o] * 2.0 / (s + se);
    if (v && j <= g.nlat - 2) v[o] = V[o] * 2.0 / (s + sqrt(gd[o + nlon]));
    if (s_out) s_out[o] = s;
  }
}

// diag_run totals (src/diag_mod.F90:71-77,98-121): per-CTA partials {sum cos dlon dlat gd, energy}
__global__ void __launch_bounds__(256) k_diag(Geo g, Tab t, const double *U, const double *V, const double *gd,
                                              const double *ghs, double dlon, double dlat, double *partials, int tseq) {
  __shared__ double red[32];
  trace_in(tseq);
  const int nlon = g.nlon;
  const size_t total = (size_t)(g.r1 - g.r0) * (size_t)nlon;
  double m = 0.0, e = 0.0;
  for (ptrdiff_t k = (ptrdiff_t)blockIdx.x * 256 + threadIdx.x; k < (ptrdiff_t)total; k += (ptrdiff_t)gridDim.x * 256) {
    const int j = g.r0 + (int)(k / (ptrdiff_t)nlon);
    const double cf = t.cosf[j];
    const double gdv = gd[k], gh = gdv + ghs[k];
    m = m + cf * dlon * dlat * gdv;
    if (j >= 1 && j <= g.nlat - 2) {
      const double Uv = U[k];
      e = e + Uv * Uv * cf;
    }
    if (j <= g.nlat - 2) {
      const double Vv = V[k];
      e = e + Vv * Vv * t.cosh[j];
    }
    e = e + gh * gh * cf;
  }
  const double rm = block_sum<256>(m, red);
  const double re = block_sum<256>(e, red);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = rm;
    partials[2 * blockIdx.x + 1] = re;
  }
  trace_out(tseq);
}

// ring[ctr % nring] = {mass * radius^2, energy, beta}; the step counter lives on the device so that a
// captured graph of one model step is replayable for any step number
__global__ void k_diag_store(const double *sums, const double *beta, double radius, double *ring, int *ctr,
                             int advance, int nring, int tseq) {
  trace_in(tseq);
  trace_out(tseq);   // (a few hundred ns; stamped before the counter advances)
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int c = *ctr;
    if (advance) {
      c += 1;
      *ctr = c;
    }
    const int slot = c % nring;
    ring[3 * slot] = sums[0] * (radius * radius);
    ring[3 * slot + 1] = sums[1];
    ring[3 * slot + 2] = *beta;
  }
}

// <a, b> with the inner-product weights (src/types_mod.F90:347-397); out partials pairs {dot, 0}
__global__ void __launch_bounds__(256) k_dot(Geo g, Tab t, const double *aU, const double *aV, const double *aG,
                                             const double *bU, const double *bV, const double *bG, int with_gd,
                                             double *partials, int slot, int accumulate = 0) {
  __shared__ double red[32];
  const int nlon = g.nlon;
  const size_t total = (size_t)(g.r1 - g.r0) * (size_t)nlon;
  double s = 0.0;
  for (ptrdiff_t k = (ptrdiff_t)blockIdx.x * 256 + threadIdx.x; k < (ptrdiff_t)total; k += (ptrdiff_t)gridDim.x * 256) {
    const int j = g.r0 + (int)(k / (ptrdiff_t)nlon);
    if (j >= 1 && j <= g.nlat - 2) s = s + aU[k] * bU[k] * t.cosf[j];
    if (j <= g.nlat - 2) s = s + aV[k] * bV[k] * t.cosh[j];
    if (with_gd) s = s + aG[k] * bG[k] * t.cosf[j];
  }
  const double r = block_sum<256>(s, red);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x + slot] = accumulate ? partials[2 * blockIdx.x + slot] + r : r;
    if (slot == 0 && !accumulate) partials[2 * blockIdx.x + 1] = 0.0;
  }
}

// y = alpha * x + beta * y on the three tendency arrays (tend algebra, src/types_mod.F90:229-345)
__global__ void __launch_bounds__(256) k_axpby3(size_t total, double alpha, const double *xU, const double *xV,
                                                const double *xG, double beta, double *yU, double *yV, double *yG) {
  for (ptrdiff_t k = (ptrdiff_t)blockIdx.x * 256 + threadIdx.x; k < (ptrdiff_t)total; k += (ptrdiff_t)gridDim.x * 256) {
    yU[k] = (beta == 0.0 ? 0.0 : beta * yU[k]) + (alpha == 0.0 ? 0.0 : alpha * xU[k]);
    yV[k] = (beta == 0.0 ? 0.0 : beta * yV[k]) + (alpha == 0.0 ? 0.0 : alpha * xV[k]);
    yG[k] = (beta == 0.0 ? 0.0 : beta * yG[k]) + (alpha == 0.0 ? 0.0 : alpha * xG[k]);
  }
}

// diag vor/div (src/diag_mod.F90:47-69) from derived u, v
__global__ void __launch_bounds__(256) k_vor_div(Geo g, Tab t, const double *u, const double *v, double *vor,
                                                 double *div) {
  const int nlon = g.nlon, nlat = g.nlat;
  const size_t total = (size_t)(g.r1 - g.r0) * (size_t)nlon;
  for (ptrdiff_t k = (ptrdiff_t)blockIdx.x * 256 + threadIdx.x; k < (ptrdiff_t)total; k += (ptrdiff_t)gridDim.x * 256) {
    const int lj = (int)(k / (ptrdiff_t)nlon), i = (int)(k - (ptrdiff_t)lj * nlon), j = g.r0 + lj;
    const int iw = (i == 0) ? nlon - 1 : i - 1, ie = (i + 1 == nlon) ? 0 : i + 1;
    const ptrdiff_t row = (ptrdiff_t)lj * nlon;
    if (j >= 1 && j <= nlat - 2) {
      const double um1 = u[row + iw], up1 = u[k];
      const double vm1 = v[k - nlon] * t.cosh[j - 1], vp1 = v[k] * t.cosh[j];
      div[k] = (up1 - um1) / t.fdlon[j] + (vp1 - vm1) / t.fdlat[j];
    } else {
      div[k] = 0.0;
    }
    if (j <= nlat - 2) {
      const double um1 = u[k], up1 = u[k + nlon];
      const double vm1 = v[k], vp1 = v[row + ie];
      vor[k] = (vp1 - vm1) / t.hdlon[j] - (up1 - um1) / t.hdlat[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// ordinary_diffusion (src/diffusion_mod.F90:74-217)
// ---------------------------------------------------------------------------------------------------------
// one scalar-Laplacian pass on (u, v, gd) -> (ud, vd, gdd), rows [r0, r1); gd pole rows by k_lap_pole
__global__ void __launch_bounds__(256) k_laplace(Geo g, Tab t, const double *u, const double *v, const double *gd,
                                                 double *ud, double *vd, double *gdd) {
  const int nlon = g.nlon, nlat = g.nlat;
  const size_t total = (size_t)(g.r1 - g.r0) * (size_t)nlon;
  for (ptrdiff_t k = (ptrdiff_t)blockIdx.x * 256 + threadIdx.x; k < (ptrdiff_t)total; k += (ptrdiff_t)gridDim.x * 256) {
    const int lj = (int)(k / (ptrdiff_t)nlon), i = (int)(k - (ptrdiff_t)lj * nlon), j = g.r0 + lj;
    const int iw = (i == 0) ? nlon - 1 : i - 1, ie = (i + 1 == nlon) ? 0 : i + 1;
    const ptrdiff_t row = (ptrdiff_t)lj * nlon;
    if (j >= 1 && j <= nlat - 2) {
      const double dl2 = t.fdlon[j] * t.fdlon[j], dt2 = t.fdlat[j] * t.fdlat[j];
      const double hc1 = t.cosh[j], hc0 = t.cosh[j - 1], fc = t.cosf[j];
      {
        const double q = gd[k];
        gdd[k] = (gd[row + ie] - 2 * q + gd[row + iw]) / dl2 +
                 ((gd[k + nlon] - q) * hc1 - (q - gd[k - nlon]) * hc0) / dt2 * fc;   // :109-114
      }
      {
        const double q = u[k];
        ud[k] = (u[row + ie] - 2 * q + u[row + iw]) / dl2 +
                ((u[k + nlon] - q) * hc1 - (q - u[k - nlon]) * hc0) / dt2 * fc;      // :135-138
      }
    } else {
      ud[k] = 0.0;  // pole rows of ud are never written in the reference (stay 0)
    }
    if (j <= nlat - 2) {
      const double q = v[k];
      const double hl2 = t.hdlon[j] * t.hdlon[j], ht2 = t.hdlat[j] * t.hdlat[j];
      double r = (v[row + ie] - 2 * q + v[row + iw]) / hl2;                          // :145-147
      if (j >= 1 && j <= nlat - 3)
        r = r + ((v[k + nlon] - q) * t.cosf[j + 1] - (q - v[k - nlon]) * t.cosf[j]) / ht2 * t.cosh[j];  // :148-157
      else if (j == 0)
        r = r + (v[k + nlon] - q) * t.cosf[j + 1] / ht2 * t.cosh[j];                 // :158-163
      else
        r = r - (q - v[k - nlon]) * t.cosf[j] / ht2 * t.cosh[j];                     // :164-170
      vd[k] = r;
    }
  }
}

// gd pole caps of the Laplacian (src/diffusion_mod.F90:116-133); blockIdx.x: 0 south, 1 north
__global__ void __launch_bounds__(PT_EW) k_lap_pole(Geo g, Tab t, const double *gd, double *gdd, int do_south,
                                                 int do_north) {
  __shared__ double red[32];
  __shared__ double bc;
  const int n = g.nlon, nlat = g.nlat;
  const bool south = (blockIdx.x == 0);
  if ((south && !do_south) || (!south && !do_north)) return;
  const int j = south ? 0 : nlat - 1;
  const ptrdiff_t off = (ptrdiff_t)(j - g.r0) * n;
  double acc = 0.0;
  if (south) {
    for (int i = threadIdx.x; i < n; i += PT_EW) acc = acc + gd[off + n + i] - gd[off + i];
  } else {
    for (int i = threadIdx.x; i < n; i += PT_EW) acc = acc - (gd[off + i] - gd[off - n + i]);
  }
  const double r = block_sum<PT_EW>(acc, red);
  if (threadIdx.x == 0) {
    const double hc = south ? t.cosh[0] : t.cosh[nlat - 2];
    bc = r * hc / (t.fdlat[j] * t.fdlat[j]) * t.cosf[j] / n;
  }
  __syncthreads();
  const double vv = bc;
  for (int i = threadIdx.x; i < n; i += PT_EW) gdd[off + i] = vv;
}

// q += sign dt coef lap(q) for gd,u,v, then iap_transform (src/diffusion_mod.F90:195-215)
__global__ void __launch_bounds__(256) k_diff_update(Geo g, const double *u, const double *v, const double *gd,
                                                     const double *ud, const double *vd, const double *gdd,
                                                     double sdc, double *NU, double *NV, double *Ngd) {
  const int nlon = g.nlon, nlat = g.nlat;
  const size_t total = (size_t)(g.r1 - g.r0) * (size_t)nlon;
  for (ptrdiff_t k = (ptrdiff_t)blockIdx.x * 256 + threadIdx.x; k < (ptrdiff_t)total; k += (ptrdiff_t)gridDim.x * 256) {
    const int lj = (int)(k / (ptrdiff_t)nlon), i = (int)(k - (ptrdiff_t)lj * nlon), j = g.r0 + lj;
    const int ie = (i + 1 == nlon) ? 0 : i + 1;
    const ptrdiff_t ke = (ptrdiff_t)lj * nlon + ie;
    const double g0 = gd[k] + sdc * gdd[k];
    const double ge = gd[ke] + sdc * gdd[ke];
    const double s0 = sqrt(g0), se = sqrt(ge);
    Ngd[k] = g0;
    const double un = u[k] + sdc * ud[k];
    NU[k] = 0.5 * (s0 + se) * un;
    if (j <= nlat - 2) {
      const double gn = gd[k + nlon] + sdc * gdd[k + nlon];
      const double vn = v[k] + sdc * vd[k];
      NV[k] = 0.5 * (s0 + sqrt(gn)) * vn;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Row-pair forms of k_derive / k_laplace / k_diff_update: one latitude row per blockIdx.y, one column pair per thread
// (16-byte accesses, the row's metric terms read once per thread, no index division -- the element-indexed kernels
// above spend ~600 / 280 / 185 instructions per column, mostly on k / nlon and the per-element table reads:
// profiles/r2_s_ncu_diffusion_summary.txt).  Every value is formed by the same expression as above, so the results are
// those of the element-indexed kernels.  Plain (coherent) loads: ghost rows of a band are written by the neighbour
// rank between launches.  num_lon is even (gmd_create), field rows are 16-byte aligned.
// ---------------------------------------------------------------------------------------------------------
constexpr int EW2 = 128;
__device__ __forceinline__ D2 ldp2(const double *p) {
  const double2 v = *reinterpret_cast<const double2 *>(p);
  D2 r;
  r.x = v.x;
  r.y = v.y;
  return r;
}

__global__ void __launch_bounds__(EW2) k_derive2(Geo g, int ja, const double *U, const double *V, const double *gd,
                                                 double *u, double *v, double *s_out) {
  const int nlon = g.nlon;
  const int i = 2 * (blockIdx.x * EW2 + threadIdx.x);
  if (i >= nlon) return;
  const int j = ja + (int)blockIdx.y;
  const int ie = (i + 2 == nlon) ? 0 : i + 2;
  const ptrdiff_t row = (ptrdiff_t)(j - g.r0) * nlon, o = row + i;
  const D2 q = ldp2(gd + o);
  const double s0 = sqrt(q.x), s1 = sqrt(q.y), s2 = sqrt(gd[row + ie]);
  if (u) {
    const D2 a = ldp2(U + o);
    st2(u + o, a.x * 2.0 / (s0 + s1), a.y * 2.0 / (s1 + s2));
  }
  if (v && j <= g.nlat - 2) {
    const D2 a = ldp2(V + o), qn = ldp2(gd + o + nlon);
    st2(v + o, a.x * 2.0 / (s0 + sqrt(qn.x)), a.y * 2.0 / (s1 + sqrt(qn.y)));
  }
  if (s_out) st2(s_out + o, s0, s1);
}

// (east - 2 q + west) / dl2 + ((north - q) c1 - (q - south) c0) / dt2 * c   (src/diffusion_mod.F90:109-114,135-138,148-157)
__device__ __forceinline__ double lap5(double q, double qe, double qw, double qn, double qs, double dl2, double dt2,
                                       double c1, double c0, double c) {
  return (qe - 2 * q + qw) / dl2 + ((qn - q) * c1 - (q - qs) * c0) / dt2 * c;
}

__global__ void __launch_bounds__(EW2) k_laplace2(Geo g, Tab t, const double *u, const double *v, const double *gd,
                                                  double *ud, double *vd, double *gdd) {
  const int nlon = g.nlon, nlat = g.nlat;
  const int i = 2 * (blockIdx.x * EW2 + threadIdx.x);
  if (i >= nlon) return;
  const int lj = (int)blockIdx.y, j = g.r0 + lj;
  const int iw = (i == 0) ? nlon - 1 : i - 1, ie = (i + 2 == nlon) ? 0 : i + 2;
  const ptrdiff_t row = (ptrdiff_t)lj * nlon, k = row + i;
  if (j >= 1 && j <= nlat - 2) {
    const double dl2 = t.fdlon[j] * t.fdlon[j], dt2 = t.fdlat[j] * t.fdlat[j];
    const double hc1 = t.cosh[j], hc0 = t.cosh[j - 1], fc = t.cosf[j];
    {
      const D2 q = ldp2(gd + k), qn = ldp2(gd + k + nlon), qs = ldp2(gd + k - nlon);
      const double qw = gd[row + iw], qe = gd[row + ie];
      st2(gdd + k, lap5(q.x, q.y, qw, qn.x, qs.x, dl2, dt2, hc1, hc0, fc),
          lap5(q.y, qe, q.x, qn.y, qs.y, dl2, dt2, hc1, hc0, fc));
    }
    {
      const D2 q = ldp2(u + k), qn = ldp2(u + k + nlon), qs = ldp2(u + k - nlon);
      const double qw = u[row + iw], qe = u[row + ie];
      st2(ud + k, lap5(q.x, q.y, qw, qn.x, qs.x, dl2, dt2, hc1, hc0, fc),
          lap5(q.y, qe, q.x, qn.y, qs.y, dl2, dt2, hc1, hc0, fc));
    }
  } else {
    st2(ud + k, 0.0, 0.0);  // pole rows of ud are never written in the reference (stay 0)
  }
  if (j <= nlat - 2) {
    const D2 q = ldp2(v + k);
    const double qw = v[row + iw], qe = v[row + ie];
    const double hl2 = t.hdlon[j] * t.hdlon[j], ht2 = t.hdlat[j] * t.hdlat[j];
    double r0 = (q.y - 2 * q.x + qw) / hl2, r1 = (qe - 2 * q.y + q.x) / hl2;                      // :145-147
    if (j >= 1 && j <= nlat - 3) {
      const D2 qn = ldp2(v + k + nlon), qs = ldp2(v + k - nlon);
      const double c1 = t.cosf[j + 1], c0 = t.cosf[j], c = t.cosh[j];
      r0 = r0 + ((qn.x - q.x) * c1 - (q.x - qs.x) * c0) / ht2 * c;                                 // :148-157
      r1 = r1 + ((qn.y - q.y) * c1 - (q.y - qs.y) * c0) / ht2 * c;
    } else if (j == 0) {
      const D2 qn = ldp2(v + k + nlon);
      const double c1 = t.cosf[j + 1], c = t.cosh[j];
      r0 = r0 + (qn.x - q.x) * c1 / ht2 * c;                                                       // :158-163
      r1 = r1 + (qn.y - q.y) * c1 / ht2 * c;
    } else {
      const D2 qs = ldp2(v + k - nlon);
      const double c0 = t.cosf[j], c = t.cosh[j];
      r0 = r0 - (q.x - qs.x) * c0 / ht2 * c;                                                       // :164-170
      r1 = r1 - (q.y - qs.y) * c0 / ht2 * c;
    }
    st2(vd + k, r0, r1);
  }
}

__global__ void __launch_bounds__(EW2) k_diff_update2(Geo g, const double *u, const double *v, const double *gd,
                                                      const double *ud, const double *vd, const double *gdd,
                                                      double sdc, double *NU, double *NV, double *Ngd) {
  const int nlon = g.nlon, nlat = g.nlat;
  const int i = 2 * (blockIdx.x * EW2 + threadIdx.x);
  if (i >= nlon) return;
  const int lj = (int)blockIdx.y, j = g.r0 + lj;
  const int ie = (i + 2 == nlon) ? 0 : i + 2;
  const ptrdiff_t row = (ptrdiff_t)lj * nlon, k = row + i;
  const D2 q = ldp2(gd + k), qd = ldp2(gdd + k);
  const double g0 = q.x + sdc * qd.x, g1 = q.y + sdc * qd.y;
  const double ge = gd[row + ie] + sdc * gdd[row + ie];
  const double s0 = sqrt(g0), s1 = sqrt(g1), se = sqrt(ge);
  st2(Ngd + k, g0, g1);
  {
    const D2 a = ldp2(u + k), ad = ldp2(ud + k);
    const double un0 = a.x + sdc * ad.x, un1 = a.y + sdc * ad.y;
    st2(NU + k, 0.5 * (s0 + s1) * un0, 0.5 * (s1 + se) * un1);
  }
  if (j <= nlat - 2) {
    const D2 qn = ldp2(gd + k + nlon), qnd = ldp2(gdd + k + nlon);
    const double gn0 = qn.x + sdc * qnd.x, gn1 = qn.y + sdc * qnd.y;
    const D2 a = ldp2(v + k), ad = ldp2(vd + k);
    const double vn0 = a.x + sdc * ad.x, vn1 = a.y + sdc * ad.y;
    st2(NV + k, 0.5 * (s0 + sqrt(gn0)) * vn0, 0.5 * (s1 + sqrt(gn1)) * vn1);
  }
}

// ---------------------------------------------------------------------------------------------------------
// WENO advection (src/weno_mod.F90:69-300), order 2 -- unfused sweeps on derived u, v (a "next" row of the
// scope table: correct first, fused later).  All arrays are band fields.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double weno2(double fp1, double fp2, double fp3, double fn2, double fn3, double fn4) {
  const double eps = 1.0e-6;
  const double c11 = -1.0 / 2.0, c21 = 3.0 / 2.0, c12 = 1.0 / 2.0, c22 = 1.0 / 2.0;
  const double wo1 = 1.0 / 3.0, wo2 = 2.0 / 3.0;
  double fs1 = c11 * fp1 + c21 * fp2;
  double fs2 = c12 * fp2 + c22 * fp3;
  double b1 = (fp2 - fp1) * (fp2 - fp1);
  double b2 = (fp3 - fp2) * (fp3 - fp2);
  double w1 = wo1 / ((eps + b1) * (eps + b1));
  double w2 = wo2 / ((eps + b2) * (eps + b2));
  double sw = w1 + w2;
  w1 = w1 / sw;
  w2 = w2 / sw;
  double f = w1 * fs1 + w2 * fs2;
  fs1 = c11 * fn4 + c21 * fn3;
  fs2 = c12 * fn3 + c22 * fn2;
  b1 = (fn3 - fn4) * (fn3 - fn4);
  b2 = (fn2 - fn3) * (fn2 - fn3);
  w1 = wo1 / ((eps + b1) * (eps + b1));
  w2 = wo2 / ((eps + b2) * (eps + b2));
  sw = w1 + w2;
  w1 = w1 / sw;
  w2 = w2 / sw;
  f = f + (w1 * fs1 + w2 * fs2);
  return f;
}

struct WenoArgs {
  Geo g;
  Tab t;
  const double *u, *v, *U, *V;    // derived u, v and IAP U, V
  double *fpu, *fnu, *fpv, *fnv;  // split fluxes
  double *fu, *fv;                // reconstructed fluxes
  double *alon_u, *alat_u, *alon_v, *alat_v;
  int dir;                        // 0 zonal, 1 meridional
};

// Lax-Friedrichs split fluxes; rows outside their definition range are written as 0 (the reference's
// never-written zero rows of the work arrays)
__global__ void __launch_bounds__(256) k_weno_split(const WenoArgs a) {
  const int nlon = a.g.nlon, nlat = a.g.nlat;
  const size_t total = (size_t)(a.g.r1 - a.g.r0) * (size_t)nlon;
  const double amax = 20.0;
  for (ptrdiff_t k = (ptrdiff_t)blockIdx.x * 256 + threadIdx.x; k < (ptrdiff_t)total; k += (ptrdiff_t)gridDim.x * 256) {
    const int lj = (int)(k / (ptrdiff_t)nlon), i = (int)(k - (ptrdiff_t)lj * nlon), j = a.g.r0 + lj;
    const int iw = (i == 0) ? nlon - 1 : i - 1, ie = (i + 1 == nlon) ? 0 : i + 1;
    const ptrdiff_t row = (ptrdiff_t)lj * nlon;
    if (j >= 1 && j <= nlat - 2) {
      double w;
      if (a.dir == 0) w = a.u[k];
      else w = 0.25 * (a.v[k - nlon] + a.v[k] + a.v[row - nlon + ie] + a.v[row + ie]);
      a.fpu[k] = 0.5 * (w + amax) * a.U[k];
      a.fnu[k] = 0.5 * (w - amax) * a.U[k];
    } else {
      a.fpu[k] = 0.0;
      a.fnu[k] = 0.0;
    }
    if (j <= nlat - 2) {
      double w;
      if (a.dir == 0) w = 0.25 * (a.u[row + iw] + a.u[row + nlon + iw] + a.u[k] + a.u[k + nlon]);
      else w = a.v[k];
      a.fpv[k] = 0.5 * (w + amax) * a.V[k];
      a.fnv[k] = 0.5 * (w - amax) * a.V[k];
    } else {
      a.fpv[k] = 0.0;
      a.fnv[k] = 0.0;
    }
  }
}

// reads a band work array with zero outside the globe rows [lo, hi]
__device__ __forceinline__ double rd(const double *p, const Geo &g, int j, int i, int lo, int hi) {
  return (j < lo || j > hi) ? 0.0 : p[(ptrdiff_t)(j - g.r0) * g.nlon + i];
}

__global__ void __launch_bounds__(256) k_weno_flux(const WenoArgs a) {
  const int nlon = a.g.nlon, nlat = a.g.nlat;
  const size_t total = (size_t)(a.g.r1 - a.g.r0) * (size_t)nlon;
  for (ptrdiff_t k = (ptrdiff_t)blockIdx.x * 256 + threadIdx.x; k < (ptrdiff_t)total; k += (ptrdiff_t)gridDim.x * 256) {
    const int lj = (int)(k / (ptrdiff_t)nlon), i = (int)(k - (ptrdiff_t)lj * nlon), j = a.g.r0 + lj;
    const int iw = (i == 0) ? nlon - 1 : i - 1, ie = (i + 1 == nlon) ? 0 : i + 1;
    const int ie2 = (ie + 1 == nlon) ? 0 : ie + 1;
    const ptrdiff_t row = (ptrdiff_t)lj * nlon;
    if (a.dir == 0) {
      if (j >= 1 && j <= nlat - 2)
        a.fu[k] = weno2(a.fpu[row + iw], a.fpu[k], a.fpu[row + ie], a.fnu[k], a.fnu[row + ie], a.fnu[row + ie2]);
      else a.fu[k] = 0.0;
      if (j <= nlat - 2)
        a.fv[k] = weno2(a.fpv[row + iw], a.fpv[k], a.fpv[row + ie], a.fnv[k], a.fnv[row + ie], a.fnv[row + ie2]);
      else a.fv[k] = 0.0;
    } else {
      // rows beyond the work arrays' range read as 0 (zero lat halos / never-written rows)
      if (j >= 1 && j <= nlat - 2)
        a.fu[k] = weno2(rd(a.fpu, a.g, j - 1, i, 1, nlat - 2), a.fpu[k], rd(a.fpu, a.g, j + 1, i, 1, nlat - 2),
                        a.fnu[k], rd(a.fnu, a.g, j + 1, i, 1, nlat - 2), rd(a.fnu, a.g, j + 2, i, 1, nlat - 2));
      else a.fu[k] = 0.0;
      if (j <= nlat - 2)
        a.fv[k] = weno2(rd(a.fpv, a.g, j - 1, i, 0, nlat - 2), a.fpv[k], rd(a.fpv, a.g, j + 1, i, 0, nlat - 2),
                        a.fnv[k], rd(a.fnv, a.g, j + 1, i, 0, nlat - 2), rd(a.fnv, a.g, j + 2, i, 0, nlat - 2));
      else a.fv[k] = 0.0;
    }
  }
}

__global__ void __launch_bounds__(256) k_weno_adv(const WenoArgs a) {
  const int nlon = a.g.nlon, nlat = a.g.nlat;
  const size_t total = (size_t)(a.g.r1 - a.g.r0) * (size_t)nlon;
  for (ptrdiff_t k = (ptrdiff_t)blockIdx.x * 256 + threadIdx.x; k < (ptrdiff_t)total; k += (ptrdiff_t)gridDim.x * 256) {
    const int lj = (int)(k / (ptrdiff_t)nlon), i = (int)(k - (ptrdiff_t)lj * nlon), j = a.g.r0 + lj;
    const int iw = (i == 0) ? nlon - 1 : i - 1, ie = (i + 1 == nlon) ? 0 : i + 1;
    const ptrdiff_t row = (ptrdiff_t)lj * nlon;
    if (a.dir == 0) {
      if (j >= 1 && j <= nlat - 2)   // B8: half_dlon(j) on a full row, src/weno_mod.F90:151-155
        a.alon_u[k] = (a.fu[k] - a.fu[row + iw] - (a.u[row + ie] - a.u[row + iw]) * a.U[k] * 0.25) / a.t.hdlon[j];
      if (j <= nlat - 2)
        a.alon_v[k] = (a.fv[k] - a.fv[row + iw] -
                       (a.u[k] + a.u[k + nlon] - a.u[row + iw] - a.u[row + nlon + iw]) * a.V[k] * 0.25) /
                      a.t.hdlon[j];
    } else {
      if (j >= 1 && j <= nlat - 2)
        a.alat_u[k] = (a.fu[k] - rd(a.fu, a.g, j - 1, i, 1, nlat - 2) -
                       (a.v[row + iw] + a.v[k] - a.v[row - nlon + iw] - a.v[k - nlon]) * a.U[k] * 0.25) /
                      a.t.fdlat[j];
      if (j <= nlat - 2)
        a.alat_v[k] = (a.fv[k] - rd(a.fv, a.g, j - 1, i, 0, nlat - 2) -
                       (rd(a.v, a.g, j + 1, i, 0, nlat - 2) - rd(a.v, a.g, j - 1, i, 0, nlat - 2)) * a.V[k] * 0.25) /
                      a.t.hdlat[j];
    }
  }
}

}  // namespace gmd

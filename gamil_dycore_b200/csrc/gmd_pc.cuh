// gmd_pc.cuh -- the three operator sweeps of ONE predict_correct (src/dycore_mod.F90:754-792) as a single marching
// kernel: the intermediate states never touch HBM.
//
//   tend1 = L(old);  A = old + dt/2 tend1          (S1, :770-773)
//   tend2 = L(A);    B = old + dt/2 tend2          (S2, :775-778)
//   tend3 = L(B);    ip1 = <tend2, tend3>, ip2 = <tend3, tend3>      (S3a, :780-785)
//
// k_stage runs these as three sweeps over the band (7..13 + 13 + 10 words per column).  Every sweep is a south -> north
// march whose stencil reaches one row south and two rows north, so the three marches can run as a WAVEFRONT a few rows
// apart: a CTA is three warps working on the SAME 64-column strip -- warp 0 evaluates S1, warp 1 S2 four rows behind
// it, warp 2 S3a another four rows behind -- and the rows of A, `old`, B and tend2 travel from warp to warp through
// small shared-memory rings (5 / 4 rows deep), never through global memory.  What S1 consumes from global memory --
// the old state, ghs, the deferred update's tendency, the per-row coefficient records -- arrives as packets of
// cp.async.bulk copies counted on mbarriers, three rows ahead, issued by one lane of the S2 warp.  Per column the
// kernel reads old (3) + ghs (1) [+ the deferred update's tendency (3)] and writes tend3 (3) [+ the materialised old
// state (3)]: 13 words instead of 36 for a fast predict_correct with the deferred update.
//
// Each warp keeps the row window of ITS stage in registers exactly as k_stage does (tend_col is shared); the warps
// of a CTA advance in lockstep, one row per tick, with one named barrier per tick.  Three chained stencils need
// (3 west, 6 east) halo columns: a strip holds 64 columns, 54 are outputs.  A CTA's chunk of rows [ja, jb) is
// evaluated by S3a; S2 covers one row more to the south and two to the north, S1 two and four (recomputed by the
// neighbouring chunk as well: no exchange between CTAs).  The same widening makes the kernel the wide-halo
// predict_correct of a latitude band without further ado.
//
// Rows that depend on a full-row operation -- zonal filter rows, reduced rows, pole caps -- cannot be part of the
// wavefront: the host keeps the fused rows at least (2 south, 4 north) rows away from them and runs the remaining
// rows next to the poles through k_stage + k_polar as before, concurrently (gmd.cu: launch_pc and the m->fz_active
// branch of stage()).
//
// Product build only (the strict build keeps the three-sweep path, whose operand order follows the reference).
#pragma once

namespace gmd {

constexpr int WOUT3 = 54;      // output columns per strip (64 held)
constexpr int PC_BX = 96;      // three warps: S1, S2, S3a
constexpr int PC_D1 = 4;       // S2 starts 4 ticks after S1
constexpr int PC_D2 = 8;       // S3a starts 8 ticks after S1
#ifndef GMD_PC_UNROLL
#define GMD_PC_UNROLL 1
#endif
#ifndef GMD_PC_MINB
#define GMD_PC_MINB 5          // resident CTAs per SM the register allocation is capped for (15 warps)
#endif
// Rings between the warps, one 512-byte line (32 lanes x 16 bytes) per (row, field), 3 fields per row:
//   A (S1 -> S2) and B (S2 -> S3a): the stage states, 5 rows deep; O (S1 -> S2): the old state, P (S2 -> S3a): tend2,
//   4 rows deep (depths checked against the tick table below by tests/test_host.py::test_pc_ring_schedule)
enum { RG_A = 0, RG_O = 1, RG_B = 2, RG_P = 3 };
constexpr int PC_DEPTH_AB = 5, PC_DEPTH_OP = 4;
static_assert((PC_DEPTH_OP & (PC_DEPTH_OP - 1)) == 0, "PC_DEPTH_OP must be a power of two");
constexpr int PC_RING_LINES = 3 * (2 * PC_DEPTH_AB + 2 * PC_DEPTH_OP);   // 54 lines = 27648 bytes
__host__ __device__ constexpr int pc_ring_line0(int kind) {
  return kind == RG_A ? 0 : kind == RG_B ? 3 * PC_DEPTH_AB : kind == RG_O ? 6 * PC_DEPTH_AB : 6 * PC_DEPTH_AB + 3 * PC_DEPTH_OP;
}
// Input ring of S1: everything one iteration of its row loop consumes from global memory -- gd(j+2), U(j+1), V(j+1),
// ghs(j+1) of the evaluated state and, with the deferred update, the tendency rows that go with them -- is a PACKET of
// up to 7 lines, fetched PC_IN_DEPTH - 1 = 3 rows ahead by bulk asynchronous copies (cp.async.bulk, the TMA engine's
// 1-D form: one thread issues the copies of a packet, completion is counted in bytes on an mbarrier).  Only 5 of the 15
// resident warps of an SM issue global loads; without the deep prefetch there are too few bytes in flight to cover
// the HBM latency (measured: 6 barrier-stall cycles per issued instruction with a one-row register prefetch).
constexpr int PC_IN_DEPTH = 4, PC_IN_NF = 7;
enum { IN_GD = 0, IN_U = 1, IN_V = 2, IN_HS = 3, IN_TG = 4, IN_TU = 5, IN_TV = 6 };
// Row records (128 bytes per latitude row, gmd_kernels.cuh RC_*): a 16-row ring per CTA, filled with S1's packets (the
// record of row j+1 travels with the packet of iteration j) and read by all three warps -- S3a is 10 rows behind S1.
constexpr int PC_REC_DEPTH = 16;
constexpr size_t PC_SMEM = (size_t)(PC_RING_LINES + PC_IN_DEPTH * PC_IN_NF) * 512 + PC_REC_DEPTH * RC_N * 8 + PC_IN_DEPTH * 8;   // 44064 bytes per CTA

// Tick table (rows relative to ja; tick t ends with barrier t):
//   S1  evaluates row x in tick x + 2            and writes A(x), old(x) into the rings
//   S2  evaluates row x in tick x + 1 + PC_D1    reading gd_A(x+2) [tick x+4], U,V_A(x+1) [x+3], old(x) [x+2];
//       its prologue (tick PC_D1) reads A rows -2..0 [ticks 0..2];  writes B(x), tend2(x)
//   S3a evaluates row x in tick x + PC_D2        reading gd_B(x+2) [tick x+7], U,V_B(x+1) [x+6], tend2(x) [x+5];
//       its prologue (tick PC_D2) reads B rows -1..1 [ticks 4..6]
// A row is overwritten `depth` ticks after it was written; the longest-lived rows are the prologue rows (4 ticks).
// The three warps reach the barrier from three different loops, each warp converged.  That is what bar.sync (=
// barrier.sync.aligned) asks for; compute-sanitizer --tool synccheck nevertheless wants a whole-CTA aligned barrier at
// ONE instruction, so tools/sanitize_fused.sh checks a build with the unaligned form (-DGMD_PC_UNALIGNED_BAR=1: same
// results, 6 % slower on B200)
#ifndef GMD_PC_UNALIGNED_BAR
#define GMD_PC_UNALIGNED_BAR 0
#endif
__device__ __forceinline__ void pc_bar() {
#if GMD_PC_UNALIGNED_BAR
  asm volatile("barrier.sync 1, 96;" ::: "memory");
#else
  asm volatile("bar.sync 1, 96;" ::: "memory");
#endif
}

// ---- mbarrier + bulk copy (sm_90+ PTX) ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PC_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PC_DONE_%=;\n"
      "bra PC_WAIT_%=;\n"
      "PC_DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar) : "memory");
}
template <int PASS, int ADV, int ROLE, int LAZY, bool PUSH>
__device__ __forceinline__ void pc_role(const StageArgs &a, const int strip, const int ja, const int jb, const bool edgeS,
                                        const bool edgeN, D2 *const ring, double &ip1, double &ip2) {
  constexpr int ES = (ROLE == 0) ? 2 : (ROLE == 1 ? 1 : 0);
  constexpr int EN = (ROLE == 0) ? 4 : (ROLE == 1 ? 2 : 0);
  constexpr int T0 = (ROLE == 0) ? 0 : (ROLE == 1 ? PC_D1 : PC_D2);
  constexpr bool need_gh = (PASS != PASS_SLOW);
  constexpr int RIN = (ROLE == 1) ? RG_A : RG_B;   // ring the evaluated state comes from (ROLE 1, 2)
  const int nlon = a.g.nlon, nlat = a.g.nlat, r0 = a.g.r0;
  const int lane = threadIdx.x & 31;
  const int rja = ja - ES, rjb = jb + EN;
  const int nticks = (jb - ja) + PC_D2;
  int c0 = strip * WOUT3 - 4 + 2 * lane;   // even; (c0, c0+1) never straddles the seam because nlon is even
  c0 %= nlon;
  if (c0 < 0) c0 += nlon;
  const bool out = (lane >= 2) && (lane <= 1 + WOUT3 / 2) && (strip * WOUT3 + 2 * (lane - 2) < nlon);
  const ptrdiff_t nl = nlon;
  // ring slot of row x (x >= ja - 2)
  const int rbase = ja - 2;
  // 16-byte shared-memory accesses, one 512-byte line per (ring, row, field): conflict-free LDS.128 / STS.128.  The
  // ring slot of a row is (row - rbase) mod depth; the row loop carries the slots of rows j, j+1, j+2 along instead of
  // dividing (slot5 / slot4 below)
  auto rld = [&](int kind, int sl, int f) -> D2 {
    const double2 v = *reinterpret_cast<const double2 *>(ring + (pc_ring_line0(kind) + sl * 3 + f) * 32);
    D2 r;
    r.x = v.x;
    r.y = v.y;
    return r;
  };
  auto rst = [&](int kind, int sl, int f, const D2 &v) {
    *reinterpret_cast<double2 *>(ring + (pc_ring_line0(kind) + sl * 3 + f) * 32) = make_double2(v.x, v.y);
  };
  auto slot5 = [&](int x) -> int { return (int)((unsigned)(x - rbase) % (unsigned)PC_DEPTH_AB); };
  auto slot4 = [&](int x) -> int { return (int)((unsigned)(x - rbase) % (unsigned)PC_DEPTH_OP); };
  auto next5 = [](int sl) -> int { return sl == PC_DEPTH_AB - 1 ? 0 : sl + 1; };

  ptrdiff_t off = (ptrdiff_t)(rja - r0) * nl + (ptrdiff_t)c0;   // element offset of (row j, column c0)
#define ATK(p, k) ((p) + (off + (ptrdiff_t)(k) * nl))
#define AT(p, jj) ATK(p, (jj) - j)
  const D2 zero2 = {0.0, 0.0};
  double bdt = 0.0;
  if (ROLE == 0 && LAZY) bdt = a.ldt * beta_from_ip(a.lip, a.lqcon);
  auto mine = [&](int r) { return (r >= ja && r < jb) || (edgeS && r < ja) || (edgeN && r >= jb); };
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

  // ---- ROLE 0: input ring ---------------------------------------------------------------------------------------
  const D2 *const in_ring = ring + PC_RING_LINES * 32;   // (this lane's 16 bytes of) line 0 of the input ring
  const unsigned in_sa = (unsigned)__cvta_generic_to_shared(ring - lane) + (unsigned)PC_RING_LINES * 512u;
  auto ldl = [&](const D2 *q) -> D2 {
    const double2 v = *reinterpret_cast<const double2 *>(q);
    D2 r;
    r.x = v.x;
    r.y = v.y;
    return r;
  };
  const unsigned rec_sa = in_sa + (unsigned)(PC_IN_DEPTH * PC_IN_NF) * 512u;
  const unsigned in_bar = rec_sa + (unsigned)(PC_REC_DEPTH * RC_N) * 8u;
  const double *const rec = reinterpret_cast<const double *>(ring - lane) + (size_t)(PC_RING_LINES + PC_IN_DEPTH * PC_IN_NF) * 64;
  // record of row x (x >= rja0 - 1, rja0 = ja - 2: the first row S1 evaluates)
  auto recp = [&](int x) -> const double * { return rec + ((unsigned)(x - (ja - 3)) % (unsigned)PC_REC_DEPTH) * RC_N; };
  // packet of iteration r: rows r+2 of gd / Lgd, r+1 of U, V, ghs, LU, LV (64 columns from this strip's first) and the
  // record of row r+1.  Everything here is warp-uniform (uniform datapath); one elected lane issues the copies.
  constexpr unsigned pk_nf = 3u + (need_gh ? 1u : 0u) + (LAZY == 1 ? 3u : (LAZY == 2 ? 2u : 0u));
  const int rja0 = ja - 2, rjb0 = jb + 4;   // the rows S1 evaluates
  int cs = 0, n1 = 64;
  if (ROLE <= 1) {
    cs = strip * WOUT3 - 4;
    cs %= nlon;
    if (cs < 0) cs += nlon;
    n1 = min(64, nlon - cs);
  }
  auto issue_packet = [&](const int r, const int nrec) {   // nrec: records of rows r+2-nrec .. r+1
    const int it = r - rja0;
    const int st = it % PC_IN_DEPTH;
    const unsigned bar = in_bar + 8u * (unsigned)st;
    const unsigned d = in_sa + (unsigned)(st * PC_IN_NF) * 512u;
    const ptrdiff_t o1 = (ptrdiff_t)(r + 1 - r0) * nl + cs, o2 = o1 + nl;
    const unsigned b1 = (unsigned)n1 * 8u;
    if (lane == 0) {
      mbar_expect_tx(bar, pk_nf * 512u + (unsigned)nrec * (unsigned)(RC_N * 8));
      bulk_g2s(d + IN_GD * 512u, a.Egd + o2, b1, bar);
      bulk_g2s(d + IN_U * 512u, a.EU + o1, b1, bar);
      bulk_g2s(d + IN_V * 512u, a.EV + o1, b1, bar);
      if (need_gh) bulk_g2s(d + IN_HS * 512u, a.ghs + o1, b1, bar);
      if (LAZY) {
        if (LAZY == 1) bulk_g2s(d + IN_TG * 512u, a.Lgd + o2, b1, bar);
        bulk_g2s(d + IN_TU * 512u, a.LU + o1, b1, bar);
        bulk_g2s(d + IN_TV * 512u, a.LV + o1, b1, bar);
      }
      if (n1 < 64) {   // the strip crosses the seam (several times on a grid narrower than a strip)
        const ptrdiff_t w1 = o1 - cs, w2 = o2 - cs;   // column 0 of the same rows
        unsigned done = b1;
        int rem = 64 - n1;
        while (rem > 0) {
          const int n = min(rem, nlon);
          const unsigned bn = (unsigned)n * 8u;
          bulk_g2s(d + IN_GD * 512u + done, a.Egd + w2, bn, bar);
          bulk_g2s(d + IN_U * 512u + done, a.EU + w1, bn, bar);
          bulk_g2s(d + IN_V * 512u + done, a.EV + w1, bn, bar);
          if (need_gh) bulk_g2s(d + IN_HS * 512u + done, a.ghs + w1, bn, bar);
          if (LAZY) {
            if (LAZY == 1) bulk_g2s(d + IN_TG * 512u + done, a.Lgd + w2, bn, bar);
            bulk_g2s(d + IN_TU * 512u + done, a.LU + w1, bn, bar);
            bulk_g2s(d + IN_TV * 512u + done, a.LV + w1, bn, bar);
          }
          done += bn;
          rem -= n;
        }
      }
      for (int q = r + 2 - nrec; q <= r + 1; q++)
        bulk_g2s(rec_sa + ((unsigned)(q - (ja - 3)) % (unsigned)PC_REC_DEPTH) * (unsigned)(RC_N * 8), a.t.rowrec + (ptrdiff_t)q * RC_N,
                 (unsigned)(RC_N * 8), bar);
    }
  };
  // (the barriers are initialised by k_pc before the warps part ways)  S1 sends for its first packets itself -- the first
  // brings the records of rows rja0-1 .. rja0+1 --, from then on S2's warp, the one with the fewest instructions per
  // row, issues the packet S1 consumes PC_IN_DEPTH - 1 ticks later: in tick t the slot S1 read in tick t-1 is free
  if (ROLE == 0) {
    for (int q = 0; q < PC_IN_DEPTH - 1; q++)
      if (rja0 + q < rjb0) issue_packet(rja0 + q, q == 0 ? 3 : 1);
  }
  for (int t = 0; t < T0; t++) {
    if (ROLE == 1 && rja0 + t + (PC_IN_DEPTH - 1) < rjb0) issue_packet(rja0 + t + (PC_IN_DEPTH - 1), 1);
    pc_bar();
  }
  // ---- prologue: rows rja-1, rja, rja+1 of sqrt(gd); rows rja-1, rja of U, V; gd + ghs of row rja -----------------
  D2 sm_, s0, sp, sq, u0, up, vm, v0, vp, Um, U0, Up, Vm, V0, Vp;
  D2 g0 = zero2, gp = zero2;
  D2 gr0, gr1;   // gd of rows j, j+1 without ghs
  {
    const int j = rja;
    D2 a0, a1, a2;
    if (ROLE == 0) {
      a0 = ld2(AT(a.Egd, j - 1));
      a1 = ld2(AT(a.Egd, j));
      a2 = ld2(AT(a.Egd, j + 1));
      Um = ld2(AT(a.EU, j - 1));
      U0 = ld2(AT(a.EU, j));
      Vm = ld2(AT(a.EV, j - 1));
      V0 = ld2(AT(a.EV, j));
      if (LAZY) {
        const D2 tUm = ld2(AT(a.LU, j - 1)), tU0 = ld2(AT(a.LU, j)), tVm = ld2(AT(a.LV, j - 1)), tV0 = ld2(AT(a.LV, j));
        if (LAZY == 1) {
          const D2 t0 = ld2(AT(a.Lgd, j - 1)), t1 = ld2(AT(a.Lgd, j)), t2 = ld2(AT(a.Lgd, j + 1));
          a0 = combG(a0, t0, j - 1);
          a1 = combG(a1, t1, j);
          a2 = combG(a2, t2, j + 1);
          if (out) {
            if (mine(j - 1)) st2(AT(a.Mgd, j - 1), a0.x, a0.y);
            if (mine(j)) st2(AT(a.Mgd, j), a1.x, a1.y);
            if (mine(j + 1)) st2(AT(a.Mgd, j + 1), a2.x, a2.y);
          }
        }
        Um = combU(Um, tUm, j - 1);
        U0 = combU(U0, tU0, j);
        Vm = combV(Vm, tVm, j - 1);
        V0 = combV(V0, tV0, j);
        if (out) {
          if (mine(j - 1)) {
            st2(AT(a.MU, j - 1), Um.x, Um.y);
            st2(AT(a.MV, j - 1), Vm.x, Vm.y);
          }
          if (mine(j)) {
            st2(AT(a.MU, j), U0.x, U0.y);
            st2(AT(a.MV, j), V0.x, V0.y);
          }
        }
      }
    } else {
      const int sm1 = slot5(j - 1), s00 = next5(sm1), sp1 = next5(s00);
      a0 = rld(RIN, sm1, 2);
      a1 = rld(RIN, s00, 2);
      a2 = rld(RIN, sp1, 2);
      Um = rld(RIN, sm1, 0);
      U0 = rld(RIN, s00, 0);
      Vm = rld(RIN, sm1, 1);
      V0 = rld(RIN, s00, 1);
    }
    gr0 = a1;
    gr1 = a2;
    sm_.x = fast_sqrt(a0.x); sm_.y = fast_sqrt(a0.y);
    s0.x = fast_sqrt(a1.x); s0.y = fast_sqrt(a1.y);
    sp.x = fast_sqrt(a2.x); sp.y = fast_sqrt(a2.y);
    if (need_gh) {
      const D2 hs = ld2(AT(a.ghs, j));   // (rja)
      g0.x = a1.x + hs.x;
      g0.y = a1.y + hs.y;
    }
    const double s0e = shfl_dn1(s0.x);
    u0.x = two_a_over_b(U0.x, s0.x + s0.y);
    u0.y = two_a_over_b(U0.y, s0.y + s0e);
    const bool vmok = (j - 1 >= 0), v0ok = (j < nlat - 1);
    vm.x = vmok ? two_a_over_b(Vm.x, sm_.x + s0.x) : 0.0;
    vm.y = vmok ? two_a_over_b(Vm.y, sm_.y + s0.y) : 0.0;
    v0.x = v0ok ? two_a_over_b(V0.x, s0.x + sp.x) : 0.0;
    v0.y = v0ok ? two_a_over_b(V0.y, s0.y + sp.y) : 0.0;
  }
  double uw_a = shfl_up1(u0.y);    // u(i-1, j)   for column a
  double Uw_a = shfl_up1(U0.y);    // U(i-1, j)
  double Vse_b = shfl_dn1(Vm.x);   // V(i+1, j-1) for column b
  double vse_b = shfl_dn1(vm.x);   // v(i+1, j-1)
  double se_b = shfl_dn1(s0.x);    // s(i+1, j)
  // ROLE 0: the packets of the first PC_IN_DEPTH - 1 iterations are already in flight (issued before the prologue);
  // ROLE 1, 2: ghs is the only global operand left, prefetched one row ahead in registers
  D2 n_hs = zero2;
  if (ROLE != 0 && need_gh) n_hs = ld2(ATK(a.ghs, 1));

  int sj5 = slot5(rja), sj4 = slot4(rja);   // ring slots of row j
  GMD_UNROLL_PRAGMA(GMD_PC_UNROLL)
  for (int j = rja; j < rjb; j++) {
    const int sj5n = next5(sj5), sj5nn = next5(sj5n);   // ... of rows j+1, j+2
    D2 c_gd2, c_U, c_V, c_hs = zero2;
    if (ROLE == 0) {
      const int it = j - rja;
      const int st = it % PC_IN_DEPTH;
      mbar_wait(in_bar + 8u * (unsigned)st, (unsigned)(it / PC_IN_DEPTH) & 1u);
      const D2 *const pk = in_ring + st * (PC_IN_NF * 32);
      c_gd2 = ldl(pk + IN_GD * 32);
      c_U = ldl(pk + IN_U * 32);
      c_V = ldl(pk + IN_V * 32);
      if (need_gh) c_hs = ldl(pk + IN_HS * 32);
      if (LAZY) {
        if (LAZY == 1) {
          const D2 c_tg = ldl(pk + IN_TG * 32);
          c_gd2 = combG(c_gd2, c_tg, j + 2);
          if (out && mine(j + 2)) st2(AT(a.Mgd, j + 2), c_gd2.x, c_gd2.y);
        }
        const D2 c_tU = ldl(pk + IN_TU * 32), c_tV = ldl(pk + IN_TV * 32);
        c_U = combU(c_U, c_tU, j + 1);
        c_V = combV(c_V, c_tV, j + 1);
        if (out && mine(j + 1)) {
          st2(AT(a.MU, j + 1), c_U.x, c_U.y);
          st2(AT(a.MV, j + 1), c_V.x, c_V.y);
        }
      }
    } else {
      if (ROLE == 1) {   // tick t = j - rja + PC_D1: the packet of S1's iteration t + PC_IN_DEPTH - 1
        const int r = rja0 + (j - rja) + PC_D1 + (PC_IN_DEPTH - 1);
        if (r < rjb0) issue_packet(r, 1);
      }
      c_gd2 = rld(RIN, sj5nn, 2);
      c_U = rld(RIN, sj5n, 0);
      c_V = rld(RIN, sj5n, 1);
      c_hs = n_hs;
      if (need_gh && j + 1 < rjb) n_hs = ld2(AT(a.ghs, j + 2));
    }
    // what this row's result is combined with: the old state (ROLE 1), the previous tendency (ROLE 2)
    D2 qU = zero2, qV = zero2, qG = zero2;
    if (ROLE == 1) {
      qU = rld(RG_O, sj4, 0);
      qV = rld(RG_O, sj4, 1);
      qG = rld(RG_O, sj4, 2);
    } else if (ROLE == 2) {
      qU = rld(RG_P, sj4, 0);
      qV = rld(RG_P, sj4, 1);
      qG = rld(RG_P, sj4, 2);
    }
    const double *__restrict__ rc = recp(j);
    const double *__restrict__ rcm = recp(j - 1);
    const double *__restrict__ rcn = recp(j + 1);
    const bool rowU = (j >= 1 && j <= nlat - 2);
    const bool rowV = (j <= nlat - 2);
    const bool rowG = rowU && (PASS != PASS_SLOW);
    // ---- advance the window: s(j+2), U(j+1), V(j+1), gh(j+1), u(j+1), v(j+1) --------------------------------------
    sq.x = fast_sqrt(c_gd2.x);
    sq.y = fast_sqrt(c_gd2.y);
    Up = c_U;
    Vp = c_V;
    if (need_gh) {
      gp.x = gr1.x + c_hs.x;
      gp.y = gr1.y + c_hs.y;
    }
    const double spe_b = shfl_dn1(sp.x);  // s(i+1, j+1) for column b
    {
      const bool uok = (j + 1 < nlat), vok = (j + 1 < nlat - 1);
      up.x = uok ? two_a_over_b(Up.x, sp.x + sp.y) : 0.0;
      up.y = uok ? two_a_over_b(Up.y, sp.y + spe_b) : 0.0;
      vp.x = vok ? two_a_over_b(Vp.x, sp.x + sq.x) : 0.0;
      vp.y = vok ? two_a_over_b(Vp.y, sp.y + sq.y) : 0.0;
    }
    // ---- longitude neighbours of row j / j+1 ----------------------------------------------------------------------
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
      const double ge_b = shfl_dn1(g0.x);
      ghdx_a = g0.y - g0.x;
      ghdx_b = ge_b - g0.y;
      ghdy_a = gp.x - g0.x;
      ghdy_b = gp.y - g0.y;
    }
    const double hc0 = rcm[RC_COSH], hc1 = rc[RC_COSH], hc2 = rcn[RC_COSH];
    double dUa, dVa, dGa, dUb, dVb, dGb;
    tend_col<PASS, ADV>(rc, rcn, rowU, rowV, rowG, a.beta_lon, a.beta_lat, hc0, hc1, hc2,
                        uw_a, u0.x, u0.y, Uw_a, U0.x, U0.y, Vw_a, V0.x, V0.y, sw_a, s0.x, s0.y, v0.x, v0.y,
                        Um.x, Vm.x, Vm.y, vm.x, vm.y, sm_.x,
                        Up.x, Unw_a, up.x, unw_a, Vp.x, vp.x, sp.x,
                        ghdx_a, ghdy_a, 0.0, 0.0, 0.0, 0.0, dUa, dVa, dGa);
    tend_col<PASS, ADV>(rc, rcn, rowU, rowV, rowG, a.beta_lon, a.beta_lat, hc0, hc1, hc2,
                        u0.x, u0.y, ue_b, U0.x, U0.y, Ue_b, V0.x, V0.y, Ve_b, s0.x, s0.y, se_b, v0.y, ve_b,
                        Um.y, Vm.y, Vse_b, vm.y, vse_b, sm_.y,
                        Up.y, Up.x, up.y, up.x, Vp.y, vp.y, sp.y,
                        ghdx_b, ghdy_b, 0.0, 0.0, 0.0, 0.0, dUb, dVb, dGb);
    if (ROLE == 0) {
      // A(j) = old(j) + dt tend1(j); pole rows keep U (src/dycore_mod.F90:623-627); a slow pass carries gd along unchanged
      D2 nU = U0, nV = V0, nG = gr0;
      if (rowU) { nU.x = U0.x + a.dt * dUa; nU.y = U0.y + a.dt * dUb; }
      if (rowV) { nV.x = V0.x + a.dt * dVa; nV.y = V0.y + a.dt * dVb; }
      if (rowG) { nG.x = gr0.x + a.dt * dGa; nG.y = gr0.y + a.dt * dGb; }
      rst(RG_A, sj5, 0, nU);
      rst(RG_A, sj5, 1, nV);
      rst(RG_A, sj5, 2, nG);
      rst(RG_O, sj4, 0, U0);
      rst(RG_O, sj4, 1, V0);
      rst(RG_O, sj4, 2, gr0);
    } else if (ROLE == 1) {
      // B(j) = old(j) + dt tend2(j); tend2(j) goes on to S3a
      D2 nU = qU, nV = qV, nG = qG;
      if (rowU) { nU.x = qU.x + a.dt * dUa; nU.y = qU.y + a.dt * dUb; }
      if (rowV) { nV.x = qV.x + a.dt * dVa; nV.y = qV.y + a.dt * dVb; }
      if (rowG) { nG.x = qG.x + a.dt * dGa; nG.y = qG.y + a.dt * dGb; }
      rst(RG_B, sj5, 0, nU);
      rst(RG_B, sj5, 1, nV);
      rst(RG_B, sj5, 2, nG);
      D2 t;
      t.x = dUa; t.y = dUb;
      rst(RG_P, sj4, 0, t);
      t.x = dVa; t.y = dVb;
      rst(RG_P, sj4, 1, t);
      t.x = dGa; t.y = dGb;
      rst(RG_P, sj4, 2, t);
    } else {
      if (out) {
        const double cfj = rc[RC_COSF];
        if (rowU) {
          st2(AT(a.TU, j), dUa, dUb);
          ip1 = ip1 + dUa * qU.x * cfj;
          ip2 = ip2 + dUa * dUa * cfj;
          ip1 = ip1 + dUb * qU.y * cfj;
          ip2 = ip2 + dUb * dUb * cfj;
        }
        if (rowV) {
          st2(AT(a.TV, j), dVa, dVb);
          ip1 = ip1 + dVa * qV.x * hc1;
          ip2 = ip2 + dVa * dVa * hc1;
          ip1 = ip1 + dVb * qV.y * hc1;
          ip2 = ip2 + dVb * dVb * hc1;
        }
        if (rowG) {
          st2(AT(a.Tgd, j), dGa, dGb);
          ip1 = ip1 + dGa * qG.x * cfj;
          ip2 = ip2 + dGa * dGa * cfj;
          ip1 = ip1 + dGb * qG.y * cfj;
          ip2 = ip2 + dGb * dGb * cfj;
        }
        // band-edge rows of the new tendency straight into the neighbours' ghost rows (as k_stage PUSH)
        if (PUSH) {
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
      }
    }
    // ---- rotate the window ----------------------------------------------------------------------------------------
    sm_ = s0; s0 = sp; sp = sq;
    u0 = up;
    vm = v0; v0 = vp;
    Um = U0; U0 = Up;
    Vm = V0; V0 = Vp;
    g0 = gp;
    gr0 = gr1;
    gr1 = c_gd2;
    uw_a = unw_a;
    Uw_a = Unw_a;
    Vse_b = Ve_b;
    vse_b = ve_b;
    se_b = spe_b;
    off += nl;
    sj5 = sj5n;
    sj4 = (sj4 + 1) & (PC_DEPTH_OP - 1);
    pc_bar();
  }
#undef AT
#undef ATK
  for (int t = T0 + (rjb - rja); t < nticks; t++) pc_bar();
}

// grid (strips, row chunks); a.rb[0] / a.re[0]: the fused rows, a.rows_per_cta rows per chunk; a.medge[0] bit 0 / 1: the
// first / last chunk also materialises the deferred update on the band's southern / northern ghost rows
template <int PASS, int ADV, int LAZY, bool PUSH>
__global__ void __launch_bounds__(PC_BX, GMD_PC_MINB) k_pc(const StageArgs a) {
  extern __shared__ __align__(16) double pc_smem[];
  __shared__ double red[2];
  trace_in(a.tseq);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ja = a.rb[0] + (int)blockIdx.y * a.rows_per_cta;
  const int jb = min(ja + a.rows_per_cta, a.re[0]);
  const size_t pslot = (size_t)a.pofs[0] + (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  if (ja < jb) {   // (CTA-uniform)
    D2 *const ring = reinterpret_cast<D2 *>(pc_smem) + lane;
    if (threadIdx.x == 0) {
      const unsigned bar0 = (unsigned)__cvta_generic_to_shared(pc_smem) + (unsigned)(PC_RING_LINES + PC_IN_DEPTH * PC_IN_NF) * 512u +
                            (unsigned)(PC_REC_DEPTH * RC_N) * 8u;
      for (int q = 0; q < PC_IN_DEPTH; q++) mbar_init(bar0 + 8u * (unsigned)q, 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const bool edgeS = (ja == a.rb[0]) && (a.medge[0] & 1), edgeN = (jb == a.re[0]) && (a.medge[0] & 2);
    double ip1 = 0.0, ip2 = 0.0;
    if (warp == 0) pc_role<PASS, ADV, 0, LAZY, PUSH>(a, blockIdx.x, ja, jb, edgeS, edgeN, ring, ip1, ip2);
    else if (warp == 1) pc_role<PASS, ADV, 1, LAZY, PUSH>(a, blockIdx.x, ja, jb, edgeS, edgeN, ring, ip1, ip2);
    else {
      pc_role<PASS, ADV, 2, LAZY, PUSH>(a, blockIdx.x, ja, jb, edgeS, edgeN, ring, ip1, ip2);
      const double r1 = warp_sum(ip1), r2 = warp_sum(ip2);
      if (lane == 0) {
        red[0] = r1;
        red[1] = r2;
      }
    }
  } else if (threadIdx.x == 64) {
    red[0] = 0.0;
    red[1] = 0.0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a.partials[2 * pslot] = red[0];
    a.partials[2 * pslot + 1] = red[1];
    // this CTA's peer stores (issued by warp 2 before the barrier) must be visible to the neighbours before the ticket
    // below lets the reduction release them; CTAs that pushed nothing skip the fence
    if (PUSH && ja < jb && ((a.hpS_U != nullptr && ja < a.push_s_end) || (a.hpN_U != nullptr && jb > a.push_n_begin))) __threadfence_system();
  }
  if (a.fold.ticket)
    fold_tail<PC_BX>(a.fold.ticket, a.fold.total, a.partials, a.fold.n, a.fold.out, a.fold.r.page, a.fold.r.rank, a.fold.r.nranks,
                     a.fold.r.k, a.tseq);
  trace_out(a.tseq);
}

}  // namespace gmd

"""Independent NumPy restatement of the hot path, used ONLY to pin the C oracle (tests/).

Written from the numerical specification in SURVEY.md appendix A, vectorised, 0-based, with
``np.roll`` for the periodic longitude and explicit zero ghost rows -- i.e. structurally unlike
oracle/gmd_oracle.c (1-based loops over halo-padded arrays) so that a transcription slip in either
shows up as a disagreement.  The zonal filter uses ``numpy.fft`` instead of the FFTPACK restatement.

Reference anchors: src/dycore_mod.F90:184-792, src/types_mod.F90:347-426, src/filter_mod.F90:35-167,
src/diffusion_mod.F90:74-217, src/diag_mod.F90:42-121, src/mesh_mod.F90:44-114, src/data_mod.F90:26-47,
src/weno_mod.F90:69-300 (WENO advection), src/dycore_mod.F90:689-752 (isp splitting).
"""
from __future__ import annotations

import numpy as np

PI = 4.0 * np.arctan(1.0)
OMEGA = 2.0 * PI / 86400.0
RADIUS = 6.37122e6
G = 9.80616


def E(a):  # value at i+1
    return np.roll(a, -1, axis=-1)


def W(a):  # value at i-1
    return np.roll(a, 1, axis=-1)


class NpModel:
    def __init__(self, nlon, nlat, dt, subcycles=4, qcon_modified=True, split="csp2", adv="center_diff",
                 beta_lon=0.0, beta_lat=0.5, use_filter=True, cutoff=(), use_diffusion=False,
                 diffusion_order=2, diffusion_coef=0.0, time_scheme="predict_correct", time_order=3):
        self.time_scheme, self.time_order = time_scheme, time_order
        self.nlon, self.nlat, self.dt, self.S = nlon, nlat, float(dt), subcycles
        self.qcon, self.split, self.adv = qcon_modified, split, adv
        self.beta_lon, self.beta_lat = beta_lon, beta_lat
        self.use_diffusion, self.diffusion_order, self.nu = use_diffusion, diffusion_order, diffusion_coef
        nh = nlat - 1
        self.dlon = 2 * PI / nlon
        self.dlat = PI / nh
        flat = -0.5 * PI + np.arange(nlat) * self.dlat
        flat[-1] = 0.5 * PI
        hlat = flat[:-1] + 0.5 * self.dlat
        self.flat, self.hlat = flat, hlat
        self.cf = np.cos(flat)
        self.sf = np.sin(flat)
        self.ch = np.cos(hlat)
        self.cf[0] = self.cf[-1] = 0.0
        self.sf[0], self.sf[-1] = -1.0, 1.0
        self.f = 2.0 * OMEGA * self.sf
        with np.errstate(divide="ignore", invalid="ignore"):
            self.c = self.sf / self.cf / RADIUS
        self.c[0] = self.c[-1] = 0.0
        # reset_cos_lat_at_poles (dycore_mod.F90:159-173) happens before the first step
        self.cf[0] = self.ch[0] * 0.25
        self.cf[-1] = self.ch[-1] * 0.25
        self.dlon_f = RADIUS * self.dlon * self.cf
        self.dlat_f = RADIUS * self.dlat * self.cf
        self.dlon_h = RADIUS * self.dlon * self.ch
        self.dlat_h = RADIUS * self.dlat * self.ch
        # filter rows (filter_mod.F90:44-59,76-99), 0-based
        self.full_cut = {}
        self.half_cut = {}
        full_flag, half_flag, fm, hm = set(), set(), {}, {}
        if use_filter:
            for k, cw in enumerate(cutoff, start=1):
                if cw == 0:
                    continue
                full_flag.add(k)              # 1-based 1+k
                half_flag.add(k - 1)          # 1-based k
                full_flag.add(nlat - 1 - k)   # 1-based nlat-k
                if nlat - k + 1 <= nh:        # 1-based nlat-k+1, out of bounds for k=1
                    half_flag.add(nlat - k)
        for k, cw in enumerate(cutoff, start=1):
            if cw == 0:
                continue
            fm[k] = max(fm.get(k, -1), cw)
            hm[k - 1] = max(hm.get(k - 1, -1), cw)
            fm[nlat - 1 - k] = max(fm.get(nlat - 1 - k, -1), cw)
            hm[nh - k] = max(hm.get(nh - k, -1), cw)   # 1-based (nlat-1)-k+1
        self.full_cut = {j: fm.get(j, -1) for j in full_flag}
        self.half_cut = {h: hm.get(h, -1) for h in half_flag}
        self.ghs = np.zeros((nlat, nlon))
        self.beta = 1.0

    # ------------------------------------------------------------------ state helpers
    def iap(self, u, v, gd):
        s = np.sqrt(gd)
        U = 0.5 * (s + E(s)) * u
        V = 0.5 * (s[:-1] + s[1:]) * v
        return U, V, s

    @staticmethod
    def _padrows(a):
        z = np.zeros((1, a.shape[1]))
        return np.vstack([z, a, z])

    def filt(self, row, c):
        n = row.size
        X = np.fft.rfft(row)
        Y = np.zeros_like(X)
        if c >= 0:
            Y[: c + 1] = X[: c + 1]
            if c + 1 < X.size:
                Y[c + 1] = X[c + 1].real  # B3: only the cosine part of wavenumber c+1 survives
        return np.fft.irfft(Y, n)

    def smooth(self, d, w, cuts):
        for j, c in cuts.items():
            if j < 0 or j >= d.shape[0]:
                continue
            s1 = float(np.sum(d[j] * w[j]))
            if abs(s1) > 1.0e-16:
                r = self.filt(d[j], c)
                s2 = float(np.sum(r * w[j]))
                d[j] = r * s1 / s2

    # ------------------------------------------------------------------ operators
    def tend(self, st, pass_="all"):
        u, v, gd, U, V, s = st
        nlat = self.nlat
        ch, cf = self.ch, self.cf
        J = slice(1, nlat - 1)
        du = np.zeros_like(u)
        dv = np.zeros_like(v)
        dgd = np.zeros_like(gd)
        chp = np.concatenate([[0.0], ch, [0.0]])          # chp[h+1] = ch[h]
        vp, Vp = self._padrows(v), self._padrows(V)      # vp[h+1] = v[h]
        if pass_ in ("all", "slow"):
            if self.adv == "center_diff":
                ual = 0.25 / self.dlon_f[J, None] * ((u + E(u))[J] * E(U)[J] - (u + W(u))[J] * W(U)[J])
                u1 = W(u)[:-1] + W(u)[1:]
                u2 = u[:-1] + u[1:]
                val = 0.25 / self.dlon_h[:, None] * (u2 * E(V) - u1 * W(V))
                v1 = (v[:-1] + E(v)[:-1])[: nlat - 2] * ch[: nlat - 2, None]
                v2 = (v[1:] + E(v)[1:]) * ch[1:, None]
                uat = 0.25 / self.dlat_f[J, None] * (v2 * U[2:] - v1 * U[:-2])
                vc = v * ch[:, None]
                v1 = vc + vp[:-2] * chp[:-2, None]
                v2 = vc + vp[2:] * chp[2:, None]
                vat = 0.25 / self.dlat_h[:, None] * (v2 * Vp[2:] - v1 * Vp[:-2])
            elif self.adv == "upwind":
                bl, bt = self.beta_lon, self.beta_lat
                u1 = (u + W(u))[J]
                u2 = (u + E(u))[J]
                Uc, Ue, Uw = U[J], E(U)[J], W(U)[J]
                ual = 0.25 / self.dlon_f[J, None] * (u2 * (Uc + Ue) - bl * np.abs(u2) * (Ue - Uc) - u1 * (Uc + Uw)
                                                     + bl * np.abs(u1) * (Uc - Uw) - (u2 - u1) * Uc)
                u1 = W(u)[:-1] + W(u)[1:]
                u2 = u[:-1] + u[1:]
                val = 0.25 / self.dlon_h[:, None] * (u2 * (V + E(V)) - bl * np.abs(u2) * (E(V) - V) - u1 * (V + W(V))
                                                     + bl * np.abs(u1) * (V - W(V)) - (u2 - u1) * V)
                v1 = (v[:-1] + E(v)[:-1])[: nlat - 2] * ch[: nlat - 2, None]
                v2 = (v[1:] + E(v)[1:]) * ch[1:, None]
                Uc, Un, Us = U[J], U[2:], U[:-2]
                uat = 0.25 / self.dlat_f[J, None] * (v2 * (Uc + Un) - bt * np.abs(v2) * (Un - Uc) - v1 * (Uc + Us)
                                                     + bt * np.abs(v1) * (Uc - Us) - (v2 - v1) * Uc)
                vc = v * ch[:, None]
                v1 = vc + vp[:-2] * chp[:-2, None]
                v2 = vc + vp[2:] * chp[2:, None]
                Vn, Vs = Vp[2:], Vp[:-2]
                vat = 0.25 / self.dlat_h[:, None] * (v2 * (V + Vn) - bt * np.abs(v2) * (Vn - V) - v1 * (V + Vs)
                                                     + bt * np.abs(v1) * (V - Vs) - (v2 - v1) * V)
            elif self.adv == "weno":
                ual, val, uat, vat = self.weno_advection(u, v, U, V)
            else:
                raise NotImplementedError(self.adv)
            du[J] += -ual - uat
            dv += -val - vat
        if pass_ in ("all", "fast"):
            p = gd + self.ghs
            c1 = (ch[:-1] / cf[1:-1])[:, None]
            c2 = (ch[1:] / cf[1:-1])[:, None]
            fcu = self.f[:, None] + self.c[:, None] * u
            fv = 0.25 * fcu[J] * (c1 * (V[:-1] + E(V)[:-1]) + c2 * (V[1:] + E(V)[1:]))
            t = fcu * U
            fu = 0.25 * (t[:-1] + W(t)[:-1] + t[1:] + W(t)[1:])
            upgf = 0.5 * (s + E(s))[J] / self.dlon_f[J, None] * (E(gd) + E(self.ghs) - gd - self.ghs)[J]
            vpgf = 0.5 * (s[:-1] + s[1:]) / self.dlat_h[:, None] * ch[:, None] * (gd[1:] + self.ghs[1:] - gd[:-1] - self.ghs[:-1])
            mdl = (((s + E(s)) * U - (s + W(s)) * W(U)) * 0.5)[J] / self.dlon_f[J, None]
            fl = (s[:-1] + s[1:]) * V * ch[:, None]   # flux through half row h
            mdt = np.zeros_like(gd)
            mdt[J] = (fl[1:] - fl[:-1]) * 0.5 / self.dlat_f[J, None]
            mdt[0] = np.sum((s[0] + s[1]) * V[0]) * 2.0 / self.nlon / RADIUS / self.dlat
            mdt[-1] = -np.sum((s[-1] + s[-2]) * V[-1]) * 2.0 / self.nlon / RADIUS / self.dlat
            du[J] += fv - upgf
            dv += -fu - vpgf
            dgd[J] -= mdl
            dgd -= mdt
            del p
        self.smooth(du, U, {j: c for j, c in self.full_cut.items() if 1 <= j <= nlat - 2})
        self.smooth(dv, V, self.half_cut)
        if pass_ != "slow":
            self.smooth(dgd, gd + self.ghs, self.full_cut)
        return du, dv, dgd

    # ------------------------------------------------------------------ WENO advection (src/weno_mod.F90:69-300)
    @staticmethod
    def _weno2(fp1, fp2, fp3, fn2, fn3, fn4):
        """weno_2nd_order_pass (:235-298): two 2-point stencils per side, optimal weights 1/3, 2/3, eps 1e-6"""
        eps = 1.0e-6
        a1, a2 = (1.0 / 3.0) / (eps + (fp2 - fp1) ** 2) ** 2, (2.0 / 3.0) / (eps + (fp3 - fp2) ** 2) ** 2
        f = (a1 * (-0.5 * fp1 + 1.5 * fp2) + a2 * (0.5 * fp2 + 0.5 * fp3)) / (a1 + a2)
        b1, b2 = (1.0 / 3.0) / (eps + (fn3 - fn4) ** 2) ** 2, (2.0 / 3.0) / (eps + (fn2 - fn3) ** 2) ** 2
        return f + (b1 * (-0.5 * fn4 + 1.5 * fn3) + b2 * (0.5 * fn3 + 0.5 * fn2)) / (b1 + b2)

    def weno_advection(self, u, v, U, V):
        """returns (u_adv_lon[J], v_adv_lon, u_adv_lat[J], v_adv_lat); Lax-Friedrichs split with a fixed 20 m/s.
        Work-array rows the reference never writes (pole rows of the u-point arrays, latitude halos) read as 0."""
        nlat = self.nlat
        J = slice(1, nlat - 1)
        amax = 20.0
        zf, zh = np.zeros_like(U), np.zeros_like(V)

        def rows(a, shift, lo, hi):
            """a[row + shift] with zeros outside rows [lo, hi]"""
            out = np.zeros_like(a)
            n = a.shape[0]
            for r in range(n):
                q = r + shift
                if lo <= q <= hi:
                    out[r] = a[q]
            return out
        # ---- zonal (:69-158)
        fpu, fnu = zf.copy(), zf.copy()
        fpu[J] = 0.5 * (u[J] + amax) * U[J]
        fnu[J] = 0.5 * (u[J] - amax) * U[J]
        ub = 0.25 * (W(u)[:-1] + W(u)[1:] + u[:-1] + u[1:])
        fpv, fnv = 0.5 * (ub + amax) * V, 0.5 * (ub - amax) * V
        fu = zf.copy()
        fu[J] = self._weno2(W(fpu), fpu, E(fpu), fnu, E(fnu), E(E(fnu)))[J]
        fv = self._weno2(W(fpv), fpv, E(fpv), fnv, E(fnv), E(E(fnv)))
        # B8: the u-row metric is half_dlon(j), sic (:153-154)
        ual = ((fu - W(fu) - (E(u) - W(u)) * U * 0.25)[J]) / self.dlon_h[1: nlat - 1, None]
        val = (fv - W(fv) - (u[:-1] + u[1:] - W(u)[:-1] - W(u)[1:]) * V * 0.25) / self.dlon_h[:, None]
        # ---- meridional (:160-233)
        nh = nlat - 1
        vp = self._padrows(v)                      # vp[h + 1] = v[h]
        vb = 0.25 * (vp[:-1] + vp[1:] + E(vp)[:-1] + E(vp)[1:])     # at full row j: half rows j-1, j
        fpu, fnu = zf.copy(), zf.copy()
        fpu[J] = 0.5 * (vb[J] + amax) * U[J]
        fnu[J] = 0.5 * (vb[J] - amax) * U[J]
        fpv, fnv = 0.5 * (v + amax) * V, 0.5 * (v - amax) * V
        fu = zf.copy()
        fu[J] = self._weno2(rows(fpu, -1, 1, nlat - 2), fpu, rows(fpu, 1, 1, nlat - 2), fnu, rows(fnu, 1, 1, nlat - 2),
                            rows(fnu, 2, 1, nlat - 2))[J]
        fv = self._weno2(rows(fpv, -1, 0, nh - 1), fpv, rows(fpv, 1, 0, nh - 1), fnv, rows(fnv, 1, 0, nh - 1),
                         rows(fnv, 2, 0, nh - 1))
        dv_u = (W(vp)[1:] + vp[1:] - W(vp)[:-1] - vp[:-1])          # at full row j: v(i-1,j)+v(i,j)-v(i-1,j-1)-v(i,j-1)
        uat = ((fu - rows(fu, -1, 1, nlat - 2) - dv_u * U * 0.25)[J]) / self.dlat_f[J, None]
        vat = (fv - rows(fv, -1, 0, nh - 1) - (vp[2:] - vp[:-2]) * V * 0.25) / self.dlat_h[:, None]
        return ual, val, uat, vat

    def update(self, dt, td, old):
        du, dv, dgd = td
        u0, v0, gd0, U0, V0, s0 = old
        gd = gd0 + dt * dgd
        s = np.sqrt(gd)
        U = U0 + dt * du
        V = V0 + dt * dv
        u = u0.copy()
        u[1:-1] = (U * 2.0 / (s + E(s)))[1:-1]
        v = V * 2.0 / (s[:-1] + s[1:])
        return u, v, gd, U, V, s

    def inner(self, a, b):
        return (np.sum(a[0][1:-1] * b[0][1:-1] * self.cf[1:-1, None]) + np.sum(a[1] * b[1] * self.ch[:, None])
                + np.sum(a[2] * b[2] * self.cf[:, None]))

    def predict_correct(self, dt, old, pass_):
        t_old = self.tend(old, pass_)
        new = self.update(0.5 * dt, t_old, old)
        t_old = self.tend(new, pass_)
        new = self.update(0.5 * dt, t_old, old)
        t_new = self.tend(new, pass_)
        ip1, ip2 = self.inner(t_old, t_new), self.inner(t_new, t_new)
        self.beta = ip1 / ip2 if (self.qcon and ip1 != 0.0 and ip2 != 0.0) else 1.0
        return self.update(dt * self.beta, t_new, old)

    def runge_kutta(self, dt, old, pass_):
        """the specified runge_kutta integrator (oracle/gmd_oracle.c runge_kutta), second reading:
        increment form with the energy fix beta = -2 <K, phi> / (dt <K, K>), <K, phi> written with the tendency products
        the stage antisymmetries <L(psi), psi> = 0 turn it into"""
        k1 = self.tend(old, pass_)
        if self.time_order == 4:
            k2 = self.tend(self.update(0.5 * dt, k1, old), pass_)
            k3 = self.tend(self.update(0.5 * dt, k2, old), pass_)
            k4 = self.tend(self.update(dt, k3, old), pass_)
            K = tuple((1.0 / 6.0) * d + (1.0 / 6.0) * ((a + 2.0 * b) + 2.0 * c) for a, b, c, d in zip(k1, k2, k3, k4))
            ip1 = self.inner(k1, k2) + self.inner(k2, k3) + self.inner(k3, k4)
        else:
            k2 = self.tend(self.update(dt, k1, old), pass_)
            k12 = tuple(a + b for a, b in zip(k1, k2))
            k3 = self.tend(self.update(0.25 * dt, k12, old), pass_)
            K = tuple((2.0 / 3.0) * c + (1.0 / 6.0) * ab for ab, c in zip(k12, k3))
            ip1 = self.inner(k1, k2) + self.inner(k12, k3)
        ip2 = self.inner(K, K)
        self.beta = ip1 / (3.0 * ip2) if (self.qcon and ip1 != 0.0 and ip2 != 0.0) else 1.0
        self.beta_direct = -2.0 * self.inner_state(K, old) / (dt * ip2)   # the ill-conditioned form, for the tests
        return self.update(dt * self.beta, K, old)

    def integrator(self, dt, old, pass_):
        return self.runge_kutta(dt, old, pass_) if self.time_scheme == "runge_kutta" else self.predict_correct(dt, old, pass_)

    def inner_state(self, t, st):
        """inner_product_tend_state (src/types_mod.F90:373-397): (du, U), (dv, V), (dgd, gd)"""
        return self.inner(t, (st[3], st[4], st[2]))

    def isp(self, F):
        """isp_splitting (src/dycore_mod.F90:689-752)"""
        S, dt = self.S, self.dt
        fast_dt, half_dt = dt / S, dt * 0.5
        add = lambda a, b: tuple(x + y for x, y in zip(a, b))
        slow = self.tend(F, "slow")
        acc = tuple(np.zeros_like(x) for x in slow)
        P = F
        for _ in range(S):
            t = add(self.tend(P, "fast"), slow)
            P1 = self.update(fast_dt * 0.5, t, P)
            t = add(self.tend(P1, "fast"), slow)
            P2 = self.update(fast_dt * 0.5, t, P)
            t2 = self.tend(P2, "fast")
            acc = add(acc, t2)
            P = self.update(fast_dt, add(t2, slow), P)
        acc = tuple(x * (2.0 / S) for x in acc)
        sub = lambda a, b: tuple(x - y for x, y in zip(a, b))
        Q1 = self.update(half_dt, sub(self.tend(P, "slow"), slow), P)
        Q2 = self.update(half_dt, sub(self.tend(Q1, "slow"), slow), P)
        R = add(add(self.tend(Q2, "slow"), slow), acc)
        ip1, ip2 = self.inner_state(R, F), self.inner(R, R)
        beta = ip1 / ip2 if (self.qcon and ip1 != 0.0 and ip2 != 0.0) else 1.0
        self.beta = beta * 4.0 / dt
        return self.update(half_dt * self.beta, R, F)

    def diffusion(self, dt, st):
        """ordinary_diffusion (src/diffusion_mod.F90:74-217), order 2 or 4"""
        u, v, gd, U, V, s = st
        if self.diffusion_order == 2:
            gdd, ud, vd = self._laplace(gd, u, v)
            sign = 1.0
        else:
            assert self.diffusion_order == 4
            g1, u1, v1 = self._laplace(gd, u, v)
            gdd, ud, vd = self._laplace(g1, u1, v1)   # the work copies become the first Laplacian (:172-179)
            sign = -1.0
        nlat = self.nlat
        for j, c in self.full_cut.items():
            if 1 <= j <= nlat - 2:
                gdd[j] = self.filt(gdd[j], c)
                ud[j] = self.filt(ud[j], c)
        for h, c in self.half_cut.items():
            vd[h] = self.filt(vd[h], c)
        gd = gd + sign * dt * self.nu * gdd
        u = u + sign * dt * self.nu * ud
        v = v + sign * dt * self.nu * vd
        U, V, s = self.iap(u, v, gd)
        return u, v, gd, U, V, s

    def _laplace(self, gd, u, v):
        """one scalar-Laplacian pass on gd, u, v (src/diffusion_mod.F90:107-171)"""
        nlat, nlon = self.nlat, self.nlon
        ch, cf = self.ch, self.cf
        J = slice(1, nlat - 1)

        def lap_full(q):
            out = np.zeros_like(q)
            out[J] = ((E(q) - 2 * q + W(q))[J] / self.dlon_f[J, None] ** 2
                      + ((q[2:] - q[1:-1]) * ch[1:, None] - (q[1:-1] - q[:-2]) * ch[:-1, None]) / self.dlat_f[J, None] ** 2 * cf[J, None])
            return out
        gdd = lap_full(gd)
        gdd[0] = np.sum(gd[1] - gd[0]) * ch[0] / self.dlat_f[0] ** 2 * cf[0] / nlon
        gdd[-1] = -np.sum(gd[-1] - gd[-2]) * ch[-1] / self.dlat_f[-1] ** 2 * cf[-1] / nlon
        ud = lap_full(u)
        vd = (E(v) - 2 * v + W(v)) / self.dlon_h[:, None] ** 2
        vd[1:-1] += ((v[2:] - v[1:-1]) * cf[2:-1, None] - (v[1:-1] - v[:-2]) * cf[1:-2, None]) / self.dlat_h[1:-1, None] ** 2 * ch[1:-1, None]
        vd[0] += (v[1] - v[0]) * cf[1] / self.dlat_h[0] ** 2 * ch[0]
        vd[-1] -= (v[-1] - v[-2]) * cf[-2] / self.dlat_h[-1] ** 2 * ch[-1]
        return gdd, ud, vd

    def step(self, st):
        if self.split == "isp":
            new = self.isp(st)
        elif self.split == "csp2":
            st1 = self.integrator(0.5 * self.dt, st, "slow")
            for _ in range(self.S):
                st1 = self.integrator(self.dt / self.S, st1, "fast")
            new = self.integrator(0.5 * self.dt, st1, "slow")
        else:
            new = self.integrator(self.dt, st, "all")
        if self.use_diffusion:
            new = self.diffusion(self.dt, new)
        return new

    def make_state(self, u, v, gd):
        U, V, s = self.iap(u, v, gd)
        return u.copy(), v.copy(), gd.copy(), U, V, s

    def mass_energy(self, st):
        u, v, gd, U, V, s = st
        mass = np.sum(self.cf[:, None] * self.dlon * self.dlat * gd) * RADIUS ** 2
        en = (np.sum(U[1:-1] ** 2 * self.cf[1:-1, None]) + np.sum(V ** 2 * self.ch[:, None])
              + np.sum((gd + self.ghs) ** 2 * self.cf[:, None]))
        return mass, en

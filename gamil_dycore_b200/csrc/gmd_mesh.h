// gmd_mesh.h -- host-side construction of the per-latitude coefficient tables and the filter row map.
//
// Follows, in the same expression order (so the tables are bit-identical to a gfortran build without
// -ffast-math): mesh_init (src/mesh_mod.F90:44-114), data_init (src/data_mod.F90:26-47),
// reset_cos_lat_at_poles (src/dycore_mod.F90:159-173) and filter_init (src/filter_mod.F90:35-103).
// Rows are 0-based here: full rows 0..nlat-1 (0 and nlat-1 are the poles), half row h lies between
// full rows h and h+1 (0..nlat-2).
#pragma once
#include <cmath>
#include <algorithm>
#include <vector>

namespace gmd {

constexpr int TPAD = 4;  // zero/benign padding entries on both ends of every device table

struct HostMesh {
  int nlon = 0, nlat = 0;
  double pi, omega, radius, g, dlon, dlat;
  // index with [j + TPAD]
  std::vector<double> full_lat, half_lat, full_cos, half_cos, full_sin, half_sin;
  std::vector<double> full_f, full_c, full_dlon, half_dlon, full_dlat, half_dlat;
  // filter map
  std::vector<int> flag_full, flag_half, cut_full, cut_half;  // [nlat], [nlat] (half uses nlat-1)
  int cutoff_max = -1;
  // moving reduced tendency: zonal reduction factor per full / half row, 0 = not reduced
  std::vector<int> red_full, red_half;

  double &at(std::vector<double> &v, int j) { return v[(size_t)(j + TPAD)]; }
  double at(const std::vector<double> &v, int j) const { return v[(size_t)(j + TPAD)]; }

  void init(int nlon_, int nlat_, bool reset_poles) {
    nlon = nlon_;
    nlat = nlat_;
    const int nhalf = nlat - 1;
    pi = std::atan(1.0) * 4.0;            // params_mod.F90:6
    omega = 2.0 * pi / 86400.0;           // :8
    radius = 6.37122e6;                   // :9
    g = 9.80616;                          // :10
    const size_t n = (size_t)nlat + 2 * TPAD;
    for (auto *v : {&full_lat, &half_lat, &full_cos, &half_cos, &full_sin, &half_sin, &full_f, &full_c,
                    &full_dlon, &half_dlon, &full_dlat, &half_dlat})
      v->assign(n, 0.0);
    dlon = 2 * pi / nlon;                 // mesh_mod.F90:64
    dlat = pi / nhalf;                    // :73
    for (int j = 0; j < nhalf; j++) {     // :74-77
      at(full_lat, j) = -0.5 * pi + j * dlat;
      at(half_lat, j) = at(full_lat, j) + 0.5 * dlat;
    }
    at(full_lat, nlat - 1) = 0.5 * pi;    // :78
    for (int j = 0; j < nhalf; j++) {     // :92-98
      at(half_cos, j) = std::cos(at(half_lat, j));
      at(half_sin, j) = std::sin(at(half_lat, j));
    }
    for (int j = 0; j < nlat; j++) {      // :100-106
      at(full_cos, j) = std::cos(at(full_lat, j));
      at(full_sin, j) = std::sin(at(full_lat, j));
    }
    at(full_cos, 0) = 0.0;                // :107-110 pole overrides
    at(full_cos, nlat - 1) = 0.0;
    at(full_sin, 0) = -1.0;
    at(full_sin, nlat - 1) = 1.0;
    for (int j = 0; j < nlat; j++) {      // data_mod.F90:31-38
      at(full_f, j) = 2.0 * omega * at(full_sin, j);
      at(full_c, j) = (j == 0 || j == nlat - 1) ? 0.0 : at(full_sin, j) / at(full_cos, j) / radius;
      at(full_dlon, j) = radius * dlon * at(full_cos, j);
      at(full_dlat, j) = radius * dlat * at(full_cos, j);
    }
    for (int j = 0; j < nhalf; j++) {     // data_mod.F90:40-46
      at(half_dlon, j) = radius * dlon * at(half_cos, j);
      at(half_dlat, j) = radius * dlat * at(half_cos, j);
    }
    if (reset_poles) {                    // dycore_mod.F90:159-173
      at(full_cos, 0) = at(half_cos, 0) * 0.25;
      at(full_dlon, 0) = radius * dlon * at(full_cos, 0);
      at(full_dlat, 0) = radius * dlat * at(full_cos, 0);
      at(full_cos, nlat - 1) = at(half_cos, nhalf - 1) * 0.25;
      at(full_dlon, nlat - 1) = radius * dlon * at(full_cos, nlat - 1);
      at(full_dlat, nlat - 1) = radius * dlat * at(full_cos, nlat - 1);
    }
  }

  // filter_init, src/filter_mod.F90:35-103.  cw = zonal_tend_filter_cutoff_wavenumber(1:20).
  // South: full row 1+k, half row k (1-based) <- c_k.  North: full row nlat-k <- c_k; half FLAG at
  // nlat-k+1 but half MASK at nlat-k (reproduced; the k=1 flag that the
  // reference writes one element past the array is dropped).  A row hit twice keeps the larger cutoff.
  void filter_init(bool use_filter, const int *cw) {
    const int nhalf = nlat - 1;
    flag_full.assign((size_t)nlat, 0);
    flag_half.assign((size_t)nlat, 0);
    cut_full.assign((size_t)nlat, -1);
    cut_half.assign((size_t)nlat, -1);
    cutoff_max = -1;
    for (int k = 1; k <= 20; k++) {
      const int c = cw[k - 1];
      if (c == 0) continue;
      // masks are built regardless of use_zonal_tend_filter (filter_mod.F90:61-99)
      if (1 + k <= nlat && c > cut_full[(size_t)k]) cut_full[(size_t)k] = c;                    // full row 1+k -> 0-based k
      if (k <= nhalf && c > cut_half[(size_t)(k - 1)]) cut_half[(size_t)(k - 1)] = c;
      if (nlat - k >= 1 && c > cut_full[(size_t)(nlat - k - 1)]) cut_full[(size_t)(nlat - k - 1)] = c;
      if (nhalf - k + 1 >= 1 && c > cut_half[(size_t)(nhalf - k)]) cut_half[(size_t)(nhalf - k)] = c;
      if (c > cutoff_max) cutoff_max = c;
      if (!use_filter) continue;
      if (1 + k <= nlat) flag_full[(size_t)k] = 1;
      if (k <= nhalf) flag_half[(size_t)(k - 1)] = 1;
      if (nlat - k >= 1) flag_full[(size_t)(nlat - k - 1)] = 1;
      if (nlat - k + 1 >= 1 && nlat - k + 1 <= nhalf) flag_half[(size_t)(nlat - k)] = 1;
    }
  }

  // Row map of the moving reduced tendency: factor r_k of zonal_reduce_factors(k) applies to the k-th full row next to
  // each pole (the pole row itself has no du and a zonally uniform dgd) and to the k-th half row from each pole.
  // Returns 0, or the 1-based k of the first factor that is negative or does not divide nlon.
  int reduce_init(bool use_reduce, const int *rf) {
    const int nhalf = nlat - 1;
    red_full.assign((size_t)nlat, 0);
    red_half.assign((size_t)nlat, 0);
    if (!use_reduce) return 0;
    for (int k = 1; k <= 20; k++) {
      const int r = rf[k - 1];
      if (r < 0 || (r > 1 && nlon % r != 0)) return k;
      if (r <= 1) continue;
      if (k <= nlat - 2) red_full[(size_t)k] = std::max(red_full[(size_t)k], r);                       // south, 0-based row k
      if (nlat - 1 - k >= 1) red_full[(size_t)(nlat - 1 - k)] = std::max(red_full[(size_t)(nlat - 1 - k)], r);
      if (k - 1 < nhalf) red_half[(size_t)(k - 1)] = std::max(red_half[(size_t)(k - 1)], r);
      if (nhalf - k >= 0) red_half[(size_t)(nhalf - k)] = std::max(red_half[(size_t)(nhalf - k)], r);
    }
    return 0;
  }
};

}  // namespace gmd

// history.h -- the `h0` history dataset of src/history_mod.F90:34-51,93-124 written through netcdf_classic.h.
// Schema (names, order, attributes, dimensions, NC_DOUBLE everywhere) follows the reference; the DATA are
// well defined where the reference reads out of bounds: v is the A-grid average on all nlat rows,
// vor carries a zero last row, div drops the (zero) north-pole row.
#pragma once
#include <string>
#include <vector>

#include "netcdf_classic.h"
#include "params.h"
#include "time_manager.h"

namespace host {

inline std::string pad_to(const std::string &s, size_t n) {
  std::string r = s.substr(0, n);
  r.resize(n, ' ');
  return r;
}

// u [nlat][nlon], v [nlat-1][nlon], gd, ghs [nlat][nlon], vor [nlat-1][nlon], div [nlat][nlon] (compact, C grid)
inline bool history_write(const Params &p, const TimeManager &tm, const std::vector<double> &u,
                          const std::vector<double> &v, const std::vector<double> &gd, const std::vector<double> &ghs,
                          const std::vector<double> &vor, const std::vector<double> &div, double total_energy,
                          double total_mass, std::string &path_out, std::string &err) {
  const int nlon = p.num_lon, nlat = p.num_lat;
  const double pi = std::atan(1.0) * 4.0, rad_to_deg = 180.0 / pi;
  NcFile nc;
  // global attributes, insertion order of src/io_mod.F90:427-429 + history_mod.F90:35-38
  nc.add_global(NcFile::att_text("dataset", pad_to("hist0", 30)));
  nc.add_global(NcFile::att_text("desc", pad_to(p.case_desc, 256)));
  nc.add_global(NcFile::att_text("author", pad_to("N/A", 256)));  // the namelist `author` is never copied (io_mod.F90:30)
  nc.add_global(NcFile::att_double("time_step_size", p.time_step_size));
  nc.add_global(NcFile::att_text("time_scheme", pad_to(p.time_scheme, 30)));
  nc.add_global(NcFile::att_text("split_scheme", pad_to(p.split_scheme, 30)));
  nc.add_global(NcFile::att_int("subcycles", p.subcycles));
  const int d_time = nc.add_dim("time", 0), d_lon = nc.add_dim("lon", (size_t)nlon), d_lat = nc.add_dim("lat", (size_t)nlat);
  const int d_ilon = nc.add_dim("ilon", (size_t)nlon), d_ilat = nc.add_dim("ilat", (size_t)(nlat - 1));
  auto atts = [](const std::string &ln, const std::string &un) {
    return std::vector<NcFile::Att>{NcFile::att_text("long_name", ln), NcFile::att_text("units", un)};
  };
  double tu = 86400.0;                                      // io_init, src/io_mod.F90:86-97
  if (p.time_units == "hours") tu = 3600.0;
  else if (p.time_units == "seconds") tu = 60.0;            // sic (B7)
  const int v_time = nc.add_var("time", {d_time}, atts("Time", p.time_units + " since " + tm.start_time_format));
  const int v_lon = nc.add_var("lon", {d_lon}, atts("Longitude", "degrees_east"));
  const int v_lat = nc.add_var("lat", {d_lat}, atts("Latitude", "degrees_north"));
  const int v_ilon = nc.add_var("ilon", {d_ilon}, atts("Longitude", "degrees_east"));
  const int v_ilat = nc.add_var("ilat", {d_ilat}, atts("Latitude", "degrees_north"));
  const int v_u = nc.add_var("u", {d_time, d_lat, d_lon}, atts("u wind component", "m s-1"));
  const int v_v = nc.add_var("v", {d_time, d_lat, d_lon}, atts("v wind component", "m s-1"));
  const int v_gh = nc.add_var("gh", {d_time, d_lat, d_lon}, atts("geopotential height", "m2 s-2"));
  const int v_ghs = nc.add_var("ghs", {d_time, d_lat, d_lon}, atts("surface geopotential", "m2 s-2"));
  const int v_vor = nc.add_var("vor", {d_time, d_lat, d_ilon}, atts("relative vorticity", "s-1"));
  const int v_div = nc.add_var("div", {d_time, d_ilat, d_lon}, atts("divergence", "s-1"));
  const int v_te = nc.add_var("te", {d_time}, atts("total energy", "m4 s-4"));
  const int v_tm = nc.add_var("tm", {d_time}, atts("total mass", "m2 s-2"));

  const double dlon = 2 * pi / nlon, dlat = pi / (nlat - 1);
  std::vector<double> lon((size_t)nlon), ilon((size_t)nlon), lat((size_t)nlat), ilat((size_t)nlat - 1);
  for (int i = 0; i < nlon; i++) { lon[i] = i * dlon * rad_to_deg; ilon[i] = (i * dlon + 0.5 * dlon) * rad_to_deg; }
  for (int j = 0; j < nlat - 1; j++) {
    const double fl = -0.5 * pi + j * dlat;
    lat[j] = fl * rad_to_deg;
    ilat[j] = (fl + 0.5 * dlat) * rad_to_deg;
  }
  lat[nlat - 1] = 0.5 * pi * rad_to_deg;
  // C grid -> A grid (src/history_mod.F90:102-108); v(i,0) = v(i,nlat) = 0 are the zero latitude halos
  const size_t nf = (size_t)nlon * nlat;
  std::vector<double> ua(nf), va(nf), gh(nf), vorp(nf, 0.0), divp((size_t)nlon * (nlat - 1));
  for (int j = 0; j < nlat; j++)
    for (int i = 0; i < nlon; i++) {
      const size_t k = (size_t)j * nlon + i;
      const int iw = (i == 0) ? nlon - 1 : i - 1;
      ua[k] = 0.5 * (u[k] + u[(size_t)j * nlon + iw]);
      const double vn = (j <= nlat - 2) ? v[k] : 0.0, vs = (j >= 1) ? v[k - nlon] : 0.0;
      va[k] = 0.5 * (vn + vs);
      gh[k] = gd[k] + ghs[k];
    }
  for (size_t k = 0; k < (size_t)nlon * (nlat - 1); k++) { vorp[k] = vor[k]; divp[k] = div[k]; }
  nc.set_data(v_time, std::vector<double>{tm.elapsed_seconds / tu});
  nc.set_data(v_lon, lon); nc.set_data(v_lat, lat); nc.set_data(v_ilon, ilon); nc.set_data(v_ilat, ilat);
  nc.set_data(v_u, ua); nc.set_data(v_v, va); nc.set_data(v_gh, gh); nc.set_data(v_ghs, ghs);
  nc.set_data(v_vor, vorp); nc.set_data(v_div, divp);
  nc.set_data(v_te, std::vector<double>{total_energy});
  nc.set_data(v_tm, std::vector<double>{total_mass});
  path_out = p.case_name + ".h0." + tm.curr_time_format + ".nc";  // src/io_mod.F90:417-419
  return nc.write(path_out, err);
}

}  // namespace host

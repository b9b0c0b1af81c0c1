// restart.h -- the `restart` dataset of src/restart_mod.F90:22-75: one classic-netCDF file per restart alert,
// <case_name>.r.<curr_time_format>.nc, holding u, v, gd, ghs on the C grid plus the global attributes
// restart_time / elapsed_seconds, and the reader behind dycore_restart (src/dycore_mod.F90:113-117).
//
// Schema (names, order, attributes, dimensions) follows the reference.  The DATA differ on purpose: the reference
// hands the halo-padded arrays state%u(:,:) to a variable of interior size (src/io_mod.F90:577-621 with
// src/restart_mod.F90:67-70), so what it writes is shifted by the two halo cells and what it reads back
// (io_input_real_2d, src/io_mod.F90:687-724, again with halo-padded bounds) is not the state it wrote
// (SURVEY.md section 5, quirk B10).  Here the interior is written and read (bit-exact round trip), so a restarted
// run reproduces the uninterrupted one to rounding -- like the reference's, a restart goes through u, v, gd and
// iap_transform (tests/test_gpu_parity.py::test_restart_run_reproduces_the_continuous_run).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "history.h"
#include "netcdf_classic.h"
#include "params.h"
#include "time_manager.h"

namespace host {

inline bool restart_write(const Params &p, const TimeManager &tm, const std::vector<double> &u,
                          const std::vector<double> &v, const std::vector<double> &gd, const std::vector<double> &ghs,
                          std::string &path_out, std::string &err) {
  const int nlon = p.num_lon, nlat = p.num_lat;
  const double pi = std::atan(1.0) * 4.0, rad_to_deg = 180.0 / pi;
  NcFile nc;
  nc.add_global(NcFile::att_text("dataset", pad_to("restart", 30)));
  nc.add_global(NcFile::att_text("desc", pad_to("Restart file", 256)));
  nc.add_global(NcFile::att_text("author", pad_to("N/A", 256)));
  nc.add_global(NcFile::att_text("restart_time", pad_to(tm.curr_time_format, 30)));   // src/restart_mod.F90:62
  nc.add_global(NcFile::att_double("elapsed_seconds", tm.elapsed_seconds));            // :63
  const int d_time = nc.add_dim("time", 0), d_lon = nc.add_dim("lon", (size_t)nlon), d_lat = nc.add_dim("lat", (size_t)nlat);
  const int d_ilon = nc.add_dim("ilon", (size_t)nlon), d_ilat = nc.add_dim("ilat", (size_t)(nlat - 1));
  auto atts = [](const std::string &ln, const std::string &un) {
    return std::vector<NcFile::Att>{NcFile::att_text("long_name", ln), NcFile::att_text("units", un)};
  };
  double tu = 86400.0;
  if (p.time_units == "hours") tu = 3600.0;
  else if (p.time_units == "seconds") tu = 60.0;  // sic (B7)
  const int v_time = nc.add_var("time", {d_time}, atts("Time", p.time_units + " since " + tm.start_time_format));
  const int v_lon = nc.add_var("lon", {d_lon}, atts("Longitude", "degrees_east"));
  const int v_lat = nc.add_var("lat", {d_lat}, atts("Latitude", "degrees_north"));
  const int v_ilon = nc.add_var("ilon", {d_ilon}, atts("Longitude", "degrees_east"));
  const int v_ilat = nc.add_var("ilat", {d_ilat}, atts("Latitude", "degrees_north"));
  const int v_u = nc.add_var("u", {d_time, d_lat, d_ilon}, atts("u wind component", "m s-1"));       // :31
  const int v_v = nc.add_var("v", {d_time, d_ilat, d_lon}, atts("v wind component", "m s-1"));       // :32
  const int v_gd = nc.add_var("gd", {d_time, d_lat, d_lon}, atts("geopotential depth", "m2 s-2"));   // :33
  const int v_ghs = nc.add_var("ghs", {d_time, d_lat, d_lon}, atts("surface geopotential", "m2 s-2"));
  const double dlon = 2 * pi / nlon, dlat = pi / (nlat - 1);
  std::vector<double> lon((size_t)nlon), ilon((size_t)nlon), lat((size_t)nlat), ilat((size_t)nlat - 1);
  for (int i = 0; i < nlon; i++) { lon[i] = i * dlon * rad_to_deg; ilon[i] = (i * dlon + 0.5 * dlon) * rad_to_deg; }
  for (int j = 0; j < nlat - 1; j++) {
    const double fl = -0.5 * pi + j * dlat;
    lat[j] = fl * rad_to_deg;
    ilat[j] = (fl + 0.5 * dlat) * rad_to_deg;
  }
  lat[nlat - 1] = 0.5 * pi * rad_to_deg;
  nc.set_data(v_time, std::vector<double>{tm.elapsed_seconds / tu});
  nc.set_data(v_lon, lon); nc.set_data(v_lat, lat); nc.set_data(v_ilon, ilon); nc.set_data(v_ilat, ilat);
  nc.set_data(v_u, u); nc.set_data(v_v, v); nc.set_data(v_gd, gd); nc.set_data(v_ghs, ghs);
  path_out = p.case_name + ".r." + tm.curr_time_format + ".nc";  // src/restart_mod.F90:24 + io_mod.F90:419
  return nc.write(path_out, err);
}

// ---- a reader for what restart_write (or netCDF's NF90_CREATE(NF90_CLOBBER)) produces: CDF-1 / CDF-2 headers,
//      NC_DOUBLE variables, the last record of the record variables ----------------------------------------------
class NcReader {
 public:
  bool open(const std::string &path, std::string &err) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { err = "io_create_dataset: Input file \"" + path + "\" does not exist!"; return false; }  // io_mod.F90:164-167
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf_.resize((size_t)std::max(0L, n));
    const bool ok = n > 0 && fread(buf_.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    if (!ok) { err = "cannot read " + path; return false; }
    return parse(err);
  }
  bool get_text_att(const std::string &name, std::string &out) const {
    auto it = gtext_.find(name);
    if (it == gtext_.end()) return false;
    out = it->second;
    while (!out.empty() && (out.back() == ' ' || out.back() == '\0')) out.pop_back();
    return true;
  }
  bool get_double_att(const std::string &name, double &out) const {
    auto it = gdouble_.find(name);
    if (it == gdouble_.end()) return false;
    out = it->second;
    return true;
  }
  // variable `name` (NC_DOUBLE) of exactly n values per record (or in total, for a fixed-size variable)
  bool get_var(const std::string &name, size_t n, std::vector<double> &out, std::string &err) const {
    auto it = vars_.find(name);
    if (it == vars_.end()) { err = "No variable \"" + name + "\" in dataset!"; return false; }  // io_mod.F90:711-713
    const VarInfo &v = it->second;
    if (v.type != 6) { err = "variable " + name + " is not NC_DOUBLE"; return false; }
    if (v.nvals != n) { err = "variable " + name + " has " + std::to_string(v.nvals) + " values, expected " + std::to_string(n); return false; }
    uint64_t off = v.begin;
    if (v.isrec) {
      if (numrecs_ == 0) { err = "no record in the restart file"; return false; }
      off += (uint64_t)(numrecs_ - 1) * recsize_;
    }
    if (off + n * 8 > buf_.size()) { err = "restart file is truncated"; return false; }
    out.resize(n);
    for (size_t k = 0; k < n; k++) {
      uint64_t w = 0;
      for (int b = 0; b < 8; b++) w = (w << 8) | buf_[(size_t)off + k * 8 + (size_t)b];
      memcpy(&out[k], &w, 8);
    }
    return true;
  }

 private:
  struct VarInfo { int type; size_t nvals; uint64_t begin, vsize; bool isrec; };
  std::vector<unsigned char> buf_;
  size_t pos_ = 0;
  uint32_t numrecs_ = 0;
  uint64_t recsize_ = 0;
  std::map<std::string, std::string> gtext_;
  std::map<std::string, double> gdouble_;
  std::map<std::string, VarInfo> vars_;

  bool need(size_t n) const { return pos_ + n <= buf_.size(); }
  uint32_t u32() { uint32_t v = 0; for (int b = 0; b < 4; b++) v = (v << 8) | buf_[pos_++]; return v; }
  uint64_t u64() { uint64_t v = 0; for (int b = 0; b < 8; b++) v = (v << 8) | buf_[pos_++]; return v; }
  bool name(std::string &s) {
    if (!need(4)) return false;
    const uint32_t n = u32();
    if (!need(n)) return false;
    s.assign((const char *)&buf_[pos_], n);
    pos_ += (n + 3) & ~3u;
    return true;
  }
  static size_t tsize(int t) { return t == 1 || t == 2 ? 1 : t == 3 ? 2 : t == 4 || t == 5 ? 4 : 8; }
  bool atts(bool global) {
    if (!need(8)) return false;
    const uint32_t tag = u32(), n = u32();
    if (tag == 0 && n == 0) return true;
    if (tag != 0x0C) return false;
    for (uint32_t k = 0; k < n; k++) {
      std::string nm;
      if (!name(nm) || !need(8)) return false;
      const int type = (int)u32();
      const uint32_t ne = u32();
      const size_t bytes = (size_t)ne * tsize(type);
      if (!need(bytes)) return false;
      if (global && type == 2) gtext_[nm] = std::string((const char *)&buf_[pos_], ne);
      if (global && type == 6 && ne >= 1) {
        uint64_t w = 0;
        for (int b = 0; b < 8; b++) w = (w << 8) | buf_[pos_ + (size_t)b];
        double d;
        memcpy(&d, &w, 8);
        gdouble_[nm] = d;
      }
      pos_ += (bytes + 3) & ~(size_t)3;
    }
    return true;
  }
  bool parse(std::string &err) {
    err = "not a netCDF classic file";
    if (buf_.size() < 8 || buf_[0] != 'C' || buf_[1] != 'D' || buf_[2] != 'F' || (buf_[3] != 1 && buf_[3] != 2)) return false;
    const bool off64 = buf_[3] == 2;
    pos_ = 4;
    numrecs_ = u32();
    if (!need(8)) return false;
    std::vector<size_t> dimsize;
    {
      const uint32_t tag = u32(), n = u32();
      if (!(tag == 0 && n == 0)) {
        if (tag != 0x0A) return false;
        for (uint32_t k = 0; k < n; k++) {
          std::string nm;
          if (!name(nm) || !need(4)) return false;
          dimsize.push_back(u32());
        }
      }
    }
    if (!atts(true)) return false;
    if (!need(8)) return false;
    const uint32_t tag = u32(), nv = u32();
    if (!(tag == 0 && nv == 0) && tag != 0x0B) return false;
    int nrecvars = 0;
    for (uint32_t k = 0; k < nv; k++) {
      std::string nm;
      if (!name(nm) || !need(4)) return false;
      const uint32_t nd = u32();
      VarInfo v{0, 1, 0, 0, false};
      for (uint32_t q = 0; q < nd; q++) {
        if (!need(4)) return false;
        const uint32_t id = u32();
        if (id >= dimsize.size()) return false;
        if (dimsize[id] == 0) v.isrec = true;
        else v.nvals *= dimsize[id];
      }
      if (!atts(false) || !need(8 + (off64 ? 8 : 4))) return false;
      v.type = (int)u32();
      v.vsize = u32();
      v.begin = off64 ? u64() : u32();
      if (v.isrec) { recsize_ += v.vsize; nrecvars++; }
      vars_[nm] = v;
    }
    (void)nrecvars;
    err.clear();
    return true;
  }
};

// restart_read (src/restart_mod.F90:39-56): u, v, gd, ghs and the restart time
inline bool restart_read(const std::string &path, int nlon, int nlat, std::vector<double> &u, std::vector<double> &v,
                         std::vector<double> &gd, std::vector<double> &ghs, std::string &restart_time, std::string &err) {
  NcReader nc;
  if (!nc.open(path, err)) return false;
  if (!nc.get_text_att("restart_time", restart_time)) {
    err = "Failed to get meta \"restart_time\" from file " + path + "!";  // io_mod.F90:679-682
    return false;
  }
  const size_t nf = (size_t)nlon * nlat, nh = (size_t)nlon * (nlat - 1);
  return nc.get_var("u", nf, u, err) && nc.get_var("v", nh, v, err) && nc.get_var("gd", nf, gd, err) &&
         nc.get_var("ghs", nf, ghs, err);
}

// "YYYY-MM-DDTHH:MM:SSZ" or "%Y-%m-%dT%H_%M_%S" -> DateTime (create_datetime of the restart_time meta)
inline bool parse_time_format(const std::string &s, DateTime &t) {
  int y, mo, d, h, mi, se;
  if (sscanf(s.c_str(), "%d-%d-%dT%d:%d:%d", &y, &mo, &d, &h, &mi, &se) != 6 &&
      sscanf(s.c_str(), "%d-%d-%dT%d_%d_%d", &y, &mo, &d, &h, &mi, &se) != 6)
    return false;
  t = DateTime::from_civil(y, mo, d, h, mi);
  t.sec += se;
  return true;
}

}  // namespace host

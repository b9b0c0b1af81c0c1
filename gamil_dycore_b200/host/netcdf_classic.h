// netcdf_classic.h -- a minimal writer for the netCDF classic format (CDF-1), enough for the reference's history
// files: the reference creates one classic-format file per output frame with NF90_CREATE(..., NF90_CLOBBER)
// (src/io_mod.F90:423).  netCDF-C/-Fortran are not available in this image, and the format is simple: a big-endian
// header (dims, global attributes, variables with attributes and data offsets) followed by the fixed-size
// variables and then the records of the variables that use the UNLIMITED dimension.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace host {

class NcFile {
 public:
  enum Type { NC_CHAR = 2, NC_INT = 4, NC_DOUBLE = 6 };
  struct Att {
    std::string name;
    Type type;
    std::vector<unsigned char> bytes;  // big-endian payload
    size_t nelems;
  };
  struct Dim {
    std::string name;
    size_t size;  // 0 = UNLIMITED
  };
  struct Var {
    std::string name;
    std::vector<int> dimids;  // C order (slowest first)
    std::vector<Att> atts;
    Type type;
    std::vector<double> data;  // doubles only (all reference variables are NC_DOUBLE, src/io_mod.F90:352-358)
  };

  int add_dim(const std::string &name, size_t size) { dims_.push_back({name, size}); return (int)dims_.size() - 1; }
  static Att att_text(const std::string &name, const std::string &v) {
    Att a{name, NC_CHAR, std::vector<unsigned char>(v.begin(), v.end()), v.size()};
    return a;
  }
  static Att att_int(const std::string &name, int32_t v) {
    Att a{name, NC_INT, {}, 1};
    put32(a.bytes, (uint32_t)v);
    return a;
  }
  static Att att_double(const std::string &name, double v) {
    Att a{name, NC_DOUBLE, {}, 1};
    put64(a.bytes, v);
    return a;
  }
  void add_global(const Att &a) { gatts_.push_back(a); }
  int add_var(const std::string &name, const std::vector<int> &dimids, const std::vector<Att> &atts) {
    Var v;
    v.name = name; v.dimids = dimids; v.atts = atts; v.type = NC_DOUBLE;
    vars_.push_back(v);
    return (int)vars_.size() - 1;
  }
  void set_data(int var, const std::vector<double> &d) { vars_[(size_t)var].data = d; }
  void set_data(int var, const double *d, size_t n) { vars_[(size_t)var].data.assign(d, d + n); }

  // writes the file with exactly one record; false on I/O error or inconsistent sizes
  bool write(const std::string &path, std::string &err) const {
    std::vector<unsigned char> h;
    h.push_back('C'); h.push_back('D'); h.push_back('F'); h.push_back(1);
    put32(h, 1);  // numrecs
    put32(h, 0x0A); put32(h, (uint32_t)dims_.size());
    for (const Dim &d : dims_) { put_name(h, d.name); put32(h, (uint32_t)d.size); }
    put_atts(h, gatts_);
    // variable sizes
    std::vector<size_t> vsize(vars_.size());
    std::vector<bool> isrec(vars_.size());
    for (size_t k = 0; k < vars_.size(); k++) {
      size_t n = 1;
      isrec[k] = false;
      for (size_t q = 0; q < vars_[k].dimids.size(); q++) {
        const Dim &d = dims_[(size_t)vars_[k].dimids[q]];
        if (d.size == 0) { if (q != 0) { err = "record dimension must be first"; return false; } isrec[k] = true; }
        else n *= d.size;
      }
      if (vars_[k].data.size() != n) { err = "variable " + vars_[k].name + " has the wrong number of values"; return false; }
      vsize[k] = n * 8;
    }
    // header size: build the var list twice (offsets are fixed width)
    auto build_vars = [&](std::vector<unsigned char> &out, const std::vector<uint64_t> &begin) {
      put32(out, 0x0B); put32(out, (uint32_t)vars_.size());
      for (size_t k = 0; k < vars_.size(); k++) {
        put_name(out, vars_[k].name);
        put32(out, (uint32_t)vars_[k].dimids.size());
        for (int id : vars_[k].dimids) put32(out, (uint32_t)id);
        put_atts(out, vars_[k].atts);
        put32(out, (uint32_t)vars_[k].type);
        put32(out, (uint32_t)vsize[k]);
        put32(out, (uint32_t)begin[k]);
      }
    };
    std::vector<uint64_t> begin(vars_.size(), 0);
    std::vector<unsigned char> tmp;
    build_vars(tmp, begin);
    uint64_t off = h.size() + tmp.size();
    for (size_t k = 0; k < vars_.size(); k++)
      if (!isrec[k]) { begin[k] = off; off += vsize[k]; }
    for (size_t k = 0; k < vars_.size(); k++)
      if (isrec[k]) { begin[k] = off; off += vsize[k]; }
    if (off > 0x7fffffffULL) { err = "file too large for the classic format"; return false; }
    build_vars(h, begin);
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { err = "Failed to create NetCDF file to output!"; return false; }  // src/io_mod.F90:424-426
    bool ok = fwrite(h.data(), 1, h.size(), f) == h.size();
    auto dump = [&](size_t k) {
      std::vector<unsigned char> b;
      b.reserve(vars_[k].data.size() * 8);
      for (double v : vars_[k].data) put64(b, v);
      ok = ok && fwrite(b.data(), 1, b.size(), f) == b.size();
    };
    for (size_t k = 0; k < vars_.size(); k++)
      if (!isrec[k]) dump(k);
    for (size_t k = 0; k < vars_.size(); k++)
      if (isrec[k]) dump(k);
    ok = (fclose(f) == 0) && ok;
    if (!ok) err = "write error on " + path;
    return ok;
  }

 private:
  static void put32(std::vector<unsigned char> &b, uint32_t v) {
    b.push_back((unsigned char)(v >> 24)); b.push_back((unsigned char)(v >> 16));
    b.push_back((unsigned char)(v >> 8)); b.push_back((unsigned char)v);
  }
  static void put64(std::vector<unsigned char> &b, double d) {
    uint64_t v;
    memcpy(&v, &d, 8);
    for (int s = 56; s >= 0; s -= 8) b.push_back((unsigned char)(v >> s));
  }
  static void pad4(std::vector<unsigned char> &b) { while (b.size() % 4) b.push_back(0); }
  static void put_name(std::vector<unsigned char> &b, const std::string &s) {
    put32(b, (uint32_t)s.size());
    b.insert(b.end(), s.begin(), s.end());
    pad4(b);
  }
  static void put_atts(std::vector<unsigned char> &b, const std::vector<Att> &atts) {
    if (atts.empty()) { put32(b, 0); put32(b, 0); return; }
    put32(b, 0x0C); put32(b, (uint32_t)atts.size());
    for (const Att &a : atts) {
      put_name(b, a.name);
      put32(b, (uint32_t)a.type);
      put32(b, (uint32_t)a.nelems);
      b.insert(b.end(), a.bytes.begin(), a.bytes.end());
      pad4(b);
    }
  }
  std::vector<Dim> dims_;
  std::vector<Att> gatts_;
  std::vector<Var> vars_;
};

}  // namespace host

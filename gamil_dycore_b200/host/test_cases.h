// test_cases.h -- the IC plugins of src/test_cases/barotropic/*_test_mod.F90 (host code, run once).
// Each fills u [nlat][nlon], v [nlat-1][nlon], gd, ghs [nlat][nlon] (compact layout) on the mesh of
// src/mesh_mod.F90:44-114 with cos(pole) = 0, i.e. BEFORE reset_cos_lat_at_poles.
#pragma once
#include <string>
#include <vector>

#include "params.h"

namespace host {

struct Fields {
  int nlon = 0, nlat = 0;
  std::vector<double> u, v, gd, ghs;
};

// returns false for an unknown test_case (src/dycore_test.F90:29-42)
bool set_initial_condition(const Params &p, Fields &f, std::string &notice, std::string &err);

// adaptive Gauss-Kronrod (21 point) quadrature with QAGS-style global error control
// (lib/quadpack.f90 `qags`, called with epsabs 1e-10, epsrel 1e-3 at jet_zonal_flow_test_mod.F90:63)
double integrate_gk21(double (*f)(double), double a, double b, double epsabs, double epsrel, int limit, int *ier);
double jet_gh_profile(double lat);

}  // namespace host

// host_capi.cpp -- C ABI of the host-side IC plugins (include/gmd_host.h); no CUDA in here.
#include "../../include/gmd_host.h"

#include <cstring>
#include <string>

#include "params.h"
#include "test_cases.h"

static thread_local std::string g_host_err;

extern "C" {

const char *gmd_host_last_error(void) { return g_host_err.c_str(); }

int gmd_host_initial_condition(const char *test_case, int num_lon, int num_lat, double *u, double *v, double *gd,
                               double *ghs) {
  if (!test_case || !u || !v || !gd || !ghs) { g_host_err = "null argument"; return 2; }
  if (num_lon < 4 || num_lat < 5) { g_host_err = "grid too small"; return 2; }
  host::Params p;
  p.num_lon = num_lon;
  p.num_lat = num_lat;
  p.test_case = test_case;
  host::Fields f;
  std::string notice, err;
  if (!host::set_initial_condition(p, f, notice, err)) { g_host_err = err; return 2; }
  std::memcpy(u, f.u.data(), f.u.size() * sizeof(double));
  std::memcpy(v, f.v.data(), f.v.size() * sizeof(double));
  std::memcpy(gd, f.gd.data(), f.gd.size() * sizeof(double));
  std::memcpy(ghs, f.ghs.data(), f.ghs.size() * sizeof(double));
  return 0;
}

}  // extern "C"

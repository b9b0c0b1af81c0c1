// params.h -- namelist /dycore_params/ and the test-case namelists (src/params_mod.F90:13-115).
// Host-side mirror of params_mod: same key names, defaults and the `restart_period` fallback; unknown keys are an
// error, as they are for a gfortran namelist read (which is why run/namelist.rh_test and namelist.jz_test of the
// reference commit cannot be read as shipped).
#pragma once
#include <map>
#include <string>
#include <vector>

namespace host {

struct Params {
  int num_lon = 0, num_lat = 0;
  int subcycles = 4;
  double time_step_size = 0.0;
  bool use_diffusion = false;
  double diffusion_coef = 0.0;
  int diffusion_order = 2;
  int run_days = 0, run_hours = 0, run_minutes = 0, run_seconds = 0;
  int start_time[5] = {0, 0, 0, 0, 0}, end_time[5] = {0, 0, 0, 0, 0};
  std::string time_units = "days";
  std::string test_case, case_name, case_desc, author;
  std::string history_periods = "6 hours";
  std::string restart_period, restart_file;
  std::string time_scheme;
  int time_order = 0;
  bool qcon_modified = false;
  std::string split_scheme;
  std::string uv_adv_scheme = "center_diff";
  double uv_adv_upwind_lon_beta = 0.0, uv_adv_upwind_lat_beta = 0.5;
  bool use_zonal_tend_filter = true;
  int zonal_tend_filter_cutoff_wavenumber[20] = {0};
  // moving reduced tendency (run/namelist.jz_test:16-19 of the reference; specified in DESIGN.md section 8)
  bool use_zonal_reduce = false, reduce_adv_lon = false, use_reduce_tend_smooth = false;
  int zonal_reduce_factors[20] = {0};
  bool is_restart_run = false;
  // test-case groups
  double rh_R = 4.0, rh_omg = 7.848e-6, rh_gd0 = 8.0e3 * 9.80616;  // rossby_haurwitz_wave_test_mod.F90:14-18
  bool smooth_mountain = false;                                    // mountain_zonal_flow_test_mod.F90:24
  std::string namelist_file;
};

// generic Fortran-namelist reader: group name -> key -> list of value tokens
typedef std::map<std::string, std::map<std::string, std::vector<std::string>>> NamelistGroups;
bool namelist_parse(const std::string &text, NamelistGroups &out, std::string &err);

// params_read (src/params_mod.F90:102-115); returns false and sets err on failure
bool params_read(const std::string &path, Params &p, std::string &err);
bool params_from_text(const std::string &text, Params &p, std::string &err);

}  // namespace host

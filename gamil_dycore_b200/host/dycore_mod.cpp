#include "dycore_mod.h"

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gmd.h"
#include "history.h"
#include "log.h"
#include "restart.h"

namespace host {

Params params;
Fields state_ic;
TimeManager timer;

static gmd_model *model = nullptr;
static double last_beta = 1.0;

static void check(int ier) {
  if (ier) log_error(gmd_last_error());
}

void dycore_init() {
  if (params.case_name.empty()) log_error("case_name is not set!");  // src/dycore_mod.F90:62-64
  // same order of module initialisation (and of their notices) as src/dycore_mod.F90:66-76
  log_notice("Log module is initialized.");
  log_notice("Mesh module is initialized.");
  timer.init(params);
  log_notice("Parallel module is initialized.");
  if (params.time_units != "days" && params.time_units != "hours" && params.time_units != "seconds")
    log_error("Invalid time_units " + params.time_units + "!");  // src/io_mod.F90:95-97
  log_notice("IO module is initialized.");
  log_notice("Diag module is initialized.");
  {
    double period = 0.0;
    if (!TimeManager::parse_period(params.history_periods, period, params.time_step_size))
      log_error("Invalid IO period " + params.history_periods + "!");
    timer.add_alert("hist0.output", period);
    log_notice("Create output dataset " + params.case_name + ".h0.");
    log_notice("Create output dataset " + params.case_name + ".debug.");
    log_notice("History module is initialized.");
    // restart_init, src/restart_mod.F90:22-35
    if (!TimeManager::parse_period(params.restart_period, period, params.time_step_size))
      log_error("Invalid IO period " + params.restart_period + "!");
    timer.add_alert("restart.output", period);
    log_notice("Create output dataset " + params.case_name + ".r.");
  }
  log_notice("Data module is initialized.");
  log_notice("Filter module is initialized.");

  gmd_config c;
  gmd_config_defaults(&c);
  c.num_lon = params.num_lon;
  c.num_lat = params.num_lat;
  c.subcycles = params.subcycles;
  c.time_step_size = params.time_step_size;
  c.qcon_modified = params.qcon_modified ? 1 : 0;
  if (params.time_scheme == "predict_correct") c.time_scheme = GMD_TIME_PREDICT_CORRECT;
  else if (params.time_scheme == "runge_kutta") c.time_scheme = GMD_TIME_RUNGE_KUTTA;   // params_mod.F90:40-44 (specified)
  else log_error("Unknown time_scheme " + params.time_scheme + "!");  // :78-83
  if (params.time_order) c.time_order = params.time_order;
  if (params.split_scheme == "csp1") c.split_scheme = GMD_SPLIT_CSP1;
  else if (params.split_scheme == "csp2") c.split_scheme = GMD_SPLIT_CSP2;
  else if (params.split_scheme == "isp") c.split_scheme = GMD_SPLIT_ISP;
  else {
    c.split_scheme = GMD_SPLIT_NONE;
    log_notice("No fast-slow split.");  // :92-95
  }
  if (params.uv_adv_scheme == "center_diff") c.uv_adv_scheme = GMD_ADV_CENTER_DIFF;
  else if (params.uv_adv_scheme == "upwind") c.uv_adv_scheme = GMD_ADV_UPWIND;
  else if (params.uv_adv_scheme == "weno") c.uv_adv_scheme = GMD_ADV_WENO;
  else log_error("Unknown uv_adv_scheme " + params.uv_adv_scheme + "!");  // :104-106
  c.uv_adv_upwind_lon_beta = params.uv_adv_upwind_lon_beta;
  c.uv_adv_upwind_lat_beta = params.uv_adv_upwind_lat_beta;
  c.use_zonal_tend_filter = params.use_zonal_tend_filter ? 1 : 0;
  memcpy(c.zonal_tend_filter_cutoff_wavenumber, params.zonal_tend_filter_cutoff_wavenumber, sizeof(int) * 20);
  c.use_zonal_reduce = params.use_zonal_reduce ? 1 : 0;
  c.reduce_adv_lon = params.reduce_adv_lon ? 1 : 0;
  c.use_reduce_tend_smooth = params.use_reduce_tend_smooth ? 1 : 0;
  memcpy(c.zonal_reduce_factors, params.zonal_reduce_factors, sizeof(int) * 20);
  c.use_diffusion = params.use_diffusion ? 1 : 0;
  c.diffusion_order = params.diffusion_order;
  c.diffusion_coef = params.diffusion_coef;
  check(gmd_create(&c, &model));
  log_notice("Dycore module is initialized.");
}

void dycore_restart() {  // restart_read, src/restart_mod.F90:39-56
  log_notice("Create input dataset " + params.restart_file + ".");
  std::string when, err;
  if (!restart_read(params.restart_file, params.num_lon, params.num_lat, state_ic.u, state_ic.v, state_ic.gd,
                    state_ic.ghs, when, err))
    log_error(err);
  DateTime t;
  if (!parse_time_format(when, t)) log_error("Invalid restart_time " + when + "!");
  timer.reset_start_time(t);   // time_reset_start_time, src/time_mod.F90:81-91
  log_notice("Reset time to " + timer.curr_time_format + ".");
}

static void output() {  // src/dycore_mod.F90:175-182
  const bool hist = timer.is_alerted("hist0.output"), rst = timer.is_alerted("restart.output");
  if (!hist && !rst) return;
  const int nlon = params.num_lon, nlat = params.num_lat;
  const size_t nf = (size_t)nlon * nlat, nh = (size_t)nlon * (nlat - 1);
  std::vector<double> u(nf), v(nh), gd(nf), vor(nh), div(nf);
  check(gmd_get_state(model, u.data(), v.data(), gd.data(), GMD_LAYOUT_COMPACT));
  std::string path, err;
  if (hist) {
    check(gmd_get_vor_div(model, vor.data(), div.data(), GMD_LAYOUT_COMPACT));
    double m = 0, e = 0, b = 0;
    check(gmd_get_diag(model, &m, &e, &b));
    if (!history_write(params, timer, u, v, gd, state_ic.ghs, vor, div, e, m, path, err)) log_error(err);
  }
  if (rst && !restart_write(params, timer, u, v, gd, state_ic.ghs, path, err)) log_error(err);
}

void dycore_run() {
  // reset_cos_lat_at_poles + iap_transform + diag_run (src/dycore_mod.F90:121-125) live behind gmd_run_init
  check(gmd_set_state(model, state_ic.u.data(), state_ic.v.data(), state_ic.gd.data(), state_ic.ghs.data(),
                      GMD_LAYOUT_COMPACT));
  check(gmd_run_init(model));
  double m = 0, e = 0, b = 0;
  check(gmd_get_diag(model, &m, &e, &b));
  output();
  log_step(timer.curr_time.format(true), {{"total_mass", m}, {"total_energy", e}});
  std::vector<double> ms, es, bs;
  while (!timer.is_finished()) {
    // steps are batched up to the next output alert: one launch sequence, one sync, one read of the diag series
    long n = std::min(std::min(timer.steps_until_alert("hist0.output"), timer.steps_until_alert("restart.output")),
                      timer.steps_until_end());
    n = std::max(1L, std::min(n, 2048L));
    check(gmd_step(model, (int)n));
    ms.assign((size_t)n, 0.0); es.assign((size_t)n, 0.0); bs.assign((size_t)n, 0.0);
    check(gmd_get_diag_series(model, (int)n, ms.data(), es.data(), bs.data()));
    for (long k = 0; k < n; k++) {
      timer.advance();
      if (k == n - 1) output();
      log_step(timer.curr_time.format(true), {{"total_mass", ms[(size_t)k]}, {"total_energy", es[(size_t)k]}, {"beta", bs[(size_t)k]}});
    }
    last_beta = bs.back();
  }
}

void dycore_final() {
  gmd_destroy(model);
  model = nullptr;
  log_notice("Mesh module is finalized.");
  log_notice("Parallel module is finalized.");
  log_notice("History module is finalized.");
  log_notice("Data module is finalized.");
  log_notice("Filter module is finalized.");
  log_notice("Dycore module is finalized.");
}

}  // namespace host

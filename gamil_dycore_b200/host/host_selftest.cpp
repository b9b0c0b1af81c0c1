// host_selftest.cpp -- command-line hooks into the host-side code (no GPU needed) for the CPU test-suite:
//   parse <namelist>                 print the parsed /dycore_params/ as "key=value" lines, or "ERROR: ..." (exit 2)
//   fmt <x>                          to_string(real8, 20) of src/string_mod.F90:64-81
//   ic <namelist> <out.bin>          run the IC plugin, dump u, v, gd, ghs as raw doubles
//   history <namelist> <nsteps>      write one h0 frame of the IC fields as if after nsteps steps; prints the path
//   clock <namelist> <nsteps>        print the log-line time stamps and alert decisions of the first nsteps steps
//   tables <namelist> <out.bin>      the product's per-latitude coefficient tables and filter / reduced row maps
//                                    (csrc/gmd_mesh.h, what gmd_create builds): 10 tables of raw doubles in the order of
//                                    gmd_get_table, then flag_full, cut_full, flag_half, cut_half, red_full, red_half as int32
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../csrc/gmd_mesh.h"
#include "history.h"
#include "restart.h"
#include "log.h"
#include "params.h"
#include "test_cases.h"
#include "time_manager.h"

using namespace host;

static int load(const char *path, Params &p) {
  std::string err;
  if (!params_read(path, p, err)) {
    printf("ERROR: %s\n", err.c_str());
    return 2;
  }
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 3) return 1;
  const std::string cmd = argv[1];
  if (cmd == "fmt") {
    printf("%s\n", to_string_r8(atof(argv[2]), 20).c_str());
    return 0;
  }
  Params p;
  if (int r = load(argv[2], p)) return r;
  if (cmd == "parse") {
    printf("test_case=%s\ncase_name=%s\ncase_desc=%s\nnum_lon=%d\nnum_lat=%d\nsubcycles=%d\ntime_step_size=%.17g\n", p.test_case.c_str(),
           p.case_name.c_str(), p.case_desc.c_str(), p.num_lon, p.num_lat, p.subcycles, p.time_step_size);
    printf("run_days=%d\nrun_hours=%d\nhistory_periods=%s\nrestart_period=%s\ntime_scheme=%s\nsplit_scheme=%s\n", p.run_days, p.run_hours,
           p.history_periods.c_str(), p.restart_period.c_str(), p.time_scheme.c_str(), p.split_scheme.c_str());
    printf("uv_adv_scheme=%s\nuv_adv_upwind_lon_beta=%.17g\nuv_adv_upwind_lat_beta=%.17g\nqcon_modified=%d\n", p.uv_adv_scheme.c_str(),
           p.uv_adv_upwind_lon_beta, p.uv_adv_upwind_lat_beta, (int)p.qcon_modified);
    printf("use_zonal_tend_filter=%d\nuse_diffusion=%d\ndiffusion_order=%d\ndiffusion_coef=%.17g\nsmooth_mountain=%d\ncutoff=", (int)p.use_zonal_tend_filter,
           (int)p.use_diffusion, p.diffusion_order, p.diffusion_coef, (int)p.smooth_mountain);
    for (int k = 0; k < 20; k++) printf("%d%s", p.zonal_tend_filter_cutoff_wavenumber[k], k == 19 ? "\n" : ",");
    printf("time_order=%d\nuse_zonal_reduce=%d\nreduce_adv_lon=%d\nuse_reduce_tend_smooth=%d\nreduce_factors=", p.time_order,
           (int)p.use_zonal_reduce, (int)p.reduce_adv_lon, (int)p.use_reduce_tend_smooth);
    for (int k = 0; k < 20; k++) printf("%d%s", p.zonal_reduce_factors[k], k == 19 ? "\n" : ",");
    return 0;
  }
  if (cmd == "ic" && argc >= 4) {
    Fields f;
    std::string notice, err;
    if (!set_initial_condition(p, f, notice, err)) { printf("ERROR: %s\n", err.c_str()); return 2; }
    FILE *o = fopen(argv[3], "wb");
    if (!o) return 3;
    fwrite(f.u.data(), 8, f.u.size(), o);
    fwrite(f.v.data(), 8, f.v.size(), o);
    fwrite(f.gd.data(), 8, f.gd.size(), o);
    fwrite(f.ghs.data(), 8, f.ghs.size(), o);
    fclose(o);
    printf("%s\n", notice.c_str());
    return 0;
  }
  if (cmd == "history" && argc >= 4) {
    Fields f;
    std::string notice, err, path;
    if (!set_initial_condition(p, f, notice, err)) { printf("ERROR: %s\n", err.c_str()); return 2; }
    TimeManager tm;
    tm.init(p);
    for (int k = 0; k < atoi(argv[3]); k++) tm.advance();
    std::vector<double> vor(f.v.size()), div(f.u.size());
    for (size_t k = 0; k < vor.size(); k++) vor[k] = 1.0e-5 * (double)(k % 7);
    for (size_t k = 0; k < div.size(); k++) div[k] = 1.0e-6 * (double)(k % 5);
    if (!history_write(p, tm, f.u, f.v, f.gd, f.ghs, vor, div, 2.5, 1.5, path, err)) { printf("ERROR: %s\n", err.c_str()); return 2; }
    printf("%s\n", path.c_str());
    return 0;
  }
  if (cmd == "restart" && argc >= 4) {
    // write a restart file from the analytic initial condition after argv[3] clock steps, read it back, compare bits
    Fields f, g;
    std::string notice, err, path, when;
    if (!set_initial_condition(p, f, notice, err)) { printf("ERROR: %s\n", err.c_str()); return 2; }
    TimeManager tm;
    tm.init(p);
    for (int k = 0; k < atoi(argv[3]); k++) tm.advance();
    if (!restart_write(p, tm, f.u, f.v, f.gd, f.ghs, path, err)) { printf("ERROR: %s\n", err.c_str()); return 2; }
    if (!restart_read(path, p.num_lon, p.num_lat, g.u, g.v, g.gd, g.ghs, when, err)) { printf("ERROR: %s\n", err.c_str()); return 2; }
    DateTime t;
    if (!parse_time_format(when, t)) { printf("ERROR: bad restart_time %s\n", when.c_str()); return 2; }
    const bool same = f.u == g.u && f.v == g.v && f.gd == g.gd && f.ghs == g.ghs && t.sec == tm.curr_time.sec;
    printf("%s\n%s\n%s\n", path.c_str(), when.c_str(), same ? "identical" : "DIFFERENT");
    return same ? 0 : 4;
  }
  if (cmd == "tables" && argc >= 4) {
    gmd::HostMesh M;
    M.init(p.num_lon, p.num_lat, /*reset_poles=*/true);
    M.filter_init(p.use_zonal_tend_filter, p.zonal_tend_filter_cutoff_wavenumber);
    if (int bad = M.reduce_init(p.use_zonal_reduce, p.zonal_reduce_factors)) { printf("ERROR: zonal_reduce_factors(%d)\n", bad); return 2; }
    FILE *o = fopen(argv[3], "wb");
    if (!o) return 3;
    const std::vector<double> *tab[10] = {&M.full_cos, &M.half_cos, &M.full_f, &M.full_c, &M.full_dlon,
                                          &M.half_dlon, &M.full_dlat, &M.half_dlat, &M.full_lat, &M.half_lat};
    for (int w = 0; w < 10; w++) {
      const bool half = (w == 1 || w == 5 || w == 7 || w == 9);
      fwrite(tab[w]->data() + gmd::TPAD, 8, (size_t)(p.num_lat - (half ? 1 : 0)), o);
    }
    for (const std::vector<int> *v : {&M.flag_full, &M.cut_full, &M.flag_half, &M.cut_half, &M.red_full, &M.red_half})
      fwrite(v->data(), 4, (size_t)p.num_lat, o);
    fclose(o);
    printf("%d %d %d\n", p.num_lon, p.num_lat, M.cutoff_max);
    return 0;
  }
  if (cmd == "clock" && argc >= 4) {
    TimeManager tm;
    tm.init(p);
    double period = 0;
    if (!TimeManager::parse_period(p.history_periods, period, p.time_step_size)) { printf("ERROR: bad period\n"); return 2; }
    tm.add_alert("hist0.output", period);
    printf("%s %d %ld %ld\n", tm.curr_time.format(true).c_str(), (int)tm.is_alerted("hist0.output"), tm.steps_until_alert("hist0.output"),
           tm.steps_until_end());
    for (int k = 0; k < atoi(argv[3]) && !tm.is_finished(); k++) {
      tm.advance();
      const bool a = tm.is_alerted("hist0.output");
      printf("%s %d %s\n", tm.curr_time.format(true).c_str(), (int)a, tm.curr_time_format.c_str());
    }
    return 0;
  }
  return 1;
}

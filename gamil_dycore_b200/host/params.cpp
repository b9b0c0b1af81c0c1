#include "params.h"

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace host {

static std::string lower(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
  return s;
}

bool namelist_parse(const std::string &text, NamelistGroups &out, std::string &err) {
  size_t i = 0;
  const size_t n = text.size();
  std::string group;
  auto skip_ws = [&]() {
    while (i < n) {
      if (text[i] == '!') {
        while (i < n && text[i] != '\n') i++;
      } else if (std::isspace((unsigned char)text[i]) || text[i] == ',') {
        i++;
      } else {
        break;
      }
    }
  };
  while (true) {
    skip_ws();
    if (i >= n) break;
    if (group.empty()) {
      if (text[i] != '&') { i++; continue; }  // text outside groups is ignored
      i++;
      size_t b = i;
      while (i < n && (std::isalnum((unsigned char)text[i]) || text[i] == '_')) i++;
      group = lower(text.substr(b, i - b));
      out[group];
      continue;
    }
    if (text[i] == '/') { group.clear(); i++; continue; }
    if (text[i] == '&') {  // &end
      size_t b = ++i;
      while (i < n && std::isalpha((unsigned char)text[i])) i++;
      if (lower(text.substr(b, i - b)) == "end") { group.clear(); continue; }
      err = "unexpected '&' inside namelist group " + group;
      return false;
    }
    size_t b = i;
    while (i < n && (std::isalnum((unsigned char)text[i]) || text[i] == '_')) i++;
    if (i == b) { err = std::string("unexpected character '") + text[i] + "' in namelist group " + group; return false; }
    std::string key = lower(text.substr(b, i - b));
    int sub = 0;   // name(k) = v1, v2, ...: the values start at element k (a single subscript; ranges are refused)
    if (i < n && text[i] == '(') {
      const size_t sb = ++i;
      while (i < n && text[i] != ')') i++;
      const std::string inside = text.substr(sb, i - sb);
      if (i < n) i++;
      char *endp = nullptr;
      const long k = std::strtol(inside.c_str(), &endp, 10);
      while (endp && *endp == ' ') endp++;
      if (inside.empty() || !endp || *endp != '\0' || k < 1) { err = "unsupported subscript (" + inside + ") on " + key; return false; }
      sub = (int)k;
    }
    skip_ws();
    if (i >= n || text[i] != '=') { err = "expected '=' after " + key; return false; }
    i++;
    std::vector<std::string> &dst = out[group][key];
    std::vector<std::string> vals;
    while (true) {
      skip_ws();
      if (i >= n) break;
      if (text[i] == '/' || text[i] == '&') break;
      if (text[i] == '\'' || text[i] == '"') {
        const char q = text[i++];
        std::string s;
        while (i < n) {
          if (text[i] == q) {
            if (i + 1 < n && text[i + 1] == q) { s += q; i += 2; continue; }
            break;
          }
          s += text[i++];
        }
        i++;
        vals.push_back(s);
        continue;
      }
      // a bare token; stop if it is followed by '=' (then it is the next key)
      size_t tb = i;
      while (i < n && !std::isspace((unsigned char)text[i]) && text[i] != ',' && text[i] != '/' && text[i] != '!' &&
             text[i] != '=')
        i++;
      size_t te = i;
      size_t k = i;
      while (k < n && (text[k] == ' ' || text[k] == '\t')) k++;
      if ((k < n && text[k] == '=') || (i < n && text[i] == '(')) { i = tb; break; }
      vals.push_back(text.substr(tb, te - tb));
    }
    if (sub <= 1) {
      dst = vals;
    } else {   // elements sub, sub+1, ... are assigned; a gap is padded with empty tokens (= "leave the element as it is")
      if (dst.size() < (size_t)(sub - 1) + vals.size()) dst.resize((size_t)(sub - 1) + vals.size());
      for (size_t q = 0; q < vals.size(); q++) dst[(size_t)(sub - 1) + q] = vals[q];
    }
  }
  return true;
}

static bool to_bool(const std::string &s, bool &v) {
  std::string t = lower(s);
  if (t == ".true." || t == "t" || t == "true" || t == ".t.") { v = true; return true; }
  if (t == ".false." || t == "f" || t == "false" || t == ".f.") { v = false; return true; }
  return false;
}
static bool to_num(const std::string &s, double &v) {
  std::string t = s;
  for (char &c : t)
    if (c == 'd' || c == 'D') c = 'e';
  char *end = nullptr;
  v = std::strtod(t.c_str(), &end);
  return end && *end == 0 && !t.empty();
}

bool params_from_text(const std::string &text, Params &p, std::string &err) {
  NamelistGroups g;
  if (!namelist_parse(text, g, err)) return false;
  if (!g.count("dycore_params")) { err = "namelist group &dycore_params not found"; return false; }
  for (auto &kv : g["dycore_params"]) {
    const std::string &k = kv.first;
    const std::vector<std::string> &v = kv.second;
    auto need1 = [&]() { return !v.empty(); };
    double x = 0;
    bool b = false;
#define NUM(field) { if (!need1() || !to_num(v[0], x)) { err = "bad value for " + k; return false; } p.field = (decltype(p.field))x; continue; }
#define STR(field) { if (!need1()) { err = "bad value for " + k; return false; } p.field = v[0]; continue; }
#define LOG(field) { if (!need1() || !to_bool(v[0], b)) { err = "bad value for " + k; return false; } p.field = b; continue; }
    if (k == "num_lon") NUM(num_lon)
    if (k == "num_lat") NUM(num_lat)
    if (k == "subcycles") NUM(subcycles)
    if (k == "run_days") NUM(run_days)
    if (k == "run_hours") NUM(run_hours)
    if (k == "run_minutes") NUM(run_minutes)
    if (k == "run_seconds") NUM(run_seconds)
    if (k == "time_units") STR(time_units)
    if (k == "time_step_size") NUM(time_step_size)
    if (k == "test_case") STR(test_case)
    if (k == "case_name") STR(case_name)
    if (k == "case_desc") STR(case_desc)
    if (k == "author") STR(author)
    if (k == "history_periods") STR(history_periods)
    if (k == "restart_period") STR(restart_period)
    if (k == "restart_file") STR(restart_file)
    if (k == "time_scheme") STR(time_scheme)
    if (k == "time_order") NUM(time_order)
    if (k == "qcon_modified") LOG(qcon_modified)
    if (k == "split_scheme") STR(split_scheme)
    if (k == "uv_adv_scheme") STR(uv_adv_scheme)
    if (k == "uv_adv_upwind_lon_beta") NUM(uv_adv_upwind_lon_beta)
    if (k == "uv_adv_upwind_lat_beta") NUM(uv_adv_upwind_lat_beta)
    if (k == "use_zonal_tend_filter") LOG(use_zonal_tend_filter)
    if (k == "use_diffusion") LOG(use_diffusion)
    if (k == "diffusion_order") NUM(diffusion_order)
    if (k == "diffusion_coef") NUM(diffusion_coef)
    if (k == "start_time" || k == "end_time") {
      int *dst = (k == "start_time") ? p.start_time : p.end_time;
      if (v.size() > 5) { err = "too many values for " + k; return false; }
      for (size_t q = 0; q < v.size(); q++) {
        if (v[q].empty()) continue;
        if (!to_num(v[q], x)) { err = "bad value for " + k; return false; }
        dst[q] = (int)x;
      }
      continue;
    }
    if (k == "use_zonal_reduce") LOG(use_zonal_reduce)
    if (k == "reduce_adv_lon") LOG(reduce_adv_lon)
    if (k == "use_reduce_tend_smooth") LOG(use_reduce_tend_smooth)
    if (k == "zonal_tend_filter_cutoff_wavenumber" || k == "zonal_reduce_factors") {
      int *dst = (k == "zonal_reduce_factors") ? p.zonal_reduce_factors : p.zonal_tend_filter_cutoff_wavenumber;
      if (v.size() > 20) { err = "too many values for " + k; return false; }
      size_t q = 0;
      for (const std::string &tok : v) {  // r*c repeat form allowed
        if (tok.empty()) { q++; continue; }
        size_t star = tok.find('*');
        int rep = 1;
        std::string val = tok;
        if (star != std::string::npos) { rep = std::atoi(tok.substr(0, star).c_str()); val = tok.substr(star + 1); }
        if (!to_num(val, x)) { err = "bad value for " + k; return false; }
        for (int r = 0; r < rep && q < 20; r++) dst[q++] = (int)x;
      }
      continue;
    }
    // gfortran: "Cannot match namelist object name"
    err = "Cannot match namelist object name " + k + " in &dycore_params";
    return false;
#undef NUM
#undef STR
#undef LOG
  }
  if (g.count("rossby_haurwitz_wave_test_params")) {
    for (auto &kv : g["rossby_haurwitz_wave_test_params"]) {
      double x;
      if (kv.second.empty() || !to_num(kv.second[0], x)) { err = "bad value for " + kv.first; return false; }
      if (kv.first == "r") p.rh_R = x;
      else if (kv.first == "omg") p.rh_omg = x;
      else if (kv.first == "gd0") p.rh_gd0 = x;
      else { err = "Cannot match namelist object name " + kv.first; return false; }
    }
  }
  if (g.count("mountain_zonal_flow_test_params")) {
    for (auto &kv : g["mountain_zonal_flow_test_params"]) {
      bool b;
      if (kv.first != "smooth_mountain") { err = "Cannot match namelist object name " + kv.first; return false; }
      if (kv.second.empty() || !to_bool(kv.second[0], b)) { err = "bad value for smooth_mountain"; return false; }
      p.smooth_mountain = b;
    }
  }
  p.is_restart_run = !p.restart_file.empty();                        // params_mod.F90:112
  if (p.restart_period.empty()) p.restart_period = p.history_periods;  // :113
  return true;
}

bool params_read(const std::string &path, Params &p, std::string &err) {
  std::ifstream f(path);
  if (!f) { err = "cannot open namelist file " + path; return false; }
  std::stringstream ss;
  ss << f.rdbuf();
  p.namelist_file = path;
  return params_from_text(ss.str(), p, err);
}

}  // namespace host

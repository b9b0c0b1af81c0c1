// log.h -- stdout formats of src/log_mod.F90:38-103 and src/string_mod.F90:64-81.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

namespace host {

// list-directed write: one leading blank (gfortran), e.g. " [Notice]: Log module is initialized."
inline void log_notice(const std::string &msg) { printf(" [Notice]: %s\n", msg.c_str()); }
inline void log_warning(const std::string &msg) { printf(" [Warning]: %s\n", msg.c_str()); }
[[noreturn]] inline void log_error(const std::string &msg) {  // log_error: message, then `stop 1`
  printf(" [Error]: %s\n", msg.c_str());
  fflush(stdout);
  exit(1);
}

// Fortran Ew.d with a 2-digit exponent: value = 0.d1d2...dd x 10^e   (to_string(real8, 20): E20.14, E20.13 if < 0)
inline std::string fortran_e(double x, int w, int d) {
  char buf[64];
  if (std::isnan(x)) { snprintf(buf, sizeof buf, "%*s", w, "NaN"); return buf; }
  if (std::isinf(x)) { snprintf(buf, sizeof buf, "%*s", w, x > 0 ? "Infinity" : "-Infinity"); return buf; }
  int e = 0;
  double m = std::fabs(x);
  char digits[64];
  if (m == 0.0) {
    snprintf(digits, sizeof digits, "%0*d", d, 0);
  } else {
    // d significant digits via %.{d-1}e, then shift the decimal point one place
    char t[64];
    snprintf(t, sizeof t, "%.*e", d - 1, m);
    int k = 0;
    for (const char *p = t; *p && *p != 'e'; p++)
      if (*p != '.') digits[k++] = *p;
    digits[k] = 0;
    e = atoi(strchr(t, 'e') + 1) + 1;
  }
  char body[96];
  if (std::abs(e) < 100) snprintf(body, sizeof body, "%s0.%sE%c%02d", x < 0 ? "-" : "", digits, e < 0 ? '-' : '+', std::abs(e));
  else snprintf(body, sizeof body, "%s0.%s%c%03d", x < 0 ? "-" : "", digits, e < 0 ? '-' : '+', std::abs(e));
  snprintf(buf, sizeof buf, "%*s", w, body);
  return buf;
}
inline std::string to_string_r8(double x, int w = 20) { return x >= 0 ? fortran_e(x, w, w - 6) : fortran_e(x, w, w - 7); }

// log_step (src/log_mod.F90:81-103): " => <iso time>" then " <value>" per diagnostic in insertion order
inline void log_step(const std::string &iso_time, const std::vector<std::pair<std::string, double>> &diags) {
  printf(" => %s", iso_time.c_str());
  for (auto &kv : diags) {
    std::string s = to_string_r8(kv.second, 20);
    size_t b = s.find_last_not_of(' ');
    printf(" %s", s.substr(0, b + 1).c_str());
  }
  printf("\n");
}

}  // namespace host

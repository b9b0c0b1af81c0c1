// time_manager.h -- step counter, model clock and output alerts (src/time_mod.F90:51-213).
// The reference's clock is lib/datetime's proleptic calendar starting at 0001-01-01T00:00:00 when
// run_days/hours/minutes are used (src/time_mod.F90:53-54), else start_time(5)/end_time(5) = y,m,d,h,min.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <map>
#include <string>

#include "params.h"

namespace host {

struct DateTime {  // seconds since 0001-01-01T00:00:00 in the proleptic Gregorian calendar
  double sec = 0.0;
  static long long days_from_civil(long long y, unsigned m, unsigned d) {
    y -= m <= 2;
    const long long era = (y >= 0 ? y : y - 399) / 400;
    const unsigned yoe = (unsigned)(y - era * 400);
    const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (long long)doe - 719468;
  }
  static void civil_from_days(long long z, long long &y, unsigned &m, unsigned &d) {
    z += 719468;
    const long long era = (z >= 0 ? z : z - 146096) / 146097;
    const unsigned doe = (unsigned)(z - era * 146097);
    const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    y = (long long)yoe + era * 400;
    const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const unsigned mp = (5 * doy + 2) / 153;
    d = doy - (153 * mp + 2) / 5 + 1;
    m = mp + (mp < 10 ? 3 : -9);
    y += (m <= 2);
  }
  static DateTime from_civil(int y, int mo, int d, int h, int mi) {
    DateTime t;
    const long long base = days_from_civil(1, 1, 1);
    t.sec = (double)(days_from_civil(y, (unsigned)mo, (unsigned)d) - base) * 86400.0 + h * 3600.0 + mi * 60.0;
    return t;
  }
  // iso: "YYYY-MM-DDTHH:MM:SSZ" (isoformat, lib/datetime/src/datetime_mod.F90:235-243); else "%Y-%m-%dT%H_%M_%S"
  std::string format(bool iso) const {
    const double s = std::floor(sec + 0.5e-6);
    long long days = (long long)std::floor(s / 86400.0);
    long long rem = (long long)(s - (double)days * 86400.0);
    long long y;
    unsigned m, d;
    civil_from_days(days + days_from_civil(1, 1, 1), y, m, d);
    char buf[64];
    if (iso)
      snprintf(buf, sizeof buf, "%04lld-%02u-%02uT%02lld:%02lld:%02lldZ", y, m, d, rem / 3600, (rem / 60) % 60, rem % 60);
    else
      snprintf(buf, sizeof buf, "%04lld-%02u-%02uT%02lld_%02lld_%02lld", y, m, d, rem / 3600, (rem / 60) % 60, rem % 60);
    return buf;
  }
};

struct Alert {
  double period = 0.0;  // seconds
  double last_time = 0.0;
  bool ring = false;
};

struct TimeManager {
  DateTime start_time, end_time, curr_time;
  double time_step_size = 0.0, elapsed_seconds = 0.0;
  int time_step = 0;
  std::string start_time_format, curr_time_format;
  std::map<std::string, Alert> alerts;

  void init(const Params &p) {  // time_init, src/time_mod.F90:51-79
    if (p.run_days > 0 || p.run_hours > 0 || p.run_minutes > 0) {
      start_time = DateTime();
      end_time.sec = p.run_days * 86400.0 + p.run_hours * 3600.0 + p.run_minutes * 60.0;
    } else {
      if (p.start_time[0] + p.start_time[1] + p.start_time[2] + p.start_time[3] + p.start_time[4] > 0)
        start_time = DateTime::from_civil(p.start_time[0], p.start_time[1], p.start_time[2], p.start_time[3], p.start_time[4]);
      if (p.end_time[0] + p.end_time[1] + p.end_time[2] + p.end_time[3] + p.end_time[4] > 0)
        end_time = DateTime::from_civil(p.end_time[0], p.end_time[1], p.end_time[2], p.end_time[3], p.end_time[4]);
    }
    time_step = 0;
    elapsed_seconds = 0.0;
    time_step_size = p.time_step_size;
    curr_time = start_time;
    start_time_format = start_time.format(false);
    curr_time_format = curr_time.format(false);
  }
  void reset_start_time(const DateTime &t) {  // time_reset_start_time, src/time_mod.F90:81-91
    start_time = t;
    curr_time = t;
    start_time_format = start_time.format(false);
    curr_time_format = curr_time.format(false);
  }
  // "<value> <units>" as in history_periods (src/io_mod.F90:202-222)
  static bool parse_period(const std::string &s, double &seconds, double step_size = 0.0) {
    double v = 0;
    char unit[32] = "";
    if (sscanf(s.c_str(), "%lf %31s", &v, unit) != 2) return false;
    const std::string u = unit;
    if (u == "days") seconds = v * 86400.0;
    else if (u == "hours") seconds = v * 3600.0;
    else if (u == "minutes") seconds = v * 60.0;
    else if (u == "seconds") seconds = v;
    else if (u == "steps") seconds = v * step_size;   // src/io_mod.F90:209-210
    else return false;
    return true;
  }
  void add_alert(const std::string &name, double period_seconds) {  // time_add_alert :144-190
    Alert a;
    a.period = period_seconds;
    a.last_time = start_time.sec;
    alerts[name] = a;
  }
  bool is_alerted(const std::string &name) {  // :192-213
    auto it = alerts.find(name);
    if (it == alerts.end()) return false;
    if (it->second.last_time + it->second.period <= curr_time.sec + 1e-9) {
      it->second.ring = true;
      return true;
    }
    return false;
  }
  void advance() {  // time_advance :106-130
    for (auto &kv : alerts)
      if (kv.second.ring) {
        kv.second.last_time = curr_time.sec;
        kv.second.ring = false;
      }
    time_step++;
    elapsed_seconds += time_step_size;
    curr_time.sec += time_step_size;
    curr_time_format = curr_time.format(true);
  }
  bool is_finished() const { return curr_time.sec >= end_time.sec - 1e-9; }  // :138-142
  // steps until the alert `name` is due (>= 1)
  long steps_until_alert(const std::string &name) const {
    auto it = alerts.find(name);
    if (it == alerts.end()) return 1L << 30;
    const double base = it->second.ring ? curr_time.sec : it->second.last_time;
    const double due = base + it->second.period - curr_time.sec;
    return (long)std::max(1.0, std::ceil(due / time_step_size - 1e-9));
  }
  long steps_until_end() const {
    return (long)std::max(0.0, std::ceil((end_time.sec - curr_time.sec) / time_step_size - 1e-9));
  }
};

}  // namespace host

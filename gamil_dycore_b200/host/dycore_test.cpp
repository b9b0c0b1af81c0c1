// dycore_test.cpp -- `program dycore_test` of the reference (src/dycore_test.F90:1-49):
//   namelist path -> params_read -> dycore_init -> IC plugin by test_case -> dycore_run -> dycore_final
#include <cstdio>
#include <string>

#include "dycore_mod.h"
#include "log.h"

int main(int argc, char **argv) {
  using namespace host;
  if (argc != 2) {
    printf(" Usage: ./dycore_test.exe <namelist_file_path>\n");  // src/dycore_test.F90:15-18
    return 1;
  }
  std::string err;
  if (!params_read(argv[1], params, err)) log_error(err);
  dycore_init();
  if (params.is_restart_run) {
    dycore_restart();
  } else {
    std::string notice;
    if (!set_initial_condition(params, state_ic, notice, err)) {
      printf(" [Error]: %s\n", err.c_str());  // the reference only prints here and carries on (:40-42); we stop
      return 1;
    }
    log_notice(notice);
  }
  dycore_run();
  dycore_final();
  return 0;
}

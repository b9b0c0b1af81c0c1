// dycore_mod.h -- host-side mirror of the reference's public interface (src/dycore_mod.F90:22-25):
//   dycore_init / dycore_restart / dycore_run / dycore_final, parameterless, talking through module-global data
// (here: the singletons below, as src/data_mod.F90:19-22 and src/params_mod.F90 do).  Everything numerical is
// delegated to the CUDA library through the C ABI of include/gmd.h.
#pragma once
#include "params.h"
#include "test_cases.h"
#include "time_manager.h"

namespace host {

extern Params params;      // params_mod globals
extern Fields state_ic;    // state(old)%{u,v,gd}, static%ghs as filled by the IC plugin
extern TimeManager timer;  // time_mod globals

void dycore_init();     // src/dycore_mod.F90:60-111
void dycore_restart();  // :113-117 (restart_read): state + clock from a `<case>.r.<time>.nc` file (host/restart.h)
void dycore_run();      // :119-142
void dycore_final();    // :144-157

}  // namespace host

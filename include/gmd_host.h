/*
 * include/gmd_host.h -- C ABI of the HOST side of the drop-in (no CUDA): the test-case plugins that fill
 * state(old)%{u,v,gd} and static%ghs before dycore_run.
 *
 * Replaces, for callers that are not the C++ `dycore_test` driver (bench.py, the tests, a Python or Fortran host):
 *     rossby_haurwitz_wave_test_set_initial_condition     src/test_cases/barotropic/rossby_haurwitz_wave_test_mod.F90:33-85
 *     steady_geostrophic_flow_test_set_initial_condition  .../steady_geostrophic_flow_test_mod.F90:20-62
 *     mountain_zonal_flow_test_set_initial_condition      .../mountain_zonal_flow_test_mod.F90:28-98
 *     jet_zonal_flow_test_set_initial_condition           .../jet_zonal_flow_test_mod.F90:26-105 (QUADPACK qags restated)
 *     shallow_water_waves_test_set_initial_condition      .../shallow_water_waves_test_mod.F90 (where built)
 * selected by name exactly as src/dycore_test.F90:29-42 does.  Library: gamil_dycore_b200/libgmd_host.so.
 */
#ifndef GMD_HOST_H
#define GMD_HOST_H

#ifdef __cplusplus
extern "C" {
#endif

/* Fills u [num_lat][num_lon], v [num_lat-1][num_lon], gd and ghs [num_lat][num_lon] (GMD_LAYOUT_COMPACT, gmd.h)
   with the named test case's initial condition at the reference's default test-case parameters.
   Returns 0, or 2 (GMD_ERR_ARG) for an unknown test case / bad size; message via gmd_host_last_error. */
int gmd_host_initial_condition(const char *test_case, int num_lon, int num_lat, double *u, double *v, double *gd,
                               double *ghs);
const char *gmd_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif

! gmd_c.F90 -- ISO_C_BINDING interfaces to the C ABI of include/gmd.h (libgmd.so).
!
! This module plus fortran/dycore_mod.F90 is what a maintainer of dongli/gamil-dycore adds to switch the
! barotropic time step to the B200 path: the program `dycore_test`, the namelist, the IC plugins, `history_mod`
! and `log_mod` stay as they are.  NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Fortran compiler exists here,
! SURVEY.md F2); the functionally identical C++ host lives in gamil_dycore_b200/host/.
module gmd_c

  use, intrinsic :: iso_c_binding

  implicit none

  integer(c_int), parameter :: GMD_OK = 0, GMD_ERR_NAN = 1
  integer(c_int), parameter :: GMD_SPLIT_NONE = 0, GMD_SPLIT_CSP1 = 1, GMD_SPLIT_CSP2 = 2, GMD_SPLIT_ISP = 3
  integer(c_int), parameter :: GMD_ADV_CENTER_DIFF = 0, GMD_ADV_UPWIND = 1, GMD_ADV_WENO = 2
  integer(c_int), parameter :: GMD_LAYOUT_COMPACT = 0, GMD_LAYOUT_REFERENCE = 1
  integer(c_int), parameter :: GMD_TIME_PREDICT_CORRECT = 0, GMD_TIME_RUNGE_KUTTA = 1

  ! struct gmd_config (include/gmd.h): numeric keys of /dycore_params/ (src/params_mod.F90:13-98)
  type, bind(c) :: gmd_config
    integer(c_int) num_lon
    integer(c_int) num_lat
    integer(c_int) subcycles
    real(c_double) time_step_size
    integer(c_int) qcon_modified
    integer(c_int) split_scheme
    integer(c_int) uv_adv_scheme
    real(c_double) uv_adv_upwind_lon_beta
    real(c_double) uv_adv_upwind_lat_beta
    integer(c_int) use_zonal_tend_filter
    integer(c_int) zonal_tend_filter_cutoff_wavenumber(20)
    integer(c_int) use_diffusion
    integer(c_int) diffusion_order
    real(c_double) diffusion_coef
    integer(c_int) rank
    integer(c_int) nranks
    integer(c_int) device
    integer(c_int) polar_band_rows
    integer(c_int) time_scheme
    integer(c_int) time_order
    integer(c_int) use_zonal_reduce
    integer(c_int) reduce_adv_lon
    integer(c_int) use_reduce_tend_smooth
    integer(c_int) zonal_reduce_factors(20)
  end type gmd_config

  interface

    subroutine gmd_config_defaults(cfg) bind(c, name='gmd_config_defaults')
      import gmd_config
      type(gmd_config), intent(out) :: cfg
    end subroutine

    integer(c_int) function gmd_create(cfg, model) bind(c, name='gmd_create')
      import gmd_config, c_ptr, c_int
      type(gmd_config), intent(in) :: cfg
      type(c_ptr), intent(out) :: model
    end function

    subroutine gmd_destroy(model) bind(c, name='gmd_destroy')
      import c_ptr
      type(c_ptr), value :: model
    end subroutine

    type(c_ptr) function gmd_last_error() bind(c, name='gmd_last_error')
      import c_ptr
    end function

    ! u, v, gd, ghs: the halo-padded module arrays state(old)%u ... static%ghs passed with c_loc
    ! (layout = GMD_LAYOUT_REFERENCE, src/parallel_mod.F90:283-315)
    integer(c_int) function gmd_set_state(model, u, v, gd, ghs, layout) bind(c, name='gmd_set_state')
      import c_ptr, c_int
      type(c_ptr), value :: model, u, v, gd, ghs
      integer(c_int), value :: layout
    end function

    integer(c_int) function gmd_run_init(model) bind(c, name='gmd_run_init')
      import c_ptr, c_int
      type(c_ptr), value :: model
    end function

    integer(c_int) function gmd_step(model, nsteps) bind(c, name='gmd_step')
      import c_ptr, c_int
      type(c_ptr), value :: model
      integer(c_int), value :: nsteps
    end function

    integer(c_int) function gmd_get_state(model, u, v, gd, layout) bind(c, name='gmd_get_state')
      import c_ptr, c_int
      type(c_ptr), value :: model, u, v, gd
      integer(c_int), value :: layout
    end function

    integer(c_int) function gmd_get_iap_state(model, iap_u, iap_v, iap_gd, layout) bind(c, name='gmd_get_iap_state')
      import c_ptr, c_int
      type(c_ptr), value :: model, iap_u, iap_v, iap_gd
      integer(c_int), value :: layout
    end function

    integer(c_int) function gmd_get_diag(model, total_mass, total_energy, beta) bind(c, name='gmd_get_diag')
      import c_ptr, c_int, c_double
      type(c_ptr), value :: model
      real(c_double), intent(out) :: total_mass, total_energy, beta
    end function

    integer(c_int) function gmd_get_diag_series(model, n, total_mass, total_energy, beta) bind(c, name='gmd_get_diag_series')
      import c_ptr, c_int, c_double
      type(c_ptr), value :: model
      integer(c_int), value :: n
      real(c_double), intent(out) :: total_mass(*), total_energy(*), beta(*)
    end function

    integer(c_int) function gmd_get_vor_div(model, vor, div, layout) bind(c, name='gmd_get_vor_div')
      import c_ptr, c_int
      type(c_ptr), value :: model, vor, div
      integer(c_int), value :: layout
    end function

    ! latitude bands over the GPUs of one node, one MPI rank per GPU (cfg%rank, cfg%nranks): every rank exports a
    ! 256-byte blob (CUDA-IPC handles of its field slab and signal page), the blobs are gathered in rank order
    ! (MPI_Allgather on character(256) buffers) and handed to gmd_peer_connect; halo rows and the two-scalar
    ! all-reduces then travel over NVLink peer memory inside gmd_step
    integer(c_int) function gmd_peer_export(model, blob) bind(c, name='gmd_peer_export')
      import c_ptr, c_int, c_char
      type(c_ptr), value :: model
      character(kind=c_char), intent(out) :: blob(256)
    end function

    integer(c_int) function gmd_peer_connect(model, blobs, nblobs) bind(c, name='gmd_peer_connect')
      import c_ptr, c_int, c_char
      type(c_ptr), value :: model
      character(kind=c_char), intent(in) :: blobs(*)
      integer(c_int), value :: nblobs
    end function

    integer(c_int) function gmd_peer_disconnect(model) bind(c, name='gmd_peer_disconnect')
      import c_ptr, c_int
      type(c_ptr), value :: model
    end function

    integer(c_int) function gmd_get_band(model, row_begin, row_end) bind(c, name='gmd_get_band')
      import c_ptr, c_int
      type(c_ptr), value :: model
      integer(c_int), intent(out) :: row_begin, row_end
    end function

    integer(c_int) function gmd_get_fused_rows(model, row_begin, row_end) bind(c, name='gmd_get_fused_rows')
      import c_ptr, c_int
      type(c_ptr), value :: model
      integer(c_int), intent(out) :: row_begin, row_end
    end function

  end interface

contains

  ! gmd_last_error() as a Fortran string, for log_error
  function gmd_error_message() result(res)
    character(:), allocatable :: res
    character(kind=c_char), pointer :: p(:)
    type(c_ptr) cp
    integer n
    cp = gmd_last_error()
    call c_f_pointer(cp, p, [1024])
    n = 0
    do while (n < 1024)
      if (p(n + 1) == c_null_char) exit
      n = n + 1
    end do
    allocate(character(n) :: res)
    res = transfer(p(1:n), res)
  end function gmd_error_message

end module gmd_c

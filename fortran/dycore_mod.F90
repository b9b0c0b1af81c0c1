! dycore_mod.F90 -- drop-in replacement for src/dycore_mod.F90 of dongli/gamil-dycore (commit 5260294).
!
! Same public interface (dycore_init / dycore_restart / dycore_run / dycore_final, src/dycore_mod.F90:22-25), same
! module-global data (state(old), static, diag; src/data_mod.F90:19-22), same output and log calls.  The step loop
! of dycore_run (:131-140) keeps its shape; `time_integrate` + `diag_run` are replaced by one call of gmd_step,
! and the state is pulled back from the GPU only when an output alert rings.
! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Fortran compiler, SURVEY.md F2); see INTEGRATION.md.
module dycore_mod

  use, intrinsic :: iso_c_binding
  use params_mod, split_scheme_in => split_scheme, uv_adv_scheme_in => uv_adv_scheme
  use data_mod
  use log_mod
  use types_mod
  use mesh_mod
  use time_mod, old => old_time_idx, new => new_time_idx
  use parallel_mod
  use io_mod
  use diag_mod
  use history_mod
  use restart_mod
  use gmd_c

  implicit none

  private

  public dycore_init
  public dycore_restart
  public dycore_run
  public dycore_final

  type(c_ptr) :: model = c_null_ptr

contains

  subroutine check(ierr)
    integer(c_int), intent(in) :: ierr
    if (ierr /= GMD_OK) call log_error(gmd_error_message())
  end subroutine check

  subroutine dycore_init()

    type(gmd_config) cfg

    if (case_name == '') call log_error('case_name is not set!')

    call log_init()
    call mesh_init()
    call time_init()
    call parallel_init()
    call io_init()
    call diag_init()
    call history_init()
    call restart_init()
    call data_init()          ! host copies of state(old), static, diag stay the exchange buffers

    call gmd_config_defaults(cfg)
    cfg%num_lon = num_lon
    cfg%num_lat = num_lat
    cfg%subcycles = subcycles
    cfg%time_step_size = time_step_size
    cfg%qcon_modified = merge(1, 0, qcon_modified)
    select case (time_scheme)
    case ('predict_correct')
      cfg%time_scheme = GMD_TIME_PREDICT_CORRECT
    case ('runge_kutta')
      cfg%time_scheme = GMD_TIME_RUNGE_KUTTA
    case default
      call log_error('Unknown time_scheme ' // trim(time_scheme) // '!')
    end select
    if (time_order /= 0) cfg%time_order = time_order
    select case (split_scheme_in)
    case ('csp1')
      cfg%split_scheme = GMD_SPLIT_CSP1
    case ('csp2')
      cfg%split_scheme = GMD_SPLIT_CSP2
    case ('isp')
      cfg%split_scheme = GMD_SPLIT_ISP
    case default
      cfg%split_scheme = GMD_SPLIT_NONE
      call log_notice('No fast-slow split.')
    end select
    select case (uv_adv_scheme_in)
    case ('center_diff')
      cfg%uv_adv_scheme = GMD_ADV_CENTER_DIFF
    case ('upwind')
      cfg%uv_adv_scheme = GMD_ADV_UPWIND
    case ('weno')
      cfg%uv_adv_scheme = GMD_ADV_WENO
    case default
      call log_error('Unknown uv_adv_scheme ' // trim(uv_adv_scheme_in) // '!')
    end select
    cfg%uv_adv_upwind_lon_beta = uv_adv_upwind_lon_beta
    cfg%uv_adv_upwind_lat_beta = uv_adv_upwind_lat_beta
    cfg%use_zonal_tend_filter = merge(1, 0, use_zonal_tend_filter)
    cfg%zonal_tend_filter_cutoff_wavenumber = zonal_tend_filter_cutoff_wavenumber
    cfg%use_diffusion = merge(1, 0, use_diffusion)
    cfg%diffusion_order = diffusion_order
    cfg%diffusion_coef = diffusion_coef

    call check(gmd_create(cfg, model))

    call log_notice('Dycore module is initialized.')

  end subroutine dycore_init

  subroutine dycore_restart()

    call restart_read(state(old), static)

  end subroutine dycore_restart

  ! state(old) on the host <- device (only when something on the host reads it)
  subroutine pull_state()

    call check(gmd_get_state(model, c_loc(state(old)%u), c_loc(state(old)%v), c_loc(state(old)%gd), GMD_LAYOUT_REFERENCE))
    call check(gmd_get_iap_state(model, c_loc(state(old)%iap%u), c_loc(state(old)%iap%v), c_loc(state(old)%iap%gd), GMD_LAYOUT_REFERENCE))
    call check(gmd_get_vor_div(model, c_loc(diag%vor), c_loc(diag%div), GMD_LAYOUT_REFERENCE))

  end subroutine pull_state

  subroutine dycore_run()

    real(c_double) beta
    integer(c_int) ierr

    ! The IC plugin (or restart_read) has filled state(old)%{u,v,gd} and static%ghs (src/dycore_test.F90:26-43).
    ! reset_cos_lat_at_poles, iap_transform and the first diag_run (src/dycore_mod.F90:121-125) happen on the device.
    call check(gmd_set_state(model, c_loc(state(old)%u), c_loc(state(old)%v), c_loc(state(old)%gd), c_loc(static%ghs), GMD_LAYOUT_REFERENCE))
    call check(gmd_run_init(model))
    call check(gmd_get_diag(model, diag%total_mass, diag%total_energy, beta))
    call output()
    call log_add_diag('total_mass', diag%total_mass)
    call log_add_diag('total_energy', diag%total_energy)
    call log_step()

    do while (.not. time_is_finished())
      ierr = gmd_step(model, 1_c_int)       ! time_integrate + diag_run; batch with nsteps > 1 between alerts if desired
      if (ierr == GMD_ERR_NAN) call log_error(gmd_error_message())   ! 'Total mass is NaN!' (src/diag_mod.F90:79-87)
      call check(ierr)
      call time_advance()
      call check(gmd_get_diag(model, diag%total_mass, diag%total_energy, beta))
      call output()
      call log_add_diag('total_mass', diag%total_mass)
      call log_add_diag('total_energy', diag%total_energy)
      call log_add_diag('beta', beta)
      call log_step()
    end do

  end subroutine dycore_run

  subroutine dycore_final()

    call gmd_destroy(model)
    model = c_null_ptr
    call mesh_final()
    call parallel_final()
    call diag_final()
    call history_final()
    call data_final()

    call log_notice('Dycore module is finalized.')

  end subroutine dycore_final

  subroutine output()

    if (time_is_alerted('hist0.output') .or. time_is_alerted('restart.output')) call pull_state()
    if (time_is_alerted('hist0.output')) call history_write(state(old), static, diag)
    if (time_is_alerted('restart.output')) call restart_write(state(old), static)

  end subroutine output

end module dycore_mod

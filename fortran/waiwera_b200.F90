!   waiwera_b200.F90 -- ISO_C_BINDING interface to libwaiwera_b200.so (include/waiwera_b200.h).
!
!   This is the thin layer the north star asks for: the reference's object-oriented Fortran 2003
!   host code (timestepper.F90, ode.F90 unchanged; flow_simulation.F90 keeps its type and its
!   type-bound hooks) calls the sm_100a CUDA path through these interfaces.  One interface per
!   exported C function, same argument order; arrays are passed as type(c_ptr) obtained with
!   c_loc() on the pointer VecGetArrayF90 / VecGetArrayReadF90 returns (host arrays), or as
!   device addresses when PETSc is built --with-cuda (VecCUDAGetArray).  Every function returns
!   integer(c_int): 0 ok, >0 recoverable physics error (the reference's `err`), <0 fatal
!   (message from wb_last_error()).
!
!   NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler (gfortran,
!   flang, nvfortran and ifx all absent, here and on the GPU box).  The struct layouts below are
!   checked field-by-field against the C header by tests/test_abi.py::test_fortran_interface_matches_header.
module waiwera_b200
  use, intrinsic :: iso_c_binding
  implicit none
  private

  integer(c_int), parameter, public :: WB_THERMO_IAPWS = 0, WB_THERMO_IFC67 = 1
  integer(c_int), parameter, public :: WB_EOS_WE = 0, WB_EOS_W = 1, WB_EOS_WCE = 2, WB_EOS_WAE = 3
  integer(c_int), parameter, public :: WB_RP_FULLY_MOBILE = 0, WB_RP_LINEAR = 1, WB_RP_PICKENS = 2, &
       WB_RP_COREY = 3, WB_RP_GRANT = 4, WB_RP_VAN_GENUCHTEN = 5, WB_RP_TABLE = 6
  integer(c_int), parameter, public :: WB_CP_ZERO = 0, WB_CP_LINEAR = 1, WB_CP_VAN_GENUCHTEN = 2, WB_CP_TABLE = 3
  integer(c_int), parameter, public :: WB_PC_NONE = 0, WB_PC_PBJACOBI = 1, WB_PC_BJACOBI_ILU0 = 2, &
       WB_PC_ASM_ILU0 = 3
  integer(c_int), parameter, public :: WB_KSP_GMRES = 0, WB_KSP_BCGS = 1
  integer(c_int), parameter, public :: WB_METHOD_BEULER = 0, WB_METHOD_BDF2 = 1, WB_METHOD_DIRECTSS = 2
  integer(c_int), parameter, public :: WB_MAX_TRACERS = 3
  integer(c_int), parameter, public :: WB_MAX_TABLE = 16

  type, bind(C), public :: wb_relperm
     integer(c_int) :: type
     real(c_double) :: p(8)
     integer(c_int) :: nl, nv
     real(c_double) :: lx(WB_MAX_TABLE), ly(WB_MAX_TABLE), vx(WB_MAX_TABLE), vy(WB_MAX_TABLE)
  end type wb_relperm

  type, bind(C), public :: wb_cappress
     integer(c_int) :: type
     real(c_double) :: p(8)
     integer(c_int) :: n
     real(c_double) :: x(WB_MAX_TABLE), y(WB_MAX_TABLE)
  end type wb_cappress

  type, bind(C), public :: wb_params
     integer(c_int) :: eos, thermo, extrapolate
     real(c_double) :: pressure_scale, temperature_scale
     real(c_double) :: partial_pressure_scale
     real(c_double) :: eos_w_temperature
     type(wb_relperm) :: relperm
     type(wb_cappress) :: cappress
     real(c_double) :: gravity(3)
  end type wb_params

  type, bind(C), public :: wb_ksp_opts
     integer(c_int) :: type, restart, maxit
     real(c_double) :: rtol, atol, dtol
  end type wb_ksp_opts

  type, bind(C), public :: wb_newton_opts
     integer(c_int) :: max_iterations, min_iterations
     real(c_double) :: rel_tol, abs_tol, update_rel_tol, update_abs_tol
     real(c_double) :: fd_err, fd_umin
     integer(c_int) :: pc_type, pc_nblocks
     type(wb_ksp_opts) :: ksp
  end type wb_newton_opts

  type, bind(C), public :: wb_newton_result
     integer(c_int) :: reason, iterations, linear_iterations
     real(c_double) :: max_residual(32)
     integer(c_int) :: lin_its(32)
     integer(c_int) :: lin_reason(32)
     real(c_double) :: lin_rnorm(32)
  end type wb_newton_result

  public :: wb_last_error, wb_version, wb_create, wb_destroy, wb_num_primary, wb_fluid_dof, wb_set_mesh, &
       wb_jacobian_pattern, wb_jacobian_get, wb_cell_faces_get, wb_comm_unique_id, wb_comm_init, wb_set_halo, wb_set_global_offset, &
       wb_comm_p2p_blob_size, wb_comm_p2p_export, wb_comm_p2p_open, wb_comm_p2p_enabled, wb_comm_p2p_disable, &
       wb_fluid_init, wb_set_boundary, wb_set_boundaries, wb_set_rock, wb_set_sources, wb_set_source_components, wb_set_source_controls, wb_set_source_recharge, wb_get_source_rates, wb_set_source_separators, wb_set_source_pressure_table, wb_separator_stage, wb_get_source_separated, wb_set_method, wb_get_fluid, wb_get_regions, wb_pre_iteration, &
       wb_pre_timestep, wb_pre_retry_timestep, wb_pre_eval, wb_cell_balances, wb_cell_inflows, wb_residual_be, &
       wb_max_scaled, wb_jacobian_be, wb_jacobian_be_colored, wb_fluid_transitions, wb_mat_create, &
       wb_mat_set_values, wb_mat_get_values, wb_mat_destroy, wb_jacobian_mat, wb_mat_mult, wb_pc_setup, wb_pc_refactor, wb_pc_apply, &
       wb_pc_destroy, wb_ksp_solve, wb_ksp_set_check_every, wb_ksp_set_fused, wb_ksp_set_fused_norm, wb_set_pc_blocks, wb_newton_solve_be, wb_timer_get, &
       wb_set_tracers, wb_set_tracer_injection, wb_tracer_cell_balances, wb_tracer_setup_linear, wb_tracer_solve, &
       wb_timer_reset, wb_timers_enable, wb_ksp_fused_profile, wb_launch_count, wb_stream

  interface

     function wb_last_error() bind(C, name="wb_last_error") result(msg)
       import :: c_ptr
       type(c_ptr) :: msg  ! NUL-terminated C string
     end function wb_last_error

     function wb_version() bind(C, name="wb_version") result(v)
       import :: c_int
       integer(c_int) :: v
     end function wb_version

     ! flow_simulation_init / destroy (src/flow_simulation.F90:882, :1049)
     function wb_create(prm, device, ctx) bind(C, name="wb_create") result(ierr)
       import :: c_int, c_ptr, wb_params
       type(wb_params), intent(in) :: prm
       integer(c_int), value :: device
       type(c_ptr), intent(out) :: ctx
       integer(c_int) :: ierr
     end function wb_create

     function wb_destroy(ctx) bind(C, name="wb_destroy") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: ierr
     end function wb_destroy

     function wb_num_primary(ctx) bind(C, name="wb_num_primary") result(n)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: n
     end function wb_num_primary

     function wb_fluid_dof(ctx) bind(C, name="wb_fluid_dof") result(n)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: n
     end function wb_fluid_dof

     ! mesh arrays: mesh%flux_face supports, face_geom, cell_geom, rock Vec arrays
     function wb_set_mesh(ctx, ncell, ninterior, nowned, nface, face_cells, face_geom, cell_geom, rock) &
          bind(C, name="wb_set_mesh") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: ncell, ninterior, nowned, nface
       type(c_ptr), value :: face_cells, face_geom, cell_geom, rock
       integer(c_int) :: ierr
     end function wb_set_mesh

     function wb_jacobian_pattern(ctx, nb, bs, nnzb, rowptr, colidx, vals) bind(C, name="wb_jacobian_pattern") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), intent(out) :: nb, bs, nnzb
       type(c_ptr), intent(out) :: rowptr, colidx, vals
       integer(c_int) :: ierr
     end function wb_jacobian_pattern

     function wb_jacobian_get(ctx, rowptr, colidx, vals) bind(C, name="wb_jacobian_get") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, rowptr, colidx, vals
       integer(c_int) :: ierr
     end function wb_jacobian_get

     ! cell -> face gather lists (order of the face loop's scatter, src/flow_simulation.F90:1410-1458)
     function wb_cell_faces_get(ctx, ncf, cf_ptr, cf_face, cf_other) bind(C, name="wb_cell_faces_get") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, cf_ptr, cf_face, cf_other
       integer(c_int), intent(out) :: ncf
       integer(c_int) :: ierr
     end function wb_cell_faces_get

     function wb_comm_unique_id(id128) bind(C, name="wb_comm_unique_id") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: id128
       integer(c_int) :: ierr
     end function wb_comm_unique_id

     function wb_comm_init(ctx, rank, nranks, id128) bind(C, name="wb_comm_init") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: rank, nranks
       type(c_ptr), value :: id128
       integer(c_int) :: ierr
     end function wb_comm_init

     function wb_set_halo(ctx, nneigh, neigh_rank, send_ptr, send_idx, recv_ptr, recv_idx) &
          bind(C, name="wb_set_halo") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: nneigh
       type(c_ptr), value :: neigh_rank, send_ptr, send_idx, recv_ptr, recv_idx
       integer(c_int) :: ierr
     end function wb_set_halo

     ! NVLink peer-to-peer exchange: export a blob, MPI_Allgather the blobs, open them
     function wb_comm_p2p_blob_size() bind(C, name="wb_comm_p2p_blob_size") result(n)
       import :: c_int
       integer(c_int) :: n
     end function wb_comm_p2p_blob_size

     function wb_comm_p2p_export(ctx, blob) bind(C, name="wb_comm_p2p_export") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, blob
       integer(c_int) :: ierr
     end function wb_comm_p2p_export

     function wb_comm_p2p_open(ctx, blobs) bind(C, name="wb_comm_p2p_open") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, blobs
       integer(c_int) :: ierr
     end function wb_comm_p2p_open

     function wb_comm_p2p_enabled(ctx) bind(C, name="wb_comm_p2p_enabled") result(on)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: on
     end function wb_comm_p2p_enabled

     function wb_comm_p2p_disable(ctx) bind(C, name="wb_comm_p2p_disable") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: ierr
     end function wb_comm_p2p_disable

     function wb_set_global_offset(ctx, first_cell, ncell_global) bind(C, name="wb_set_global_offset") result(ierr)
       import :: c_int, c_int64_t, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int64_t), value :: first_cell, ncell_global
       integer(c_int) :: ierr
     end function wb_set_global_offset

     ! fluid_init (src/flow_simulation.F90:2171-2287)
     function wb_fluid_init(ctx, y, region) bind(C, name="wb_fluid_init") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, y, region
       integer(c_int) :: ierr
     end function wb_fluid_init

     function wb_set_boundary(ctx, ghost_cell, interior_cell, primary, region) bind(C, name="wb_set_boundary") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: ghost_cell, interior_cell
       type(c_ptr), value :: primary
       integer(c_int), value :: region
       integer(c_int) :: ierr
     end function wb_set_boundary

     function wb_set_boundaries(ctx, n, ghost_cells, interior_cells, primary, region) &
          bind(C, name="wb_set_boundaries") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: n
       type(c_ptr), value :: ghost_cells, interior_cells, primary, region
       integer(c_int) :: ierr
     end function wb_set_boundaries

     ! flow_simulation_update_rock_properties (src/flow_simulation.F90:2051-2089): rock records of the interior cells
     function wb_set_rock(ctx, rock) bind(C, name="wb_set_rock") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       type(c_ptr), value :: rock
       integer(c_int) :: ierr
     end function wb_set_rock

     ! source_network%assemble_cell_inflows for fixed-rate sources (src/source.F90:375-480)
     function wb_set_sources(ctx, n, cell, component, rate, enthalpy) bind(C, name="wb_set_sources") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: n
       type(c_ptr), value :: cell, component, rate, enthalpy
       integer(c_int) :: ierr
     end function wb_set_sources

     ! source controls re-evaluated at every function evaluation: deliverability (src/source_control.F90:322-507),
     ! direction (:596-620), total limiter (src/source_network_node.F90:245-315); direction / limit may be c_null_ptr
     ! injection / production component of every source (get_components, src/source_setup.F90:2052-2083)
     function wb_set_source_components(ctx, n, injection_component, production_component) &
          bind(C, name="wb_set_source_components") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, injection_component, production_component
       integer(c_int), value :: n
       integer(c_int) :: ierr
     end function wb_set_source_components

     function wb_set_source_controls(ctx, n, source, productivity, reference_pressure, direction, limit) &
          bind(C, name="wb_set_source_controls") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: n
       type(c_ptr), value :: source, productivity, reference_pressure, direction, limit
       integer(c_int) :: ierr
     end function wb_set_source_controls

     function wb_get_source_rates(ctx, rate) bind(C, name="wb_get_source_rates") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, rate
       integer(c_int) :: ierr
     end function wb_get_source_rates

     ! recharge / injectivity controls: rate = -coefficient (P - reference pressure)
     function wb_set_source_recharge(ctx, n, source, coefficient, reference_pressure) &
          bind(C, name="wb_set_source_recharge") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx
       integer(c_int), value :: n
       type(c_ptr), value :: source, coefficient, reference_pressure
       integer(c_int) :: ierr
     end function wb_set_source_recharge

     ! separators and limiters on the separated water / steam flows (source "separator" / "limiter" input)
     function wb_set_source_separators(ctx, n, source, nstage, pressure, limit_water, limit_steam) &
          bind(C, name="wb_set_source_separators") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx
       integer(c_int), value :: n
       type(c_ptr), value :: source, nstage, pressure, limit_water, limit_steam
       integer(c_int) :: ierr
     end function wb_set_source_separators

     ! reference pressure of sources on deliverability tabulated against the flowing enthalpy or the pressure
     function wb_set_source_pressure_table(ctx, n, source, coordinate, step, npts, table) &
          bind(C, name="wb_set_source_pressure_table") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx
       integer(c_int), value :: n
       type(c_ptr), value :: source, coordinate, step, npts, table
       integer(c_int) :: ierr
     end function wb_set_source_pressure_table

     function wb_separator_stage(ctx, pressure, ref_water_enthalpy, ref_steam_enthalpy) &
          bind(C, name="wb_separator_stage") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       real(c_double), value :: pressure
       real(c_double), intent(out) :: ref_water_enthalpy, ref_steam_enthalpy
       integer(c_int) :: ierr
     end function wb_separator_stage

     function wb_get_source_separated(ctx, out5) bind(C, name="wb_get_source_separated") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx, out5
       integer(c_int) :: ierr
     end function wb_get_source_separated

     ! context%residual selection (src/timestepper.F90:345-452): BE / BDF2 / direct steady state
     function wb_set_method(ctx, method, dt_last, lhs_last2) bind(C, name="wb_set_method") result(ierr)
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: method
       real(c_double), value :: dt_last
       type(c_ptr), value :: lhs_last2
       integer(c_int) :: ierr
     end function wb_set_method

     function wb_get_fluid(ctx, fluid) bind(C, name="wb_get_fluid") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, fluid
       integer(c_int) :: ierr
     end function wb_get_fluid

     function wb_get_regions(ctx, region) bind(C, name="wb_get_regions") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, region
       integer(c_int) :: ierr
     end function wb_get_regions

     ! ode hooks pre_iteration / pre_timestep / pre_retry_timestep (src/flow_simulation.F90:2108, :2022, :2093)
     function wb_pre_iteration(ctx) bind(C, name="wb_pre_iteration") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: ierr
     end function wb_pre_iteration

     function wb_pre_timestep(ctx) bind(C, name="wb_pre_timestep") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: ierr
     end function wb_pre_timestep

     function wb_pre_retry_timestep(ctx) bind(C, name="wb_pre_retry_timestep") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: ierr
     end function wb_pre_retry_timestep

     ! pre_eval (src/flow_simulation.F90:2126)
     function wb_pre_eval(ctx, y, perturbed, nperturbed) bind(C, name="wb_pre_eval") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, y, perturbed
       integer(c_int), value :: nperturbed
       integer(c_int) :: ierr
     end function wb_pre_eval

     ! lhs = cell_balances (:1242), rhs = cell_inflows (:1334)
     function wb_cell_balances(ctx, lhs) bind(C, name="wb_cell_balances") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, lhs
       integer(c_int) :: ierr
     end function wb_cell_balances

     function wb_cell_inflows(ctx, rhs) bind(C, name="wb_cell_inflows") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, rhs
       integer(c_int) :: ierr
     end function wb_cell_inflows

     ! SNES_residual + backwards_Euler_residual (src/timestepper.F90:587, :345)
     function wb_residual_be(ctx, y, lhs_last, dt, perturbed, nperturbed, lhs, rhs, r) &
          bind(C, name="wb_residual_be") result(ierr)
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: ctx, y, lhs_last
       real(c_double), value :: dt
       type(c_ptr), value :: perturbed
       integer(c_int), value :: nperturbed
       type(c_ptr), value :: lhs, rhs, r
       integer(c_int) :: ierr
     end function wb_residual_be

     ! vec_max_pointwise_abs_scale (src/dm_utils.F90:644-685)
     function wb_max_scaled(ctx, v, scale, tol, maxval, maxloc) bind(C, name="wb_max_scaled") result(ierr)
       import :: c_int, c_int64_t, c_double, c_ptr
       type(c_ptr), value :: ctx, v, scale
       real(c_double), value :: tol
       real(c_double), intent(out) :: maxval
       integer(c_int64_t), intent(out) :: maxloc
       integer(c_int) :: ierr
     end function wb_max_scaled

     ! SNESComputeJacobianDefaultColor replacement (src/timestepper.F90:1584-1611)
     function wb_jacobian_be(ctx, y, lhs_last, dt, fd_err, fd_umin, vals_out) bind(C, name="wb_jacobian_be") result(ierr)
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: ctx, y, lhs_last
       real(c_double), value :: dt, fd_err, fd_umin
       type(c_ptr), value :: vals_out
       integer(c_int) :: ierr
     end function wb_jacobian_be

     function wb_jacobian_be_colored(ctx, y, lhs_last, dt, fd_err, fd_umin, vals_out, ncolors) &
          bind(C, name="wb_jacobian_be_colored") result(ierr)
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: ctx, y, lhs_last
       real(c_double), value :: dt, fd_err, fd_umin
       type(c_ptr), value :: vals_out
       integer(c_int), intent(out) :: ncolors
       integer(c_int) :: ierr
     end function wb_jacobian_be_colored

     ! post_linesearch -> fluid_transitions (src/flow_simulation.F90:2419-2576)
     function wb_fluid_transitions(ctx, y_old, search, y, changed_search, changed_y) &
          bind(C, name="wb_fluid_transitions") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, y_old, search, y
       integer(c_int), intent(out) :: changed_search, changed_y
       integer(c_int) :: ierr
     end function wb_fluid_transitions

     ! Mat / PC / KSP (PETSc plug-in seam, src/timestepper.F90:1645-1836)
     function wb_mat_create(ctx, nb, ncolb, bs, nnzb, rowptr, colidx, vals, mat) bind(C, name="wb_mat_create") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: nb, ncolb, bs, nnzb
       type(c_ptr), value :: rowptr, colidx, vals
       type(c_ptr), intent(out) :: mat
       integer(c_int) :: ierr
     end function wb_mat_create

     function wb_mat_set_values(mat, vals) bind(C, name="wb_mat_set_values") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: mat, vals
       integer(c_int) :: ierr
     end function wb_mat_set_values

     function wb_mat_get_values(mat, vals) bind(C, name="wb_mat_get_values") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: mat, vals
       integer(c_int) :: ierr
     end function wb_mat_get_values

     function wb_mat_destroy(mat) bind(C, name="wb_mat_destroy") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: mat
       integer(c_int) :: ierr
     end function wb_mat_destroy

     function wb_jacobian_mat(ctx, mat) bind(C, name="wb_jacobian_mat") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       type(c_ptr), intent(out) :: mat
       integer(c_int) :: ierr
     end function wb_jacobian_mat

     function wb_mat_mult(mat, x, y) bind(C, name="wb_mat_mult") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: mat, x, y
       integer(c_int) :: ierr
     end function wb_mat_mult

     function wb_pc_setup(mat, type, nblocks, block_of_row, pc) bind(C, name="wb_pc_setup") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: mat
       integer(c_int), value :: type, nblocks
       type(c_ptr), value :: block_of_row
       type(c_ptr), intent(out) :: pc
       integer(c_int) :: ierr
     end function wb_pc_setup

     function wb_pc_refactor(pc) bind(C, name="wb_pc_refactor") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: pc
       integer(c_int) :: ierr
     end function wb_pc_refactor

     function wb_pc_apply(pc, r, z) bind(C, name="wb_pc_apply") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: pc, r, z
       integer(c_int) :: ierr
     end function wb_pc_apply

     function wb_pc_destroy(pc) bind(C, name="wb_pc_destroy") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: pc
       integer(c_int) :: ierr
     end function wb_pc_destroy

     function wb_ksp_solve(mat, pc, opts, b, x, its, reason, rnorm) bind(C, name="wb_ksp_solve") result(ierr)
       import :: c_int, c_double, c_ptr, wb_ksp_opts
       type(c_ptr), value :: mat, pc
       type(wb_ksp_opts), intent(in) :: opts
       type(c_ptr), value :: b, x
       integer(c_int), intent(out) :: its, reason
       real(c_double), intent(out) :: rnorm
       integer(c_int) :: ierr
     end function wb_ksp_solve

     function wb_ksp_set_check_every(k) bind(C, name="wb_ksp_set_check_every") result(ierr)
       import :: c_int
       integer(c_int), value :: k
       integer(c_int) :: ierr
     end function wb_ksp_set_check_every

     function wb_set_pc_blocks(ctx, block_of_row) bind(C, name="wb_set_pc_blocks") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, block_of_row
       integer(c_int) :: ierr
     end function wb_set_pc_blocks

     ! setup_tracers (src/tracer.F90:64-150): phase(nt) 1-based phase index; diffusion / decay / activation may be c_null_ptr
     function wb_set_tracers(ctx, nt, phase, diffusion, decay, activation) bind(C, name="wb_set_tracers") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: nt
       type(c_ptr), value :: phase, diffusion, decay, activation
       integer(c_int) :: ierr
     end function wb_set_tracers

     ! source%tracer_injection_rate for every source of wb_set_sources: rate(nt, nsources)
     function wb_set_tracer_injection(ctx, rate) bind(C, name="wb_set_tracer_injection") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, rate
       integer(c_int) :: ierr
     end function wb_set_tracer_injection

     ! aux_lhs (src/flow_simulation.F90:1489-1556)
     function wb_tracer_cell_balances(ctx, al) bind(C, name="wb_tracer_cell_balances") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, al
       integer(c_int) :: ierr
     end function wb_tracer_cell_balances

     ! method%setup_linear (src/timestepper.F90:458-581) + aux_pre_solve (src/flow_simulation.F90:1837-1959)
     function wb_tracer_setup_linear(ctx, dt, al_last, x_last, al_last2, x_last2, x_boundary, al, b, mat) &
          bind(C, name="wb_tracer_setup_linear") result(ierr)
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: ctx
       real(c_double), value :: dt
       type(c_ptr), value :: al_last, x_last, al_last2, x_last2, x_boundary, al, b
       type(c_ptr), intent(out) :: mat
       integer(c_int) :: ierr
     end function wb_tracer_setup_linear

     ! the auxiliary step of timestepper_step (src/timestepper.F90:2347-2353): setup_linear, aux_pre_solve, KSPSolve
     function wb_tracer_solve(ctx, ksp, pc_type, pc_nblocks, dt, al_last, x_last, al_last2, x_last2, x_boundary, al, x, &
          its, reason) bind(C, name="wb_tracer_solve") result(ierr)
       import :: c_int, c_double, c_ptr, wb_ksp_opts
       type(c_ptr), value :: ctx
       type(wb_ksp_opts), intent(in) :: ksp
       integer(c_int), value :: pc_type, pc_nblocks
       real(c_double), value :: dt
       type(c_ptr), value :: al_last, x_last, al_last2, x_last2, x_boundary, al, x
       integer(c_int), intent(out) :: its, reason
       integer(c_int) :: ierr
     end function wb_tracer_solve

     ! SNESSolve as configured by timestepper.F90:1552-1641, one backward-Euler step
     function wb_newton_solve_be(ctx, opts, dt, lhs_last, y, res) bind(C, name="wb_newton_solve_be") result(ierr)
       import :: c_int, c_double, c_ptr, wb_newton_opts, wb_newton_result
       type(c_ptr), value :: ctx
       type(wb_newton_opts), intent(in) :: opts
       real(c_double), value :: dt
       type(c_ptr), value :: lhs_last, y
       type(wb_newton_result), intent(out) :: res
       integer(c_int) :: ierr
     end function wb_newton_solve_be

     ! PetscLogEvent equivalents (src/profiling.F90:42-65)
     function wb_timer_get(ctx, name, ms, count) bind(C, name="wb_timer_get") result(ierr)
       import :: c_int, c_int64_t, c_double, c_char, c_ptr
       type(c_ptr), value :: ctx
       character(kind=c_char), intent(in) :: name(*)  ! NUL-terminated
       real(c_double), intent(out) :: ms
       integer(c_int64_t), intent(out) :: count
       integer(c_int) :: ierr
     end function wb_timer_get

     function wb_ksp_set_fused(on) bind(C, name="wb_ksp_set_fused") result(ierr)
       import :: c_int
       integer(c_int), value :: on
       integer(c_int) :: ierr
     end function wb_ksp_set_fused

     ! 0: VecNorm as a second reduction; 1: norm from the dot-product pass on multi-GPU solves; 2: always
     function wb_ksp_set_fused_norm(mode) bind(C, name="wb_ksp_set_fused_norm") result(ierr)
       import :: c_int
       integer(c_int), value :: mode
       integer(c_int) :: ierr
     end function wb_ksp_set_fused_norm

     ! per-phase device time of the persistent GMRES kernel (nanoseconds, 7 values)
     function wb_ksp_fused_profile(ctx, ns7, reset) bind(C, name="wb_ksp_fused_profile") result(ierr)
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: ctx
       real(c_double), intent(out) :: ns7(7)
       integer(c_int), value :: reset
       integer(c_int) :: ierr
     end function wb_ksp_fused_profile

     function wb_timer_reset(ctx) bind(C, name="wb_timer_reset") result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: ierr
     end function wb_timer_reset

     function wb_timers_enable(on) bind(C, name="wb_timers_enable") result(ierr)
       import :: c_int
       integer(c_int), value :: on
       integer(c_int) :: ierr
     end function wb_timers_enable

     function wb_launch_count(ctx) bind(C, name="wb_launch_count") result(n)
       import :: c_int64_t, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int64_t) :: n
     end function wb_launch_count

     function wb_stream(ctx) bind(C, name="wb_stream") result(s)
       import :: c_ptr
       type(c_ptr), value :: ctx
       type(c_ptr) :: s
     end function wb_stream

  end interface

end module waiwera_b200

!> bvode_module -- Fortran bind(C) interfaces of the B200-native batched VODE integrator
!! (include/bvode.h), plus a thin `bvode_t` type whose methods mirror
!! `dvode_t%initialize / %solve / %dvindy / %dvsrco` of jacobwilliams/dvode
!! (src/dvode_module.F90:276-302, 377, 604, 1878, 3198) for a batch of independent systems.
!!
!! SOURCE ONLY: the build image has no Fortran compiler, so this file is not compiled or
!! tested here; tests/test_cabi_exports.py checks that every symbol named below is exported
!! by libbvode.so with the C signature of include/bvode.h.
!!
!! Differences from the scalar reference a caller must know:
!!   * y(neq, nsys), t(nsys), istate(nsys): one column / entry per system
!!     (Fortran column-major y(neq,nsys) is exactly the C layout y[nsys][neq]).
!!   * rwork(lrw, nsys), iwork(liw, nsys) with lrw >= 20, liw >= 30 hold only the reference's
!!     header slots (optional inputs read from column 1, optional outputs written per system);
!!     the reference's rwork(21:) lives in GPU memory owned by the handle.
!!   * f / jac are __device__ functors compiled into the library and chosen by name.
!!   * REAL32 (the reference's -DREAL32, dvode_kinds_module.F90:29-31): link libbvode32.so and replace c_double
!!     by c_float throughout this file (every `bvode_real` of include/bvode.h).
module bvode_module
   use, intrinsic :: iso_c_binding
   implicit none
   private

   integer(c_int), parameter, public :: BVODE_OK = 0, BVODE_EINVAL = -1, BVODE_ENOTIMPL = -2, &
                                        BVODE_ECUDA = -3, BVODE_ENOMEM = -4, BVODE_EISTATE = -5
   integer(c_int), parameter, public :: BVODE_FLAG_STRICT = 1
   integer(c_int), parameter, public :: BVODE_FLAG_COMPACT = 2
   integer(c_int), parameter, public :: BVODE_FLAG_PER_SYSTEM = 4   !! rtol / atol / optional inputs per system

   interface
      !> A user problem library (csrc/bvode_user_problem.cuh) loaded by the application: pass what its
      !> bvode_user_entry_fast() / bvode_user_entry_strict() / bvode_user_abi_fast() return.
      function bvode_register_problem(entry_fast, entry_strict, abi) bind(C, name='bvode_register_problem') result(id)
         import :: c_int, c_ptr
         type(c_ptr), value :: entry_fast, entry_strict
         integer(c_int), value :: abi
         integer(c_int) :: id
      end function
      function bvode_problem_lookup(name) bind(C, name='bvode_problem_lookup') result(id)
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: name(*)
         integer(c_int) :: id
      end function
      function bvode_problem_info(problem, neq, npar, name) bind(C, name='bvode_problem_info') result(rc)
         import :: c_int, c_ptr
         integer(c_int), value :: problem
         integer(c_int), intent(out) :: neq, npar
         type(c_ptr), intent(out) :: name
         integer(c_int) :: rc
      end function
      function bvode_create(handle, problem, nsys, device, flags) bind(C, name='bvode_create') result(rc)
         import :: c_int, c_ptr
         type(c_ptr), intent(out) :: handle
         integer(c_int), value :: problem, nsys, device, flags
         integer(c_int) :: rc
      end function
      subroutine bvode_destroy(handle) bind(C, name='bvode_destroy')
         import :: c_ptr
         type(c_ptr), value :: handle
      end subroutine
      function bvode_last_error() bind(C, name='bvode_last_error') result(msg)
         import :: c_ptr
         type(c_ptr) :: msg
      end function
      function bvode_set_params(handle, params) bind(C, name='bvode_set_params') result(rc)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         real(c_double), intent(in) :: params(*)      !! params(npar, nsys)
         integer(c_int) :: rc
      end function
      function bvode_solve(handle, neq, y, t, tout, tout_per_system, itol, rtol, atol, itask, &
                           istate, iopt, rwork, lrw, iwork, liw, mf) bind(C, name='bvode_solve') result(rc)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: neq, tout_per_system, itol, itask, iopt, lrw, liw, mf
         real(c_double), intent(inout) :: y(*), t(*), rwork(*)
         real(c_double), intent(in) :: tout(*), rtol(*), atol(*)
         integer(c_int), intent(inout) :: istate(*), iwork(*)
         integer(c_int) :: rc
      end function
      function bvode_solve_schedule(handle, neq, y, t, nout, touts, itol, rtol, atol, itask, istate, &
                                    iopt, rwork, lrw, iwork, liw, mf, y_out) &
            bind(C, name='bvode_solve_schedule') result(rc)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: neq, nout, itol, itask, iopt, lrw, liw, mf
         real(c_double), intent(inout) :: y(*), t(*), rwork(*)
         real(c_double), intent(in) :: touts(*), rtol(*), atol(*)
         integer(c_int), intent(inout) :: istate(*), iwork(*)
         real(c_double), intent(out) :: y_out(*)      !! y_out(neq, nsys, nout)
         integer(c_int) :: rc
      end function
      function bvode_dvindy(handle, t, t_per_system, k, dky, iflag) bind(C, name='bvode_dvindy') result(rc)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         real(c_double), intent(in) :: t(*)
         integer(c_int), value :: t_per_system, k
         real(c_double), intent(out) :: dky(*)        !! dky(neq, nsys)
         integer(c_int), intent(out) :: iflag(*)
         integer(c_int) :: rc
      end function
      function bvode_get_stats(handle, rout, iout) bind(C, name='bvode_get_stats') result(rc)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         real(c_double), intent(out) :: rout(*)       !! rout(4, nsys)  = rwork(11:14)
         integer(c_int), intent(out) :: iout(*)       !! iout(12, nsys) = iwork(11:22)
         integer(c_int) :: rc
      end function
      function bvode_state_bytes(handle) bind(C, name='bvode_state_bytes') result(n)
         import :: c_ptr, c_long_long
         type(c_ptr), value :: handle
         integer(c_long_long) :: n
      end function
      function bvode_dvsrco(handle, sav, job) bind(C, name='bvode_dvsrco') result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle, sav
         integer(c_int), value :: job
         integer(c_int) :: rc
      end function
      function bvode_dvsrco_n(handle, sav, nbytes, job) bind(C, name='bvode_dvsrco_n') result(rc)
         import :: c_int, c_ptr, c_long_long
         type(c_ptr), value :: handle, sav
         integer(c_long_long), value :: nbytes        !! length of sav, checked against bvode_state_bytes()
         integer(c_int), value :: job
         integer(c_int) :: rc
      end function
      function bvode_reset(handle) bind(C, name='bvode_reset') result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int) :: rc
      end function
      !> ---- multi-GPU (include/bvode.h): index sharding and the final gather over NCCL ----
      subroutine bvode_shard_range(nsys_total, rank, world, first, last) bind(C, name='bvode_shard_range')
         import :: c_int, c_long_long
         integer(c_long_long), value :: nsys_total
         integer(c_int), value :: rank, world
         integer(c_long_long), intent(out) :: first, last   !! 0-based [first, last)
      end subroutine
      function bvode_comm_unique_id(id) bind(C, name='bvode_comm_unique_id') result(rc)
         import :: c_int, c_char
         character(kind=c_char), intent(out) :: id(128)     !! rank 0 creates it; MPI_Bcast it to the others
         integer(c_int) :: rc
      end function
      function bvode_comm_create(comm, id, world, rank, device) bind(C, name='bvode_comm_create') result(rc)
         import :: c_int, c_ptr, c_char
         type(c_ptr), intent(out) :: comm
         character(kind=c_char), intent(in) :: id(128)
         integer(c_int), value :: world, rank, device
         integer(c_int) :: rc
      end function
      function bvode_comm_adopt(comm, nccl_comm, world, rank, device) bind(C, name='bvode_comm_adopt') result(rc)
         import :: c_int, c_ptr
         type(c_ptr), intent(out) :: comm
         type(c_ptr), value :: nccl_comm                   !! an ncclComm_t the application made itself
         integer(c_int), value :: world, rank, device
         integer(c_int) :: rc
      end function
      subroutine bvode_comm_destroy(comm) bind(C, name='bvode_comm_destroy')
         import :: c_ptr
         type(c_ptr), value :: comm
      end subroutine
      !> Host arrays on the root: y_all(neq, nsys_total), t_all, istate_all, rout_all(4, :), iout_all(12, :)
      function bvode_gather(handle, comm, root, nsys_total, y_all, t_all, istate_all, rout_all, iout_all) &
            bind(C, name='bvode_gather') result(rc)
         import :: c_int, c_ptr, c_long_long
         type(c_ptr), value :: handle, comm
         integer(c_int), value :: root
         integer(c_long_long), value :: nsys_total
         type(c_ptr), value :: y_all, t_all, istate_all, rout_all, iout_all   !! c_loc(array) or c_null_ptr
         integer(c_int) :: rc
      end function
      !> One process, several GPUs: arrays of nsys_total columns, the arguments of bvode_solve[_schedule]
      function bvode_multi_create(multi, problem, nsys_total, ndev, devices, flags) bind(C, name='bvode_multi_create') result(rc)
         import :: c_int, c_ptr, c_long_long
         type(c_ptr), intent(out) :: multi
         integer(c_int), value :: problem, ndev, flags
         integer(c_long_long), value :: nsys_total
         integer(c_int), intent(in) :: devices(*)
         integer(c_int) :: rc
      end function
      subroutine bvode_multi_destroy(multi) bind(C, name='bvode_multi_destroy')
         import :: c_ptr
         type(c_ptr), value :: multi
      end subroutine
      function bvode_multi_set_params(multi, params) bind(C, name='bvode_multi_set_params') result(rc)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: multi
         real(c_double), intent(in) :: params(*)
         integer(c_int) :: rc
      end function
      function bvode_multi_solve(multi, neq, y, t, tout, tout_per_system, itol, rtol, atol, itask, &
                                 istate, iopt, rwork, lrw, iwork, liw, mf) bind(C, name='bvode_multi_solve') result(rc)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: multi
         integer(c_int), value :: neq, tout_per_system, itol, itask, iopt, lrw, liw, mf
         real(c_double), intent(inout) :: y(*), t(*), rwork(*)
         real(c_double), intent(in) :: tout(*), rtol(*), atol(*)
         integer(c_int), intent(inout) :: istate(*), iwork(*)
         integer(c_int) :: rc
      end function
      function bvode_multi_solve_schedule(multi, neq, y, t, nout, touts, itol, rtol, atol, itask, istate, &
                                          iopt, rwork, lrw, iwork, liw, mf, y_out) &
            bind(C, name='bvode_multi_solve_schedule') result(rc)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: multi
         integer(c_int), value :: neq, nout, itol, itask, iopt, lrw, liw, mf
         real(c_double), intent(inout) :: y(*), t(*), rwork(*)
         real(c_double), intent(in) :: touts(*), rtol(*), atol(*)
         integer(c_int), intent(inout) :: istate(*), iwork(*)
         real(c_double), intent(out) :: y_out(*)
         integer(c_int) :: rc
      end function
      !> xerrwd (dvode_module.F90:3351) as a host-side message stream
      function bvode_istate_summary(handle, counts) bind(C, name='bvode_istate_summary') result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), intent(out) :: counts(8)     !! (ok, istate=-1, ..., istate=-6, other)
         integer(c_int) :: rc
      end function
      function bvode_format_message(handle, sys, buf, buflen) bind(C, name='bvode_format_message') result(n)
         import :: c_int, c_ptr, c_char
         type(c_ptr), value :: handle
         integer(c_int), value :: sys, buflen         !! sys is 0-based
         character(kind=c_char), intent(out) :: buf(*)
         integer(c_int) :: n
      end function
   end interface

   public :: bvode_register_problem, bvode_problem_lookup, bvode_problem_info, bvode_create, bvode_destroy, &
             bvode_last_error, bvode_set_params, bvode_solve, bvode_solve_schedule, &
             bvode_dvindy, bvode_get_stats, bvode_state_bytes, bvode_dvsrco, bvode_dvsrco_n, bvode_reset, &
             bvode_istate_summary, bvode_format_message, bvode_shard_range, bvode_comm_unique_id, &
             bvode_comm_create, bvode_comm_adopt, bvode_comm_destroy, bvode_gather, bvode_multi_create, &
             bvode_multi_destroy, bvode_multi_set_params, bvode_multi_solve, bvode_multi_solve_schedule

   !> Batched counterpart of `type(dvode_t)` (dvode_module.F90:251-314).
   type, public :: bvode_t
      type(c_ptr) :: handle = c_null_ptr
      integer :: nsys = 0, neq = 0, npar = 0
   contains
      procedure :: initialize => bv_initialize   !! dvode_t%initialize(f, jac), :377
      procedure :: solve => bv_solve             !! dvode_t%solve(...), :604
      procedure :: dvindy => bv_dvindy           !! dvode_t%dvindy(...), :1878
      procedure :: dvsrco => bv_dvsrco           !! dvode_t%dvsrco(sav, job), :3198
      procedure :: destroy => bv_destroy
   end type

contains

   !> `call solver%initialize(f=fex, jac=jex)` becomes
   !! `call solver%initialize('robertson', nsys, params)`.
   subroutine bv_initialize(me, problem, nsys, params, device, ierr)
      class(bvode_t), intent(inout) :: me
      character(len=*), intent(in) :: problem
      integer, intent(in) :: nsys
      real(c_double), intent(in) :: params(:, :)        !! (npar, nsys)
      integer, intent(in), optional :: device
      integer, intent(out), optional :: ierr
      integer(c_int) :: id, rc, dev
      type(c_ptr) :: cname
      dev = 0
      if (present(device)) dev = device
      id = bvode_problem_lookup(trim(problem)//c_null_char)
      rc = BVODE_EINVAL
      if (id >= 0) then
         rc = bvode_problem_info(id, me%neq, me%npar, cname)
         if (rc == BVODE_OK) rc = bvode_create(me%handle, id, int(nsys, c_int), dev, 0_c_int)
         if (rc == BVODE_OK) rc = bvode_set_params(me%handle, params)
         if (rc == BVODE_OK) me%nsys = nsys
      end if
      if (present(ierr)) ierr = rc
   end subroutine

   !> Same argument list as dvode_t%solve (dvode_module.F90:604-605); tout is one value for the
   !! whole batch.  Per-system results are in istate(:); ierr reports API misuse only.
   subroutine bv_solve(me, neq, y, t, tout, itol, rtol, atol, itask, istate, iopt, rwork, lrw, &
                       iwork, liw, mf, ierr)
      class(bvode_t), intent(inout) :: me
      integer, intent(in) :: neq, itol, itask, iopt, lrw, liw, mf
      real(c_double), intent(inout) :: y(neq, *), t(*), rwork(lrw, *)
      real(c_double), intent(in) :: tout, rtol(*), atol(*)
      integer(c_int), intent(inout) :: istate(*), iwork(liw, *)
      integer, intent(out), optional :: ierr
      integer(c_int) :: rc
      real(c_double) :: tout1(1)
      tout1(1) = tout
      rc = bvode_solve(me%handle, int(neq, c_int), y, t, tout1, 0_c_int, int(itol, c_int), rtol, atol, &
                       int(itask, c_int), istate, int(iopt, c_int), rwork, int(lrw, c_int), iwork, &
                       int(liw, c_int), int(mf, c_int))
      if (present(ierr)) ierr = rc
   end subroutine

   !> dvindy(t, k, yh, ldyh, dky, iflag): yh / ldyh are the handle's device state.
   subroutine bv_dvindy(me, t, k, dky, iflag, ierr)
      class(bvode_t), intent(inout) :: me
      real(c_double), intent(in) :: t
      integer, intent(in) :: k
      real(c_double), intent(out) :: dky(me%neq, *)
      integer(c_int), intent(out) :: iflag(*)
      integer, intent(out), optional :: ierr
      integer(c_int) :: rc
      real(c_double) :: t1(1)
      t1(1) = t
      rc = bvode_dvindy(me%handle, t1, 0_c_int, int(k, c_int), dky, iflag)
      if (present(ierr)) ierr = rc
   end subroutine

   !> dvsrco(sav, job): sav is a byte buffer of bvode_state_bytes(handle) bytes.
   subroutine bv_dvsrco(me, sav, job, ierr)
      class(bvode_t), intent(inout) :: me
      integer(c_signed_char), intent(inout), target :: sav(*)
      integer, intent(in) :: job
      integer, intent(out), optional :: ierr
      integer(c_int) :: rc
      rc = bvode_dvsrco(me%handle, c_loc(sav), int(job, c_int))
      if (present(ierr)) ierr = rc
   end subroutine

   subroutine bv_destroy(me)
      class(bvode_t), intent(inout) :: me
      if (c_associated(me%handle)) call bvode_destroy(me%handle)
      me%handle = c_null_ptr
   end subroutine

end module bvode_module

!> examples/robertson.f90 -- the reference's example program (test/example.f90:30-77) on the batched
!! solver, through fortran/bvode_module.f90: the same declarations, the same call loop, the same
!! table, for nsys systems at once.  SOURCE ONLY in this repository's image (no Fortran compiler);
!! examples/robertson.c is the same program in C and is built and run by the tests.
!!
!!   gfortran fortran/bvode_module.f90 examples/robertson.f90 -L dvode_b200 -lbvode -Wl,-rpath,$PWD/dvode_b200
program robertson_batch
   use, intrinsic :: iso_c_binding
   use, intrinsic :: iso_fortran_env, only: output_unit
   use bvode_module
   implicit none

   integer, parameter :: nsys = 4, neq = 3, lrw = 20, liw = 30   ! the headers only: work arrays live on the GPU
   real(c_double) :: atol(3), rtol(1), rwork(lrw, nsys), t(nsys), tout, y(neq, nsys), k(3, nsys)
   integer(c_int) :: iwork(liw, nsys), istate(nsys)
   integer :: itol, itask, iopt, mf, iout, i, ierr
   type(bvode_t) :: solver

   do i = 1, nsys                       ! system 1: the reference's rate constants; the others scaled
      k(:, i) = [0.04_c_double, 1.0e4_c_double, 3.0e7_c_double]*(1.0_c_double + 0.01_c_double*(i - 1))
   end do
   y(1, :) = 1.0_c_double
   y(2:3, :) = 0.0_c_double
   t = 0.0_c_double
   tout = 0.4_c_double
   itol = 2
   rtol = 1.0e-4_c_double
   atol = [1.0e-8_c_double, 1.0e-14_c_double, 1.0e-6_c_double]
   itask = 1
   istate = 1
   iopt = 0
   rwork = 0.0_c_double
   iwork = 0
   mf = 21

   call solver%initialize('robertson', nsys, k, ierr=ierr)      ! was: call solver%initialize(f=fex, jac=jex)
   if (ierr /= BVODE_OK) error stop 'bvode: initialize failed'
   do iout = 1, 12
      call solver%solve(neq, y, t, tout, itol, rtol, atol, itask, istate, iopt, rwork, lrw, iwork, liw, mf)
      write (output_unit, '(a,d12.4,a,3d14.6)') ' at t =', t(1), '   y =', y(1, 1), y(2, 1), y(3, 1)
      if (any(istate < 0)) then
         write (output_unit, '(///a,i3)') ' error halt: istate =', minval(istate)
         stop
      end if
      tout = tout*10.0_c_double
   end do

   write (output_unit, '(/a,i4,a,i4,a,i4,a,i4/a,i4/a,i4/a,i4)') &
      ' no. steps =', iwork(11, 1), '   no. f-s =', iwork(12, 1), '   no. j-s =', iwork(13, 1), &
      '   no. lu-s =', iwork(19, 1), '  no. nonlinear iterations =', iwork(20, 1), &
      '  no. nonlinear convergence failures =', iwork(21, 1), '  no. error test failures =', iwork(22, 1)
   call solver%destroy()
end program robertson_batch

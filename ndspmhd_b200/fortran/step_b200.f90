!----------------------------------------------------------------------------
! step_b200.f90 -- drop-in for STEP (src/Makefile:263 selects the integrator the
! same way: STEP=stepND_leapfrog_mhd.f90): the leapfrog of
! src/stepND_leapfrog_mhd.f90:39-300 run on the GPU-resident state.
!
! Between dumps no particle array crosses PCIe: `step` is one C call.  The host
! arrays are refreshed from the device only when something on the host reads
! them -- before `output`/`evwrite` (src/evolve.f90:62,165) -- by calling
! b200_sync_to_host.  The first call uploads the module arrays and runs one
! derivs on the device (the reference enters `step` with the rates of the
! derivs call made by `initialise`, src/initialiseND_mhd.f90:337).
!
! SOURCE FOR A MAINTAINER: this image has no Fortran compiler, so this file is
! exercised through the Python mirror (ndspmhd_b200/lib.py: Hotpath.step) that
! makes the same C calls in the same order (tests/test_gpu_step.py).
!----------------------------------------------------------------------------
subroutine step
 use, intrinsic :: iso_c_binding
 use ndspmhd_b200
 use dimen_mhd, only:ndim
 use part,      only:npart,ntotal,pmass
 use timestep,  only:dt,C_cour,C_force,dtfixed,dtcourant,dtforce,dtav,dtdrag,dtvisc,vsig2max
 implicit none
 type(nd_step_opts) :: so
 type(nd_scalars)   :: s
 type(nd_options)   :: o
 type(nd_arrays)    :: a
 logical, save      :: first = .true.

 call b200_fill_options(o)
 o%device_ghosts = 1                      ! derivs inside the step regenerates the ghosts on the device (src/derivs.f90:78)
 if (first) then
    if (.not.c_associated(b200_ctx)) call b200_check(ndspmhd_b200_create(o,int(ndim,c_int),0_c_int,b200_ctx),'create')
    call b200_fill_arrays(a)
    call b200_check(ndspmhd_b200_upload(b200_ctx,a,int(npart,c_int),int(npart,c_int),int(size(pmass),c_int)),'upload')
    call b200_check(ndspmhd_b200_derivs(b200_ctx,s),'derivs')
    first = .false.
 else
    call b200_check(ndspmhd_b200_set_options(b200_ctx,o),'set_options')
 endif
 so%C_cour = C_cour; so%C_force = C_force; so%reserved = 0
 so%dtfixed = 0; if (dtfixed) so%dtfixed = 1
 ! predictor, boundary, derivs, corrector, boundary, new dt (:253); errors print to iprint and `call quit` like the reference
 call b200_check(ndspmhd_b200_step(b200_ctx,so,dt,s),'step')
 ntotal = s%ntotal
 dtcourant = s%dtcourant; dtforce = s%dtforce; dtav = s%dtav; dtdrag = s%dtdrag; dtvisc = s%dtvisc; vsig2max = s%vsig2max
end subroutine step

!--refresh the module arrays from the device before output/evwrite read them
subroutine b200_sync_to_host
 use, intrinsic :: iso_c_binding
 use ndspmhd_b200
 use part, only:x,vel,hh,en,Bevol,alpha,psi,rho
 implicit none
 type(nd_state_out) :: st
 type(nd_arrays)    :: a
 ! b200_loc_r: c_loc through an assumed-size TARGET dummy (the module arrays have no TARGET attribute)
 st%x = b200_loc_r(x); st%vel = b200_loc_r(vel); st%hh = b200_loc_r(hh); st%en = b200_loc_r(en)
 st%Bevol = b200_loc_r(Bevol); st%alpha = b200_loc_r(alpha); st%psi = b200_loc_r(psi); st%rho = b200_loc_r(rho)
 st%dustevol = c_null_ptr; st%deltav = c_null_ptr
 call b200_check(ndspmhd_b200_download_state(b200_ctx,st,int(size(hh),c_int)),'download_state')
 call b200_fill_arrays(a)                 ! pointers to dens, pr, ..., force, divB, curlB
 call b200_check(ndspmhd_b200_download(b200_ctx,a,15_c_int,int(size(hh),c_int)),'download')
end subroutine b200_sync_to_host

!------------------------------------------------------------------------------!
! Drop-in replacements for the reference's hot-path call sites in `derivs`
! (src/derivs.f90:82, :92, :98, :156): the same argument-less external
! subroutines, defined in THIS file instead of
!     linkND.f90  iterate_density.f90  density_sums.f90  ratesND_mhd.f90
! in the SOURCES list (src/Makefile:231-259) -- the reference's own plug-in
! mechanism (cf. STEP=, SETUP?D=, CONS2PRIM=).
!
! Division of labour in this "host-ghost" mode: the Fortran side keeps doing
! boundary, set_ghost_particles, the integrator and all I/O; the library does
! link + density iteration + cons2prim + rates on the GPU.  State is uploaded in
! set_linklist and downloaded after get_rates; in between it stays resident.
!
! Option tuples outside the compiled set return ND_ERR_UNSUPPORTED_OPTION at
! create time; the shim then calls the original CPU routines, which must be
! kept in the build under the names *_cpu (rename in the four files above).
!------------------------------------------------------------------------------!
subroutine set_linklist
 use ndspmhd_b200
 use dimen_mhd, only:ndim
 use part,      only:npart,ntotal,pmass
 use options,   only:iavlim,idust,icompute_d2v,imhd,iprterm,ibiascorrection
 implicit none
 type(nd_options) :: o
 type(nd_arrays)  :: a
 integer(c_int)   :: ierr
 logical, save    :: cpu_only = .false.
 !--CPU consumers of ll/ifirstincell/iamincell (get_divB, smooth, dust_diffusion, get_2ndderivs) need the host link
 !  list: keep the original routine for those option values (SURVEY.md 8b).  iavlim(3)=2 (get_curl inside
 !  conservative2primitive) runs on the device when B is the evolved variable (imhd >= 11); with B/rho the library
 !  refuses the tuple (ND_ERR_UNSUPPORTED_OPTION below) and the CPU routines take over.
 if (cpu_only .or. idust==3 .or. icompute_d2v>0 .or. imhd<0 .or. iprterm==12 .or. ibiascorrection>0) then
    call set_linklist_cpu
    b200_resident = .false.
    return
 endif
 call b200_fill_options(o)
 if (.not.c_associated(b200_ctx)) then
    ierr = ndspmhd_b200_create(o,int(ndim,c_int),0_c_int,b200_ctx)
    if (ierr == ND_ERR_UNSUPPORTED_OPTION) then
       cpu_only = .true.; call set_linklist_cpu; return
    endif
    call b200_check(ierr,'create')
 else
    ierr = ndspmhd_b200_set_options(b200_ctx,o)
    if (ierr == ND_ERR_UNSUPPORTED_OPTION) then
       cpu_only = .true.; call set_linklist_cpu; return
    endif
    call b200_check(ierr,'set_options')
 endif
 call b200_fill_arrays(a)
 ! size(pmass) = idim, the allocated length (src/allocateND.f90:313)
 call b200_check(ndspmhd_b200_upload(b200_ctx,a,int(npart,c_int),int(ntotal,c_int),int(size(pmass),c_int)),'upload')
 call b200_check(ndspmhd_b200_link(b200_ctx),'set_linklist')
 b200_resident = .true.
end subroutine set_linklist

subroutine iterate_density
 use ndspmhd_b200
 use part,     only:npart,ntotal,hh,pmass
 use bound,    only:hhmax
 use hterms,   only:itsdensity
 implicit none
 type(nd_arrays)  :: a
 type(nd_scalars) :: s
 integer(c_int)   :: ierr, iresume
 if (.not.b200_resident) then
    call iterate_density_cpu
    return
 endif
 iresume = 0
 do
    ierr = ndspmhd_b200_iterate_density(b200_ctx,iresume,s)
    if (ierr /= ND_NEED_RELINK) exit
    !--h grew past hhmax (src/iterate_density.f90:122-126): the host remakes the ghosts with the current h and hands them over
    call b200_fill_arrays(a)
    call b200_check(ndspmhd_b200_download(b200_ctx,a,ND_DL_DENSITY,int(size(pmass),c_int)),'download h')
    call set_ghost_particles
    call b200_fill_arrays(a)          ! alloc() may have moved the arrays
    call b200_check(ndspmhd_b200_update_ghosts(b200_ctx,a,int(ntotal,c_int),int(size(pmass),c_int),hhmax),'update_ghosts')
    iresume = 1
 enddo
 call b200_check(ierr,'iterate_density')
 itsdensity = s%itsdensity
end subroutine iterate_density

!--conservative2primitive is a module procedure (module cons2prim, src/conservative2primitive.f90:42): swap the file
!  through the Makefile's CONS2PRIM variable (src/Makefile:227) for one whose body is:
subroutine conservative2primitive_b200
 use ndspmhd_b200
 implicit none
 if (b200_resident) call b200_check(ndspmhd_b200_cons2prim(b200_ctx),'conservative2primitive')
end subroutine conservative2primitive_b200

subroutine get_rates
 use ndspmhd_b200
 use part,     only:pmass
 use timestep, only:dtcourant,dtforce,dtav,dtdrag,dtvisc,vsig2max
 implicit none
 type(nd_arrays)  :: a
 type(nd_scalars) :: s
 integer(c_int)   :: mask
 if (.not.b200_resident) then
    call get_rates_cpu
    return
 endif
 call b200_check(ndspmhd_b200_get_rates(b200_ctx,s),'get_rates')
 dtcourant = s%dtcourant; dtforce = s%dtforce; dtav = s%dtav; dtdrag = s%dtdrag; dtvisc = s%dtvisc; vsig2max = s%vsig2max
 !--what the integrator (src/stepND_leapfrog_mhd.f90:70-216) and evwrite (src/evwrite_mhd.f90:124-284) read goes back to the
 !  module arrays.  With b200_lean_download (default) the arrays only a dump reads -- gradh, dens, spsound, gradpsi, dudt,
 !  numneigh -- and the ghost rows of every output array stay on the device: 192 instead of 252 bytes per particle and no
 !  ghost rows over PCIe; call b200_sync_to_host (step_b200.f90) before `output` to fetch everything.
 call b200_fill_arrays(a)
 mask = ior(ior(ND_DL_DENSITY,ND_DL_PRIM),ND_DL_RATES)
 if (b200_lean_download) then
    a%gradh = c_null_ptr; a%dens = c_null_ptr; a%spsound = c_null_ptr; a%gradpsi = c_null_ptr
    a%dudt = c_null_ptr; a%numneigh = c_null_ptr
    mask = ior(mask,ND_DL_REAL_ROWS)
 endif
 call b200_check(ndspmhd_b200_download(b200_ctx,a,mask,int(size(pmass),c_int)),'download')
 b200_resident = .false.
end subroutine get_rates

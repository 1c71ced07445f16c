! ------------------------------------------------------------------------------
! mod_blomgpu: ISO_C_BINDING shim between BLOM's Fortran host code and the
! B200 CUDA library (include/blomgpu.h, libblomgpu.so).
!
! Host code stays Fortran: blom_step keeps calling the entry points with the
! reference's own names and argument lists (m,n,mm,nn,k1m,k1n); this module
! provides them and forwards to the C ABI.  Arrays are registered once (the
! module variables of mod_state, mod_grid, ... keep their layout
! a(1-nbdy:idm+nbdy,1-nbdy:jdm+nbdy[,nlev]), phy/mod_xc.F90:45) and stay
! resident on the device; blomgpu_download brings a field back before an
! out-of-scope CPU routine touches it.
!
! NOTE: shipped as source.  The image this library is developed in has no
! Fortran compiler, so this file is not compiled or tested here; every
! interface below is a 1:1 transcription of include/blomgpu.h.
! ------------------------------------------------------------------------------
module mod_blomgpu

   use, intrinsic :: iso_c_binding, only: c_int, c_double, c_char, c_ptr, &
                                          c_null_char, c_loc, c_long, c_int32_t
   use mod_xc, only: xchalt, lp, mnproc

   implicit none
   private

   interface
      integer(c_int) function blomgpu_init(dims, tile, device) bind(C, name='blomgpu_init')
         import :: c_int
         integer(c_int), intent(in) :: dims(8), tile(6)
         integer(c_int), value :: device
      end function
      integer(c_int) function blomgpu_finalize() bind(C, name='blomgpu_finalize')
         import :: c_int
      end function
      function blomgpu_last_error() bind(C, name='blomgpu_last_error') result(msg)
         import :: c_ptr
         type(c_ptr) :: msg
      end function
      integer(c_int) function blomgpu_comm_unique_id(id) bind(C, name='blomgpu_comm_unique_id')
         import :: c_int, c_char
         character(kind=c_char), intent(out) :: id(128)
      end function
      integer(c_int) function blomgpu_comm_init(id, rank, nranks) bind(C, name='blomgpu_comm_init')
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: id(128)
         integer(c_int), value :: rank, nranks
      end function
      integer(c_int) function blomgpu_register(name, host, nlev) bind(C, name='blomgpu_register')
         import :: c_int, c_char, c_double
         character(kind=c_char), intent(in) :: name(*)
         real(c_double), intent(inout) :: host(*)
         integer(c_int), value :: nlev
      end function
      integer(c_int) function blomgpu_register_int(name, host, nlev) bind(C, name='blomgpu_register_int')
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: name(*)
         integer(c_int), intent(inout) :: host(*)
         integer(c_int), value :: nlev
      end function
      integer(c_int) function blomgpu_upload(name) bind(C, name='blomgpu_upload')
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: name(*)
      end function
      integer(c_int) function blomgpu_download(name) bind(C, name='blomgpu_download')
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: name(*)
      end function
      integer(c_int) function blomgpu_download_async(name) bind(C, name='blomgpu_download_async')
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: name(*)
      end function
      integer(c_int) function blomgpu_download_levels_async(name, koff, nlev) bind(C, name='blomgpu_download_levels_async')
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: name(*)
         integer(c_int), value :: koff, nlev
      end function
      integer(c_int) function blomgpu_upload_async(name, koff, nlev) bind(C, name='blomgpu_upload_async')
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: name(*)
         integer(c_int), value :: koff, nlev
      end function
      integer(c_int) function blomgpu_wait_upload(name) bind(C, name='blomgpu_wait_upload')
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: name(*)
      end function
      integer(c_int) function blomgpu_upload_all() bind(C, name='blomgpu_upload_all')
         import :: c_int
      end function
      integer(c_int) function blomgpu_download_all() bind(C, name='blomgpu_download_all')
         import :: c_int
      end function
      integer(c_int) function blomgpu_sync() bind(C, name='blomgpu_sync')
         import :: c_int
      end function
      integer(c_int) function blomgpu_set_option(key, val) bind(C, name='blomgpu_set_option')
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: key(*), val(*)
      end function
      integer(c_int) function blomgpu_set_scalar(key, val) bind(C, name='blomgpu_set_scalar')
         import :: c_int, c_char, c_double
         character(kind=c_char), intent(in) :: key(*)
         real(c_double), value :: val
      end function
      integer(c_int) function blomgpu_get_scalar(key, val) bind(C, name='blomgpu_get_scalar')
         import :: c_int, c_char, c_double
         character(kind=c_char), intent(in) :: key(*)
         real(c_double), intent(out) :: val
      end function
      integer(c_int) function blomgpu_xctilr(name, koff, l1, ld, mh, nh, itype) bind(C, name='blomgpu_xctilr')
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: name(*)
         integer(c_int), value :: koff, l1, ld, mh, nh, itype
      end function
      integer(c_int) function blomgpu_xcsum(name, lev, mask, s) bind(C, name='blomgpu_xcsum')
         import :: c_int, c_char, c_double
         character(kind=c_char), intent(in) :: name(*), mask(*)
         integer(c_int), value :: lev
         real(c_double), intent(out) :: s
      end function
      integer(c_int) function blomgpu_xcmax(name, lev, mask, s) bind(C, name='blomgpu_xcmax')
         import :: c_int, c_char, c_double
         character(kind=c_char), intent(in) :: name(*), mask(*)
         integer(c_int), value :: lev
         real(c_double), intent(out) :: s
      end function
      integer(c_int) function blomgpu_xcmin(name, lev, mask, s) bind(C, name='blomgpu_xcmin')
         import :: c_int, c_char, c_double
         character(kind=c_char), intent(in) :: name(*), mask(*)
         integer(c_int), value :: lev
         real(c_double), intent(out) :: s
      end function
      integer(c_int) function blomgpu_nreg() bind(C, name='blomgpu_nreg')
         import :: c_int
      end function
      integer(c_int) function blomgpu_chksum(name, kcsd, itype, crc) bind(C, name='blomgpu_chksum')
         import :: c_int, c_char, c_int32_t
         character(kind=c_char), intent(in) :: name(*)
         integer(c_int), value :: kcsd, itype
         integer(c_int32_t), intent(out) :: crc
      end function
      integer(c_int) function blomgpu_chksum_at(name, koff, kcsd, itype, crc) bind(C, name='blomgpu_chksum_at')
         import :: c_int, c_char, c_int32_t
         character(kind=c_char), intent(in) :: name(*)
         integer(c_int), value :: koff, kcsd, itype
         integer(c_int32_t), intent(out) :: crc
      end function
      integer(c_int) function blomgpu_bigrid(depth_name) bind(C, name='blomgpu_bigrid')
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: depth_name(*)
      end function
      integer(c_int) function blomgpu_init_cppm() bind(C, name='blomgpu_init_cppm')
         import :: c_int
      end function
      integer(c_int) function blomgpu_inieos() bind(C, name='blomgpu_inieos')
         import :: c_int
      end function
      integer(c_int) function blomgpu_numerical_bounds() bind(C, name='blomgpu_numerical_bounds')
         import :: c_int
      end function
      integer(c_int) function blomgpu_tmsmt1(nn) bind(C, name='blomgpu_tmsmt1')
         import :: c_int
         integer(c_int), value :: nn
      end function
      integer(c_int) function blomgpu_tmsmt2(m, mm, nn, k1m) bind(C, name='blomgpu_tmsmt2')
         import :: c_int
         integer(c_int), value :: m, mm, nn, k1m
      end function
      integer(c_int) function blomgpu_budget_init(mass0) bind(C, name='blomgpu_budget_init')
         import :: c_int, c_double
         real(c_double), intent(out) :: mass0
      end function
      integer(c_int) function blomgpu_budget_sums(ncall, n, nn, out) bind(C, name='blomgpu_budget_sums')
         import :: c_int, c_double
         integer(c_int), value :: ncall, n, nn
         real(c_double), intent(inout) :: out(4)
      end function
   end interface

   ! the nine entry points with the common (m,n,mm,nn,k1m,k1n) signature
   abstract interface
      integer(c_int) function six_int_entry(m, n, mm, nn, k1m, k1n) bind(C)
         import :: c_int
         integer(c_int), value :: m, n, mm, nn, k1m, k1n
      end function
   end interface
   procedure(six_int_entry), bind(C, name='blomgpu_init_fluxes') :: blomgpu_init_fluxes
   procedure(six_int_entry), bind(C, name='blomgpu_difest_halos') :: blomgpu_difest_halos
   procedure(six_int_entry), bind(C, name='blomgpu_eddtra') :: blomgpu_eddtra
   procedure(six_int_entry), bind(C, name='blomgpu_advect') :: blomgpu_advect
   procedure(six_int_entry), bind(C, name='blomgpu_pbcor1') :: blomgpu_pbcor1
   procedure(six_int_entry), bind(C, name='blomgpu_diffus') :: blomgpu_diffus
   procedure(six_int_entry), bind(C, name='blomgpu_pgforc') :: blomgpu_pgforc
   procedure(six_int_entry), bind(C, name='blomgpu_momtum') :: blomgpu_momtum
   procedure(six_int_entry), bind(C, name='blomgpu_barotp') :: blomgpu_barotp
   procedure(six_int_entry), bind(C, name='blomgpu_pbcor2') :: blomgpu_pbcor2
   procedure(six_int_entry), bind(C, name='blomgpu_ndiff') :: blomgpu_ndiff
   procedure(six_int_entry), bind(C, name='blomgpu_cmnfld2') :: blomgpu_cmnfld2
   procedure(six_int_entry), bind(C, name='blomgpu_cmnfld_bfsqf_ale') :: blomgpu_cmnfld_bfsqf_ale
   procedure(six_int_entry), bind(C, name='blomgpu_cmnfld_nslope_ale') :: blomgpu_cmnfld_nslope_ale
   procedure(six_int_entry), bind(C, name='blomgpu_cmnfld_nnslope_ale') :: blomgpu_cmnfld_nnslope_ale

   public :: gpu_setup, gpu_register, gpu_register_int, gpu_upload, gpu_download, gpu_download_async, &
             gpu_download_levels_async, gpu_upload_async, gpu_wait_upload, &
             gpu_option, gpu_scalar, gpu_xctilr, gpu_xcsum, gpu_xcmax, gpu_xcmin, gpu_chksum, gpu_chksum_at, gpu_nreg, &
             init_fluxes, tmsmt1, difest_halos, eddtra, advect, pbcor1, diffus, pgforc, momtum, &
             barotp, pbcor2, tmsmt2, ndiff, cmnfld2, cmnfld_bfsqf_ale, cmnfld_nslope_ale, &
             cmnfld_nnslope_ale, budget_init, budget_sums

contains

   ! -- error convention of the reference: print + xchalt + stop
   !    (phy/mod_advect.F90:166-171)
   subroutine check(rc, where)
      integer(c_int), intent(in) :: rc
      character(len=*), intent(in) :: where
      if (rc /= 0) then
         write (lp,*) 'blomgpu: failure in ', where
         call xchalt('('//where//')')
         stop 'blomgpu'
      end if
   end subroutine check

   pure function cstr(s) result(c)
      character(len=*), intent(in) :: s
      character(kind=c_char, len=len_trim(s)+1) :: c
      c = trim(s)//c_null_char
   end function cstr

   subroutine gpu_setup(itdm, jtdm, kdm, idm, jdm, nbdy, ntr, nreg, i0, j0, ii, jj, device)
      integer, intent(in) :: itdm, jtdm, kdm, idm, jdm, nbdy, ntr, nreg, i0, j0, ii, jj, device
      integer(c_int) :: dims(8), tile(6)
      dims = [itdm, jtdm, kdm, idm, jdm, nbdy, ntr, nreg]
      tile = [i0, j0, ii, jj, mnproc - 1, 1]
      call check(blomgpu_init(dims, tile, int(device, c_int)), 'blomgpu_init')
   end subroutine gpu_setup

   subroutine gpu_register(name, a, nlev)
      character(len=*), intent(in) :: name
      real(c_double), intent(inout) :: a(*)
      integer, intent(in) :: nlev
      call check(blomgpu_register(cstr(name), a, int(nlev, c_int)), 'register '//name)
   end subroutine gpu_register

   subroutine gpu_register_int(name, a, nlev)
      character(len=*), intent(in) :: name
      integer(c_int), intent(inout) :: a(*)
      integer, intent(in) :: nlev
      call check(blomgpu_register_int(cstr(name), a, int(nlev, c_int)), 'register '//name)
   end subroutine gpu_register_int

   subroutine gpu_upload(name)
      character(len=*), intent(in) :: name
      call check(blomgpu_upload(cstr(name)), 'upload '//name)
   end subroutine gpu_upload

   subroutine gpu_download(name)
      character(len=*), intent(in) :: name
      call check(blomgpu_download(cstr(name)), 'download '//name)
   end subroutine gpu_download

   subroutine gpu_download_async(name)
      ! device -> host on the copy stream; the array is valid after the next gpu_sync
      character(len=*), intent(in) :: name
      call check(blomgpu_download_async(cstr(name)), 'download_async '//name)
   end subroutine gpu_download_async

   subroutine gpu_download_levels_async(name, koff, nlev)
      ! levels koff..koff+nlev-1 (e.g. k1n,kk: the new time level) device -> host on the copy stream
      character(len=*), intent(in) :: name
      integer, intent(in) :: koff, nlev
      call check(blomgpu_download_levels_async(cstr(name), koff, nlev), 'download_levels_async '//name)
   end subroutine gpu_download_levels_async

   subroutine gpu_upload_async(name, koff, nlev)
      ! host -> device on the upload stream, overlapping the routines called next
      character(len=*), intent(in) :: name
      integer, intent(in) :: koff, nlev
      call check(blomgpu_upload_async(cstr(name), koff, nlev), 'upload_async '//name)
   end subroutine gpu_upload_async

   subroutine gpu_wait_upload(name)
      ! the library stream waits for the last gpu_upload_async of this field (the host does not block)
      character(len=*), intent(in) :: name
      call check(blomgpu_wait_upload(cstr(name)), 'wait_upload '//name)
   end subroutine gpu_wait_upload

   subroutine gpu_option(key, val)
      character(len=*), intent(in) :: key, val
      call check(blomgpu_set_option(cstr(key), cstr(val)), 'option '//key)
   end subroutine gpu_option

   subroutine gpu_scalar(key, val)
      character(len=*), intent(in) :: key
      real(c_double), intent(in) :: val
      call check(blomgpu_set_scalar(cstr(key), val), 'scalar '//key)
   end subroutine gpu_scalar

   ! xctilr(a(1-nbdy,1-nbdy,koff),l1,ld,mh,nh,itype) on a registered field
   subroutine gpu_xctilr(name, koff, l1, ld, mh, nh, itype)
      character(len=*), intent(in) :: name
      integer, intent(in) :: koff, l1, ld, mh, nh, itype
      call check(blomgpu_xctilr(cstr(name), koff, l1, ld, mh, nh, itype), 'xctilr '//name)
   end subroutine gpu_xctilr

   ! xcsum / xcmax / xcmin of level lev of a registered field over mask 'ip'|'iu'|'iv'|'iq'
   ! (phy/mod_xc.F90:2071, 1157, 1285): same strip order as the reference, bit-reproducible
   subroutine gpu_xcsum(s, name, lev, mask)
      real(c_double), intent(out) :: s
      character(len=*), intent(in) :: name, mask
      integer, intent(in) :: lev
      call check(blomgpu_xcsum(cstr(name), int(lev, c_int), cstr(mask), s), 'xcsum '//name)
   end subroutine gpu_xcsum

   subroutine gpu_xcmax(s, name, lev, mask)
      real(c_double), intent(out) :: s
      character(len=*), intent(in) :: name, mask
      integer, intent(in) :: lev
      call check(blomgpu_xcmax(cstr(name), int(lev, c_int), cstr(mask), s), 'xcmax '//name)
   end subroutine gpu_xcmax

   subroutine gpu_xcmin(s, name, lev, mask)
      real(c_double), intent(out) :: s
      character(len=*), intent(in) :: name, mask
      integer, intent(in) :: lev
      call check(blomgpu_xcmin(cstr(name), int(lev, c_int), cstr(mask), s), 'xcmin '//name)
   end subroutine gpu_xcmin

   ! chksum(a, kcsd, text) of the reference (phy/mod_checksum.F90:41-74) on the device copy; prints the
   ! reference's line ' chksum: <text>: 0x%08X' so that csdiag logs can be diffed
   subroutine gpu_chksum(name, kcsd, itype, text)
      character(len=*), intent(in) :: name, text
      integer, intent(in) :: kcsd, itype
      integer(c_int32_t) :: crc
      call check(blomgpu_chksum(cstr(name), int(kcsd, c_int), int(itype, c_int), crc), 'chksum '//name)
      if (mnproc == 1) write (lp, '(3a,z8.8)') ' chksum: ', text, ': 0x', crc
   end subroutine gpu_chksum

   ! chksum(a(1-nbdy,1-nbdy,koff), kcsd, itype, text), e.g. the k1m level block of the flux arrays
   subroutine gpu_chksum_at(name, koff, kcsd, itype, text)
      character(len=*), intent(in) :: name, text
      integer, intent(in) :: koff, kcsd, itype
      integer(c_int32_t) :: crc
      call check(blomgpu_chksum_at(cstr(name), int(koff, c_int), int(kcsd, c_int), int(itype, c_int), crc), &
                 'chksum '//name)
      if (mnproc == 1) write (lp, '(3a,z8.8)') ' chksum: ', text, ': 0x', crc
   end subroutine gpu_chksum_at

   ! region type found by the device bigrid (0 closed ... 4 periodic in j), phy/mod_xc.F90:54-92
   integer function gpu_nreg()
      gpu_nreg = blomgpu_nreg()
   end function gpu_nreg

   ! -- the reference entry points, same names and argument lists ---------------
   subroutine init_fluxes(m, n, mm, nn, k1m, k1n)   ! phy/mod_state.F90:341
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_init_fluxes(m, n, mm, nn, k1m, k1n), 'init_fluxes')
   end subroutine
   subroutine tmsmt1(nn)                            ! phy/mod_tmsmt.F90:209
      integer, intent(in) :: nn
      call check(blomgpu_tmsmt1(nn), 'tmsmt1')
   end subroutine
   ! halo refreshes of difest_lateral_hybrid / cmnfld2 on the device-resident state
   subroutine difest_halos(m, n, mm, nn, k1m, k1n)  ! phy/mod_difest.F90:826-831, phy/mod_cmnfld_routines.F90:1171
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_difest_halos(m, n, mm, nn, k1m, k1n), 'difest_halos')
   end subroutine
   subroutine eddtra(m, n, mm, nn, k1m, k1n)        ! phy/mod_eddtra.F90:1808
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_eddtra(m, n, mm, nn, k1m, k1n), 'eddtra')
   end subroutine
   subroutine advect(m, n, mm, nn, k1m, k1n)        ! phy/mod_advect.F90:59
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_advect(m, n, mm, nn, k1m, k1n), 'advect')
   end subroutine
   subroutine pbcor1(m, n, mm, nn, k1m, k1n)        ! phy/mod_pbcor.F90:66
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_pbcor1(m, n, mm, nn, k1m, k1n), 'pbcor1')
   end subroutine
   subroutine diffus(m, n, mm, nn, k1m, k1n)        ! phy/mod_diffus.F90:41
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_diffus(m, n, mm, nn, k1m, k1n), 'diffus')
   end subroutine
   subroutine pgforc(m, n, mm, nn, k1m, k1n)        ! phy/mod_pgforc.F90:438
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_pgforc(m, n, mm, nn, k1m, k1n), 'pgforc')
   end subroutine
   subroutine momtum(m, n, mm, nn, k1m, k1n)        ! phy/mod_momtum.F90:215
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_momtum(m, n, mm, nn, k1m, k1n), 'momtum')
   end subroutine
   subroutine barotp(m, n, mm, nn, k1m, k1n)        ! phy/mod_barotp.F90:148
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_barotp(m, n, mm, nn, k1m, k1n), 'barotp')
   end subroutine
   subroutine pbcor2(m, n, mm, nn, k1m, k1n)        ! phy/mod_pbcor.F90:416
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_pbcor2(m, n, mm, nn, k1m, k1n), 'pbcor2')
   end subroutine
   subroutine tmsmt2(m, mm, nn, k1m)                ! phy/mod_tmsmt.F90:281
      integer, intent(in) :: m, mm, nn, k1m
      call check(blomgpu_tmsmt2(m, mm, nn, k1m), 'tmsmt2')
   end subroutine
   ! neutral diffusion over the whole tile: replaces the ndiff_*_jslice calls of the slice pipeline
   ! (phy/mod_ale_regrid_remap.F90:1607-1690, phy/mod_ndiff.F90:959-1175); its slice products are
   ! registered as whole-domain arrays nd_* (include/blomgpu.h)
   subroutine ndiff(m, n, mm, nn, k1m, k1n)
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_ndiff(m, n, mm, nn, k1m, k1n), 'ndiff')
   end subroutine
   ! cmnfld2, hybrid (ALE) branch: halo refresh of temp/saln, filtered buoyancy frequency and the
   ! neutral slope that eddtra/ndiff read (phy/mod_cmnfld_routines.F90:1158-1238, :229-350, :654-883)
   subroutine cmnfld2(m, n, mm, nn, k1m, k1n)
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_cmnfld2(m, n, mm, nn, k1m, k1n), 'cmnfld2')
   end subroutine
   subroutine cmnfld_bfsqf_ale(m, n, mm, nn, k1m, k1n)     ! phy/mod_cmnfld_routines.F90:229
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_cmnfld_bfsqf_ale(m, n, mm, nn, k1m, k1n), 'cmnfld_bfsqf_ale')
   end subroutine
   subroutine cmnfld_nslope_ale(m, n, mm, nn, k1m, k1n)    ! phy/mod_cmnfld_routines.F90:654
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_cmnfld_nslope_ale(m, n, mm, nn, k1m, k1n), 'cmnfld_nslope_ale')
   end subroutine
   subroutine cmnfld_nnslope_ale(m, n, mm, nn, k1m, k1n)   ! phy/mod_cmnfld_routines.F90:813
      integer, intent(in) :: m, n, mm, nn, k1m, k1n
      call check(blomgpu_cmnfld_nnslope_ale(m, n, mm, nn, k1m, k1n), 'cmnfld_nnslope_ale')
   end subroutine
   ! conservation diagnostics, phy/mod_budget.F90:74-196; the caller (mod_budget) stores
   ! out(1:4) into sdp(ncall,n), tdp(ncall,n), trdp(ncall,n) and sc(n)
   subroutine budget_init(mass0)
      real(c_double), intent(out) :: mass0
      call check(blomgpu_budget_init(mass0), 'budget_init')
   end subroutine
   subroutine budget_sums(ncall, n, nn, out)
      integer, intent(in) :: ncall, n, nn
      real(c_double), intent(inout) :: out(4)
      call check(blomgpu_budget_sums(ncall, n, nn, out), 'budget_sums')
   end subroutine

end module mod_blomgpu

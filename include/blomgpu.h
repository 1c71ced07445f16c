/*
 * blomgpu.h — C ABI of the B200-native BLOM horizontal stencil step.
 *
 * The reference (NorESMhub/BLOM) has no FFI / plugin layer: the hot-path
 * routines are Fortran module procedures `X(m,n,mm,nn,k1m,k1n)` that work on
 * module-global arrays (phy/mod_blom_step.F90:146-227).  The drop-in boundary
 * is therefore (entry-point signature + global array layout); this header is
 * what an ISO_C_BINDING shim binds (blom_b200/fortran/mod_blomgpu.F90, and the
 * reference-side stub in INTEGRATION.md).
 *
 * Conventions
 *  - plain C, plain pointers and sizes; every function returns 0 on success,
 *    non-zero on error (the Fortran shim then calls xchalt, matching the
 *    reference's "print + xcstop/xchalt + stop" convention,
 *    phy/mod_advect.F90:166-171).  blomgpu_last_error() gives the message.
 *  - arrays are the reference's own: real(8) a(1-nbdy:idm+nbdy,
 *    1-nbdy:jdm+nbdy [,nlev]), column-major, i fastest (phy/mod_xc.F90:45,
 *    phy/mod_state.F90:34-86).  Host memory stays owned by the caller; the
 *    library keeps a device-resident copy per registered name and only moves
 *    data on explicit upload/download.
 *  - time-level selectors (m,n,mm,nn,k1m,k1n) have the meaning of
 *    phy/mod_blom_step.F90:89-94.
 *  - no CPU fallback: every compute entry point runs CUDA kernels for sm_100a
 *    and fails loudly without a device.
 */
#ifndef BLOMGPU_H
#define BLOMGPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- life cycle ------------------------------------------------------- */
/* dims = {itdm,jtdm,kdm,idm,jdm,nbdy,ntr,nreg}  (dimensions.F, bld/blom_dimensions:150-183)
 * tile = {i0,j0,ii,jj,rank,nranks}: this process' tile in the global grid
 *        (phy/mod_xc.F90:1407-1442); j-band decomposition: i0=0, ii=itdm.
 * device: CUDA device ordinal. */
int blomgpu_init(const int dims[8], const int tile[6], int device);
int blomgpu_finalize(void);
const char* blomgpu_last_error(void);
/* 1 if the library was built with -fmad=false (parity build), else 0 */
int blomgpu_parity_build(void);

/* ---- multi-GPU plumbing (replaces MPI in mod_xc, phy/mod_xc.F90:1332-1700)
 * NCCL bootstrap: rank 0 fills a 128-byte unique id, the host side broadcasts
 * it (any transport), every rank calls comm_init. */
int blomgpu_comm_unique_id(char id[128]);
int blomgpu_comm_init(const char id[128], int rank, int nranks);

/* ---- array registration (module variables of mod_state, mod_grid, ...) -- */
int blomgpu_register(const char* name, double* host, int nlev);
int blomgpu_register_int(const char* name, int* host, int nlev);
int blomgpu_upload(const char* name);      /* host -> device */
int blomgpu_download(const char* name);    /* device -> host */
int blomgpu_upload_all(void);
int blomgpu_download_all(void);
/* device -> host on a second stream, ordered after everything enqueued so far; returns at once
 * and overlaps the routines called afterwards (which must not write the field).  The host array
 * is valid after the next blomgpu_sync().  Used for fields a later out-of-scope CPU routine needs
 * (u,v are final after momtum, phy/mod_blom_step.F90:169-227). */
int blomgpu_download_async(const char* name);
/* the same for levels koff..koff+nlev-1 (1-based) of a field: e.g. the new time level of dp/temp/saln is final
 * after pbcor2 and the mid level after tmsmt2 (phy/mod_pbcor.F90:416, phy/mod_tmsmt.F90:281) */
int blomgpu_download_levels_async(const char* name, int koff, int nlev);
/* host -> device of levels koff..koff+nlev-1 on a third stream: starts when everything enqueued so far has
 * finished and overlaps the routines called next.  Before the first routine that reads the field call
 * blomgpu_wait_upload(name) (the library stream then waits for the copy, the host does not).  Lets the host
 * hand over u,v while tmsmt1..pgforc already run (their first reader is momtum). */
int blomgpu_upload_async(const char* name, int koff, int nlev);
int blomgpu_wait_upload(const char* name);
int blomgpu_sync(void);
/* raw device pointer of a registered/owned array (for zero-copy interop) */
int blomgpu_device_ptr(const char* name, void** dptr, int* nlev);

/* ---- options: namelist strings / scalars (phy/mod_rdlim.F90:137,
 *      phy/mod_time.F90:121-142).  Keys: advmth, pgfmth, mommth, bmcmth, eitmth,
 *      ltedtp, ...; scalars: baclin, batrop, delt1, dlt, lstep, nstep, ... */
int blomgpu_set_option(const char* key, const char* value);
int blomgpu_set_scalar(const char* key, double value);
/* read back a scalar (e.g. btdtmx estimated by numerical_bounds) */
int blomgpu_get_scalar(const char* key, double* value);

/* ---- mod_xc (serial/MPI comm layer) ------------------------------------ */
/* xctilr(a(1-nbdy,1-nbdy,koff),l1,ld,mh,nh,itype)  phy/mod_xc.F90:2342,4222 */
int blomgpu_xctilr(const char* name, int koff, int l1, int ld, int mh, int nh, int itype);
/* xcsum(sum,a(:,:,lev),mask) bit-reproducible order  phy/mod_xc.F90:2071,4116 */
int blomgpu_xcsum(const char* name, int lev, const char* mask, double* sum);
/* xcmax/xcmin over mask==1 interior points of level lev  phy/mod_xc.F90:1157,1285 */
int blomgpu_xcmax(const char* name, int lev, const char* mask, double* out);
int blomgpu_xcmin(const char* name, int lev, const char* mask, double* out);
/* chksum(a,kcsd,itype,text) -> crc  phy/mod_checksum.F90:41-74, phy/mod_xc.F90:2195,4164 */
int blomgpu_chksum(const char* name, int kcsd, int itype, uint32_t* crc);
/* the same with the Fortran actual argument a(1-nbdy,1-nbdy,koff), e.g. chksum(utflld(1-nbdy,1-nbdy,k1m),kk,...)
 * (phy/mod_diffus.F90:175) */
int blomgpu_chksum_at(const char* name, int koff, int kcsd, int itype, uint32_t* crc);

/* ---- setup ------------------------------------------------------------- */
/* bigrid(depth): masks ip,iu,iv,iq (+nreg resolution)  phy/mod_bigrid.F90:44-317 */
int blomgpu_bigrid(const char* depth_name);
int blomgpu_nreg(void);
/* init_cppm  phy/mod_cppm.F90:2504-2746 */
int blomgpu_init_cppm(void);
/* inieos  phy/mod_eos.F90:83-155 */
int blomgpu_inieos(void);
/* numerical_bounds  phy/mod_blom_init.F90:446-555 */
int blomgpu_numerical_bounds(void);
/* init_fluxes  phy/mod_state.F90:341-383 */
int blomgpu_init_fluxes(int m, int n, int mm, int nn, int k1m, int k1n);

/* ---- hot-path entry points, same argument lists as the reference -------- */
int blomgpu_tmsmt1(int nn);                                   /* phy/mod_tmsmt.F90:209 */
/* The halo refreshes of the (out-of-scope) lateral-diffusivity and common-field routines that run between
 * tmsmt1 and eddtra: xctilr of u,v (2*kk levels), ubflxs_p, vbflxs_p, pbu, pbv (2 levels) with (2,2)
 * (phy/mod_difest.F90:826-831) and of temp, saln (2*kk levels) with (3,3)
 * (phy/mod_cmnfld_routines.F90:1171-1172).  momtum and eddtra rely on them; a host that keeps the state on
 * the device calls this where difest_lateral_hybrid / cmnfld2 would have refreshed the host arrays. */
int blomgpu_difest_halos(int m, int n, int mm, int nn, int k1m, int k1n);
int blomgpu_eddtra(int m, int n, int mm, int nn, int k1m, int k1n);  /* phy/mod_eddtra.F90:1808 */
int blomgpu_advect(int m, int n, int mm, int nn, int k1m, int k1n);  /* phy/mod_advect.F90:59 */
int blomgpu_pbcor1(int m, int n, int mm, int nn, int k1m, int k1n);  /* phy/mod_pbcor.F90:66 */
int blomgpu_diffus(int m, int n, int mm, int nn, int k1m, int k1n);  /* phy/mod_diffus.F90:41 */
int blomgpu_pgforc(int m, int n, int mm, int nn, int k1m, int k1n);  /* phy/mod_pgforc.F90:438 */
int blomgpu_momtum(int m, int n, int mm, int nn, int k1m, int k1n);  /* phy/mod_momtum.F90:215 */
int blomgpu_barotp(int m, int n, int mm, int nn, int k1m, int k1n);  /* phy/mod_barotp.F90:148 */
int blomgpu_pbcor2(int m, int n, int mm, int nn, int k1m, int k1n);  /* phy/mod_pbcor.F90:416 */
int blomgpu_tmsmt2(int m, int mm, int nn, int k1m);           /* phy/mod_tmsmt.F90:281 */
/* Neutral diffusion (ltedtp='neutral'): ndiff_prep_jslice, ndiff_uflx_jslice, ndiff_vflx_jslice and
 * ndiff_update_trc_jslice (phy/mod_ndiff.F90:959-1175) over the whole tile, in the order of the slice
 * pipeline that calls them (phy/mod_ale_regrid_remap.F90:1607-1690).  Neutral diffusion inputs are that
 * pipeline's products, registered as whole-domain arrays in the common (i,j,level) layout, T = 2+ntr:
 *   nd_p_src   (kdm+1)    p_src_js(k,i,js)          source interface pressures
 *   nd_ksmx    int (1)    ksmx_js(i,js)             deepest source layer with mass
 *   nd_t_srcdi (2*kdm*T)  t_srcdi_js(is,k,nt,i,js)  level ((nt-1)*kdm+k-1)*2+is
 *   nd_tpc_src (5*kdm*T)  tpc_src_js(c,k,nt,i,js)   level ((nt-1)*kdm+k-1)*5+c
 *   nd_p_dst   (kdm+1)    p_dst_js(k,i,js)          destination interface pressures
 *   nd_trc_rm  (kdm*T)    trc_rm(k,nt,i)            level (nt-1)*kdm+k, updated in place
 *   dpml       (1)        mixed-layer pressure thickness (option ndiff_surface_align, default '1')
 * plus temp, saln, trc, difiso, pu, pv; updates u|v t|s flld, u|v t|s flx (level k+mm), nslpx, nslpy.
 * Library-owned scratch of the call: (8*T + 8)*kdm + 2*(kdm+1) + 4*T*kdm levels (the per-column copy of the inputs the
 * searches read, and the face buffers; 192 bytes per cell and layer for T = 2, 24 GB in all at tnx0.25v4). */
int blomgpu_ndiff(int m, int n, int mm, int nn, int k1m, int k1n);

/* cmnfld2 (phy/mod_cmnfld_routines.F90:1158-1238), hybrid/ALE branch: the producer of nslpx/nslpy that
 * eddtra and ndiff consume, called between tmsmt1 and eddtra (phy/mod_blom_step.F90:136).  Refreshes the
 * temp/saln halos (3,3), then cmnfld_bfsqf_ale (:229-350; bfsqi, bfsqf kdm+1 levels, bfsql kdm) and, for
 * edritp='large scale' or eitmth='gm', cmnfld_nslope_ale (:654-811; phi, nslpx, nslpy, nnslpx, nnslpy)
 * or, with ltedtp='neutral', cmnfld_nnslope_ale (:813-883).  Scalars sls0, bfsqmn default to
 * phy/mod_cmnfld.F90:36,46.  vcoord='isopyc_bulkml' fails with the reference-style message. */
int blomgpu_cmnfld2(int m, int n, int mm, int nn, int k1m, int k1n);
int blomgpu_cmnfld_bfsqf_ale(int m, int n, int mm, int nn, int k1m, int k1n);
int blomgpu_cmnfld_nslope_ale(int m, int n, int mm, int nn, int k1m, int k1n);
int blomgpu_cmnfld_nnslope_ale(int m, int n, int mm, int nn, int k1m, int k1n);

/* Conservation diagnostics (cnsvdi): budget_init (phy/mod_budget.F90:74-93) returns the global mass
 * xcsum(pb(:,:,1)*scp2); budget_sums (phy/mod_budget.F90:95-196) the thickness-weighted global sums at
 * time level nn: out[0]=sdp(ncall,n)  out[1]=tdp(ncall,n)  out[2]=trdp(ncall,n) (1st tracer, ntr>0)
 * out[3]=sc(n) (salt_corr*scp2, only on ncall 4, or 5 for vcoord='isopyc_bulkml', and only when a
 * field 'salt_corr' is registered).  Entries that are not evaluated are left untouched.  Column sums run
 * in k order and the horizontal sum is the strip-ordered xcsum, so the values are bit-reproducible and
 * independent of the j-band decomposition.  Work arrays util1/util2 (mod_utility) are overwritten. */
int blomgpu_budget_init(double* mass0);
int blomgpu_budget_sums(int ncall, int n, int nn, double out[4]);

/* ---- instrumentation ---------------------------------------------------- */
/* kernels launched by this library since the last reset */
long blomgpu_launch_count(void);
void blomgpu_launch_count_reset(void);
/* per-routine device timers (CUDA events on the library stream), the analogue
 * of mod_timing (phy/mod_timing.F90:107-494).  enable!=0 turns them on. */
int blomgpu_timers_enable(int enable);
/* writes up to cap entries; returns number of routines with samples */
int blomgpu_timers_get(int cap, char names[][32], double* ms_total, long* calls, long* launches);
void blomgpu_timers_reset(void);
/* per-kernel device timers: one CUDA-event pair around every launch while enabled
 * (used by bench.py to time the dominant kernel live for the roofline figure). */
int blomgpu_ktimers_enable(int enable);
int blomgpu_ktimers_get(int cap, char names[][64], double* ms_total, long* launches);
/* the CUDA stream (cudaStream_t) all kernels are launched on */
void* blomgpu_stream(void);

#ifdef __cplusplus
}
#endif
#endif /* BLOMGPU_H */

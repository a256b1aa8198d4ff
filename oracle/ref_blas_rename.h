/* TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 * Force-included (gcc -include) when compiling the UNMODIFIED reference sources from
 * /root/reference/src into oracle/_ref/.  The reference calls BLAS/LAPACK through the plain
 * cblas_xxx / xposv_ names (src/cmfrec.h:232-294, 320-372); the only BLAS in this image is the OpenBLAS
 * 0.3.30 that SciPy 1.18.1 bundles, which exports every symbol with a "scipy_" prefix.  This header
 * maps one onto the other; it contains no reference code. */
#ifndef REF_BLAS_RENAME_H
#define REF_BLAS_RENAME_H
#define cblas_ddot scipy_cblas_ddot
#define cblas_dcopy scipy_cblas_dcopy
#define cblas_daxpy scipy_cblas_daxpy
#define cblas_dscal scipy_cblas_dscal
#define cblas_dsyr scipy_cblas_dsyr
#define cblas_dsyrk scipy_cblas_dsyrk
#define cblas_dnrm2 scipy_cblas_dnrm2
#define cblas_dgemm scipy_cblas_dgemm
#define cblas_dgemv scipy_cblas_dgemv
#define cblas_dger scipy_cblas_dger
#define cblas_dsymv scipy_cblas_dsymv
#define dlacpy_ scipy_dlacpy_
#define dposv_ scipy_dposv_
#define dpotrf_ scipy_dpotrf_
#define dpotrs_ scipy_dpotrs_
#define dgelsd_ scipy_dgelsd_
#define cblas_sdot scipy_cblas_sdot
#define cblas_scopy scipy_cblas_scopy
#define cblas_saxpy scipy_cblas_saxpy
#define cblas_sscal scipy_cblas_sscal
#define cblas_ssyr scipy_cblas_ssyr
#define cblas_ssyrk scipy_cblas_ssyrk
#define cblas_snrm2 scipy_cblas_snrm2
#define cblas_sgemm scipy_cblas_sgemm
#define cblas_sgemv scipy_cblas_sgemv
#define cblas_sger scipy_cblas_sger
#define cblas_ssymv scipy_cblas_ssymv
#define slacpy_ scipy_slacpy_
#define sposv_ scipy_sposv_
#define spotrf_ scipy_spotrf_
#define spotrs_ scipy_spotrs_
#define sgelsd_ scipy_sgelsd_
#define openblas_set_num_threads scipy_openblas_set_num_threads
#define openblas_get_num_threads scipy_openblas_get_num_threads
#endif

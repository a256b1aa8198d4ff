"""ctypes signatures of the C ABI declared in include/cmfrec_b200.h.

The reference-named entry points (PART 1 of the header) have exactly the reference's argument lists
(reference src/cmfrec.h:1851-1921, 1152-1179), so the same table also binds the reference's own shared
library -- which is how the tests call the oracle build with identical arguments.
"""
import ctypes as C

import numpy as np

c_bool = C.c_bool
c_int = C.c_int
c_size_t = C.c_size_t
c_void_p = C.c_void_p


def real_ctype(dtype):
    return C.c_float if np.dtype(dtype) == np.float32 else C.c_double


def _P():  # any pointer argument (NULL allowed)
    return c_void_p


def fit_explicit_argtypes(real):
    P = _P()
    return [
        P, P, P, P, P, P, P, P,                    # biasA biasB A B C D Ai Bi
        c_bool, c_bool, c_int,                     # add_implicit_features reset_values seed
        P, P, P,                                   # glob_mean U_colmeans I_colmeans
        c_int, c_int, c_int,                       # m n k
        P, P, P, c_size_t,                         # ixA ixB X nnz
        P, P,                                      # Xfull weight
        c_bool, c_bool, c_bool,                    # user_bias item_bias center
        real, P, real, P,                          # lam lam_unique l1_lam l1_lam_unique
        c_bool, c_bool, c_bool,                    # scale_lam scale_lam_sideinfo scale_bias_const
        P, P,                                      # scaling_biasA scaling_biasB
        P, c_int, c_int,                           # U m_u p
        P, c_int, c_int,                           # II n_i q
        P, P, P, c_size_t,                         # U_row U_col U_sp nnz_U
        P, P, P, c_size_t,                         # I_row I_col I_sp nnz_I
        c_bool, c_bool, c_bool,                    # NA_as_zero_X/U/I
        c_int, c_int, c_int,                       # k_main k_user k_item
        real, real, real, real,                    # w_main w_user w_item w_implicit
        c_int, c_int,                              # niter nthreads
        c_bool, c_bool,                            # verbose handle_interrupt
        c_bool, c_int, c_bool, c_bool,             # use_cg max_cg_steps precondition_cg finalize_chol
        c_bool, c_int, c_bool, c_bool,             # nonneg max_cd_steps nonneg_C nonneg_D
        c_bool, c_bool,                            # precompute_for_predictions include_all_X
        P, P, P, P, P, P, P, P, P,                 # B_plus_bias BtB TransBtBinvBt BtXbias BeTBeChol BiTBi TransCtCinvCt CtCw CtUbias
    ]


def fit_implicit_argtypes(real):
    P = _P()
    return [
        P, P, P, P,                                # A B C D
        c_bool, c_int,                             # reset_values seed
        P, P,                                      # U_colmeans I_colmeans
        c_int, c_int, c_int,                       # m n k
        P, P, P, c_size_t,                         # ixA ixB X nnz
        real, P, real, P,                          # lam lam_unique l1_lam l1_lam_unique
        P, c_int, c_int,                           # U m_u p
        P, c_int, c_int,                           # II n_i q
        P, P, P, c_size_t,                         # U sparse
        P, P, P, c_size_t,                         # I sparse
        c_bool, c_bool,                            # NA_as_zero_U/I
        c_int, c_int, c_int,                       # k_main k_user k_item
        real, real, real,                          # w_main w_user w_item
        P,                                         # w_main_multiplier
        real, c_bool, c_bool,                      # alpha adjust_weight apply_log_transf
        c_int, c_int,                              # niter nthreads
        c_bool, c_bool,                            # verbose handle_interrupt
        c_bool, c_int, c_bool, c_bool,             # use_cg max_cg_steps precondition_cg finalize_chol
        c_bool, c_int, c_bool, c_bool,             # nonneg max_cd_steps nonneg_C nonneg_D
        c_bool,                                    # precompute_for_predictions
        P, P, P, P,                                # BtB BeTBe BeTBeChol CtUbias
    ]


def fit_most_popular_argtypes(real):
    P = _P()
    return [P, P, P, real, real, c_bool, c_bool, real, c_int, c_int, P, P, P, c_size_t, P, P,
            c_bool, c_bool, c_bool, c_bool, c_bool, P, c_int]


def topn_argtypes(real):
    P = _P()
    return [P, c_int, P, c_int, P, real, real, c_int, c_int, P, c_int, P, c_int, P, P, c_int, c_int, c_int]


class AlsOptions(C.Structure):
    """struct cmfb200_als_options; the real_t fields are bound per library in bind_product()."""


def als_options_type(real):
    class _Opt(C.Structure):
        _fields_ = [
            ("implicit", c_int),
            ("m", c_int), ("n", c_int), ("k", c_int),
            ("user_bias", c_int), ("item_bias", c_int),
            ("lam_A", real), ("lam_B", real), ("lam_biasA", real), ("lam_biasB", real),
            ("scale_lam", c_int), ("max_cg_steps", c_int),
            ("rank", c_int), ("world", c_int),
            ("nccl_id", c_void_p), ("stream", c_void_p),
        ]
    return _Opt


REFERENCE_ENTRY_POINTS = ("fit_collective_explicit_als", "fit_collective_implicit_als", "fit_most_popular", "topN",
                          "get_has_openmp", "predict_multiple", "predict_X_old_collective_explicit",
                          "predict_X_old_collective_implicit", "topN_old_collective_explicit", "topN_old_collective_implicit",
                          "factors_collective_explicit_multiple", "factors_collective_implicit_multiple",
                          "precompute_collective_explicit", "precompute_collective_implicit")

PRODUCT_ENTRY_POINTS = REFERENCE_ENTRY_POINTS + (
    "cmfb200_real_name", "cmfb200_device_count", "cmfb200_random_init", "cmfb200_random_init_threads", "cmfb200_coo_to_csr_and_csc",
    "cmfb200_global_mean", "cmfb200_init_biases_twosided", "cmfb200_partition_rows", "cmfb200_nccl_unique_id", "cmfb200_als_create",
    "cmfb200_gram", "cmfb200_set_world", "cmfb200_trim_pool", "cmfb200_debug_poison_smem", "cmfb200_als_destroy", "cmfb200_als_set_factors", "cmfb200_als_get_factors", "cmfb200_als_half_sweep",
    "cmfb200_als_iterate", "cmfb200_als_timed_iterate", "cmfb200_als_set_profile",
    "cmfb200_als_create_from_device_coo", "cmfb200_als_random_factors", "cmfb200_serve_create", "cmfb200_serve_destroy", "cmfb200_serve_predict", "cmfb200_serve_topn", "cmfb200_gemm_nt",
    "cmfb200_als_read_profile", "cmfb200_als_attach_collective", "cmfb200_als_get_collective", "cmfb200_als_sync", "cmfb200_als_launch_count", "cmfb200_als_local_counts",
)


def bind_reference_names(lib, dtype):
    """Attach argtypes/restype for the reference-named entry points to `lib` (product or reference build)."""
    real = real_ctype(dtype)
    lib.fit_collective_explicit_als.argtypes = fit_explicit_argtypes(real)
    lib.fit_collective_explicit_als.restype = c_int
    lib.fit_collective_implicit_als.argtypes = fit_implicit_argtypes(real)
    lib.fit_collective_implicit_als.restype = c_int
    lib.fit_most_popular.argtypes = fit_most_popular_argtypes(real)
    lib.fit_most_popular.restype = c_int
    lib.topN.argtypes = topn_argtypes(real)
    lib.topN.restype = c_int
    lib.get_has_openmp.argtypes = []
    lib.get_has_openmp.restype = c_bool
    P = _P()
    # reference src/cmfrec.h: predict_multiple returns void there; the product returns its error code
    lib.predict_multiple.argtypes = [P, c_int, P, c_int, P, P, real, c_int, c_int, c_int, c_int, P, P, c_size_t, P, c_int]
    lib.predict_multiple.restype = c_int
    lib.predict_X_old_collective_explicit.argtypes = [P, P, P, c_size_t, P, P, P, P, real, c_int, c_int, c_int, c_int, c_int, c_int, c_int]
    lib.predict_X_old_collective_explicit.restype = c_int
    lib.predict_X_old_collective_implicit.argtypes = [P, P, P, c_size_t, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int]
    lib.predict_X_old_collective_implicit.restype = c_int
    lib.topN_old_collective_explicit.argtypes = [P, real, P, P, c_int, P, P, real, c_int, c_int, c_int, c_int, P, c_int, P, c_int, P, P,
                                                 c_int, c_int, c_int, c_bool, c_int]
    lib.topN_old_collective_explicit.restype = c_int
    lib.topN_old_collective_implicit.argtypes = [P, P, c_int, P, c_int, c_int, c_int, c_int, P, c_int, P, c_int, P, P, c_int, c_int, c_int]
    lib.topN_old_collective_implicit.restype = c_int
    lib.factors_collective_explicit_multiple.argtypes = [
        P, P, c_int, P, c_int, c_int, c_bool, c_bool, c_bool,        # A biasA m U m_u p NA_as_zero_U NA_as_zero_X nonneg
        P, P, P, c_size_t, P, P, P,                                  # U_row U_col U_sp nnz_U U_csr_p U_csr_i U_csr
        P, c_int, c_int, P, P, real, P, P,                           # Ub m_ubin pbin C Cb glob_mean biasB U_colmeans
        P, P, P, c_size_t, P, P, P, P, c_int, P, P,                  # X ixA ixB nnz Xcsr_p Xcsr_i Xcsr Xfull n weight B
        P, c_bool, c_int, c_int, c_int, c_int,                       # Bi add_implicit_features k k_user k_item k_main
        real, P, real, P, c_bool, c_bool, c_bool, real,              # lam lam_unique l1_lam l1_lam_unique scale_lam scale_lam_sideinfo scale_bias_const scaling_biasA
        real, real, real, c_int, c_bool,                             # w_main w_user w_implicit n_max include_all_X
        P, P, P, P, P, P, P, P, P, c_int]                            # BtB TransBtBinvBt BtXbias BeTBeChol BiTBi TransCtCinvCt CtCw CtUbias B_plus_bias nthreads
    lib.factors_collective_explicit_multiple.restype = c_int
    lib.factors_collective_implicit_multiple.argtypes = [
        P, c_int, P, c_int, c_int, c_bool, c_bool, P, P, P, c_size_t, P, P, P,   # A m U m_u p NA_as_zero_U nonneg U_row U_col U_sp nnz_U U_csr_*
        P, P, P, c_size_t, P, P, P, P, c_int, P, P,                               # X ixA ixB nnz Xcsr_p Xcsr_i Xcsr B n C U_colmeans
        c_int, c_int, c_int, c_int, real, real, real, real, real, real, c_bool,   # k k_user k_item k_main lam l1_lam alpha w_main w_user w_main_multiplier apply_log_transf
        P, P, P, P, c_int]                                                        # BeTBe BtB BeTBeChol CtUbias nthreads
    lib.factors_collective_implicit_multiple.restype = c_int
    lib.precompute_collective_explicit.argtypes = [
        P, c_int, c_int, c_bool, P, c_int, P, c_bool, P, real, c_bool, P, c_bool, c_int, c_int, c_int, c_int, c_bool, c_bool,
        real, P, c_bool, c_bool, c_bool, real, real, real, real, P, P, P, P, P, P, P, P, P]
    lib.precompute_collective_explicit.restype = c_int
    lib.precompute_collective_implicit.argtypes = [P, c_int, P, c_int, P, c_bool, c_int, c_int, c_int, c_int, real, real, real, real,
                                                   c_bool, c_bool, P, P, P, P]
    lib.precompute_collective_implicit.restype = c_int
    return lib


def bind_product(lib, dtype):
    real = real_ctype(dtype)
    P = c_void_p
    bind_reference_names(lib, dtype)
    lib.cmfb200_real_name.restype = C.c_char_p
    lib.cmfb200_real_name.argtypes = []
    lib.cmfb200_device_count.restype = c_int
    lib.cmfb200_device_count.argtypes = []
    lib.cmfb200_random_init.restype = None
    lib.cmfb200_random_init.argtypes = [P, c_size_t, P, c_size_t, c_int, c_bool]
    lib.cmfb200_random_init_threads.restype = None
    lib.cmfb200_random_init_threads.argtypes = [P, c_size_t, P, c_size_t, c_int, c_bool, c_int]
    lib.cmfb200_coo_to_csr_and_csc.restype = None
    lib.cmfb200_coo_to_csr_and_csc.argtypes = [P, P, P, c_int, c_int, c_size_t, P, P, P, P, P, P]
    lib.cmfb200_global_mean.restype = real
    lib.cmfb200_global_mean.argtypes = [P, c_size_t, c_int]
    lib.cmfb200_init_biases_twosided.restype = None
    lib.cmfb200_init_biases_twosided.argtypes = [c_int, c_int, P, P, P, P, P, P, real, real, c_bool, c_bool, P, P, c_int]
    lib.cmfb200_partition_rows.restype = c_int
    lib.cmfb200_partition_rows.argtypes = [P, c_int, c_int, P, C.POINTER(c_int)]
    lib.cmfb200_nccl_unique_id.restype = c_int
    lib.cmfb200_nccl_unique_id.argtypes = [P]
    lib.AlsOptions = als_options_type(real)
    lib.cmfb200_als_create.restype = c_int
    lib.cmfb200_als_create.argtypes = [C.POINTER(c_void_p), C.POINTER(lib.AlsOptions), P, P, P, P, P, P]
    lib.cmfb200_als_destroy.restype = None
    lib.cmfb200_als_create_from_device_coo.restype = c_int
    lib.cmfb200_als_create_from_device_coo.argtypes = [C.POINTER(c_void_p), C.POINTER(lib.AlsOptions), P, P, P, c_size_t, real, real]
    lib.cmfb200_als_random_factors.restype = c_int
    lib.cmfb200_als_random_factors.argtypes = [P, C.c_ulonglong, real]
    lib.cmfb200_gram.restype = c_int
    lib.cmfb200_gram.argtypes = [P, c_int, c_int, P, c_int, C.POINTER(C.c_float)]
    lib.cmfb200_trim_pool.restype = None
    lib.cmfb200_trim_pool.argtypes = []
    lib.cmfb200_serve_create.restype = c_void_p
    lib.cmfb200_serve_create.argtypes = [P, c_int, c_int, P, c_int, c_int, P, P, real, c_int, c_int, C.POINTER(c_int)]
    lib.cmfb200_serve_destroy.restype = None
    lib.cmfb200_serve_destroy.argtypes = [P]
    lib.cmfb200_serve_predict.restype = c_int
    lib.cmfb200_serve_predict.argtypes = [P, P, P, c_size_t, P]
    lib.cmfb200_serve_topn.restype = c_int
    lib.cmfb200_serve_topn.argtypes = [P, P, c_int, P, P, c_int, P, P, C.POINTER(C.c_float)]
    lib.cmfb200_gemm_nt.restype = c_int
    lib.cmfb200_gemm_nt.argtypes = [P, c_int, c_int, P, c_int, c_int, c_int, P, c_int, C.POINTER(C.c_float)]
    lib.cmfb200_set_world.restype = C.c_int
    lib.cmfb200_set_world.argtypes = [C.c_int, C.c_int, C.c_void_p]
    lib.cmfb200_debug_poison_smem.restype = C.c_int
    lib.cmfb200_debug_poison_smem.argtypes = [C.c_uint]
    lib.cmfb200_als_destroy.argtypes = [P]
    lib.cmfb200_als_set_factors.restype = c_int
    lib.cmfb200_als_set_factors.argtypes = [P, P, P, P, P]
    lib.cmfb200_als_get_factors.restype = c_int
    lib.cmfb200_als_get_factors.argtypes = [P, P, P, P, P]
    lib.cmfb200_als_half_sweep.restype = c_int
    lib.cmfb200_als_half_sweep.argtypes = [P, c_int, c_int, c_int]
    lib.cmfb200_als_iterate.restype = c_int
    lib.cmfb200_als_iterate.argtypes = [P, c_int, c_int, c_int, c_int, c_int]
    lib.cmfb200_als_timed_iterate.restype = c_int
    lib.cmfb200_als_timed_iterate.argtypes = [P, c_int, c_int, c_int, c_int, c_int, C.POINTER(C.c_float)]
    lib.cmfb200_als_set_profile.restype = None
    lib.cmfb200_als_set_profile.argtypes = [P, c_int]
    lib.cmfb200_als_read_profile.restype = c_int
    lib.cmfb200_als_read_profile.argtypes = [P, c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    lib.cmfb200_als_attach_collective.restype = c_int
    lib.cmfb200_als_attach_collective.argtypes = [P, P, c_int, P, c_int, c_int, real, real, real, real, real, real, real]
    lib.cmfb200_als_get_collective.restype = c_int
    lib.cmfb200_als_get_collective.argtypes = [P, P, P, P, P]
    lib.cmfb200_als_sync.restype = c_int
    lib.cmfb200_als_sync.argtypes = [P]
    lib.cmfb200_als_launch_count.restype = C.c_longlong
    lib.cmfb200_als_launch_count.argtypes = [P]
    lib.cmfb200_als_local_counts.restype = None
    lib.cmfb200_als_local_counts.argtypes = [P, P, P, P, P]
    return lib

"""ctypes binding of libcrm_b200.so (C ABI declared in include/crm_b200.h).

There is no CPU fallback: if the library is missing or cannot be loaded the import of the
compute entry points raises, and every call checks the status code and raises CrmError."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcrm_b200.so")

c_double_p = ctypes.c_void_p  # device / host addresses are passed as integers
c_int32_p = ctypes.c_void_p


class CrmError(RuntimeError):
    """A call into libcrm_b200 returned a non-zero status."""

    def __init__(self, status, message):
        super().__init__(f"libcrm_b200 status {status}: {message}")
        self.status = status


class ScanDiag(ctypes.Structure):
    """crm_scan_diag_t"""
    _fields_ = [(name, ctypes.c_void_p) for name in (
        "lml", "delta", "scale", "Q", "lam", "nlam", "M", "liu", "ifault", "flags", "nfev",
        "ov_rho_idx", "ov_v0", "ov_v1")]


# name -> (restype, argtypes); every symbol declared in include/crm_b200.h
SIGNATURES = {
    "crm_version": (ctypes.c_int, []),
    "crm_last_error": (ctypes.c_char_p, []),
    "crm_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "crm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "crm_trim_pool": (ctypes.c_int, [ctypes.c_int]),
    "crm_setup": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_int64, c_double_p, ctypes.c_int64,
                                 c_double_p, ctypes.c_int64, c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.POINTER(ctypes.c_double), ctypes.c_int,
                                 ctypes.c_void_p]),
    "crm_setup_partial": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_int64, c_double_p, ctypes.c_int64,
                                         c_double_p, ctypes.c_int64, c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.POINTER(ctypes.c_double), ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "crm_basis_record_size": (ctypes.c_int64, [ctypes.c_void_p]),
    "crm_export_basis": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p, ctypes.c_void_p]),
    "crm_import_basis": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p, ctypes.c_void_p]),
    "crm_setup_finish": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "crm_set_test_contexts": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_int64, ctypes.c_void_p]),
    "crm_set_background_factors": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                                  ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]),
    "crm_hint_integer_genotypes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "crm_rotation_rows": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "crm_update_phenotype": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_void_p]),
    "crm_set_donors": (ctypes.c_int, [ctypes.c_void_p, c_int32_p, c_int32_p, ctypes.c_int64, ctypes.c_void_p]),
    "crm_get_dims": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]),
    "crm_get_spectrum": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p, ctypes.c_void_p]),
    "crm_scan_interaction": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                            c_double_p, ctypes.c_int64, c_double_p, c_double_p, c_double_p, c_double_p,
                                            c_double_p, ctypes.POINTER(ScanDiag), ctypes.c_void_p]),
    "crm_scan_association": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                            ctypes.c_int, c_double_p, c_double_p, c_double_p, c_double_p, ctypes.c_void_p]),
    "crm_predict_interaction": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, c_double_p,
                                               ctypes.c_int, c_double_p, c_double_p, ctypes.c_int64, c_double_p, ctypes.c_void_p]),
    "crm_launch_count": (ctypes.c_longlong, []),
    "crm_profile": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                   ctypes.POINTER(ctypes.c_int64)]),
    "crm_profile_int8": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
    "crm_stage_genotypes": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    "crm_stage_genotypes_typed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                 ctypes.c_void_p]),
    "crm_host_threads": (ctypes.c_int, []),
    "crm_feeder_blocks": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64), ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]),
    "crm_host_narrow": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "crm_fp64_tensor_peak": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.c_void_p]),
    "crm_eigh_batched": (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, ctypes.POINTER(ctypes.c_double),
                                        ctypes.POINTER(ctypes.c_float), ctypes.c_void_p]),
    "crm_int8_split_gemm": (ctypes.c_int, [c_double_p, ctypes.c_int64, ctypes.c_int64, c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                           ctypes.c_int, c_double_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float),
                                           ctypes.c_void_p]),
    "crm_gemm": (ctypes.c_int, [ctypes.c_int, c_double_p, ctypes.c_int64, ctypes.c_int64, c_double_p, ctypes.c_int64,
                                ctypes.c_int64, c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                ctypes.c_int, ctypes.c_int64, ctypes.c_int64, c_double_p, ctypes.c_int64, ctypes.c_int,
                                ctypes.c_void_p]),
    "crm_lmm_fit_rotated": (ctypes.c_int, [c_double_p] * 8 + [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                              ctypes.c_int64, ctypes.c_double, ctypes.c_int, c_double_p,
                                                              c_double_p, c_double_p, c_double_p, c_int32_p, c_int32_p,
                                                              ctypes.c_void_p]),
    "crm_davies_pvalues": (ctypes.c_int, [c_double_p, c_double_p, c_int32_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                          ctypes.c_double, c_double_p, c_double_p, c_int32_p, c_int32_p, c_double_p,
                                          ctypes.c_void_p]),
    "crm_liu_params": (ctypes.c_int, [c_double_p, c_double_p, c_int32_p, ctypes.c_int, ctypes.c_int64, c_double_p, ctypes.c_void_p]),
    "crm_qmin": (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_int64, c_double_p, ctypes.c_void_p]),
    "crm_lrt_pvalues": (ctypes.c_int, [c_double_p, ctypes.c_double, ctypes.c_int64, c_double_p, ctypes.c_void_p]),
    "crm_lrt_pvalues_dof": (ctypes.c_int, [c_double_p, ctypes.c_double, ctypes.c_int64, ctypes.c_double, c_double_p, ctypes.c_void_p]),
}

_lib = None


def load():
    """Load the shared library (built in-tree by cellregmap_b200.build) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m cellregmap_b200.build` "
            "(cellregmap_b200 has no CPU or PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drift apart
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().crm_last_error()
        text = msg.decode("utf-8", "replace") if msg else ""
        if status == -4:       # CRM_ERR_NONFINITE: glimix_core.lmm.LMM raises ValueError on a non-finite design [W g]
            raise ValueError("There are non-finite values in the covariates matrix. " + text)
        raise CrmError(status, text)


def call(name, *args):
    check(getattr(load(), name)(*args))

"""ctypes binding of libvus.so (include/vus.h).  The CUDA library is the only implementation of the
hot path: if it is missing or no B200 is visible this module raises -- there is no CPU fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvus.so")

c_double_p = C.POINTER(C.c_double)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_u64_p = C.POINTER(C.c_uint64)


KERNEL_CLASSES = ("linearize", "error", "linerr", "assemble", "stereo_assemble", "schur", "bcr_factor", "bcr_solve",
                  "matvec", "border", "vector", "retract")


class LmParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("relative_error_tol", C.c_double), ("absolute_error_tol", C.c_double),
                ("error_tol", C.c_double), ("lambda_initial", C.c_double), ("lambda_factor", C.c_double),
                ("lambda_upper_bound", C.c_double), ("lambda_lower_bound", C.c_double), ("min_model_fidelity", C.c_double),
                ("pcg_max_iterations", C.c_int32), ("pcg_rel_tol", C.c_double), ("max_supernode", C.c_int32),
                ("verbose", C.c_int32), ("profile_kernels", C.c_int32), ("band_chunks", C.c_int32)]


class LmResult(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("inner_iterations", C.c_int32), ("linearizations", C.c_int32),
                ("pcg_iterations", C.c_int32), ("solve_failures", C.c_int32), ("pcg_not_converged", C.c_int32),
                ("initial_error", C.c_double), ("final_error", C.c_double), ("final_lambda", C.c_double),
                ("ms_total", C.c_double), ("ms_linearize", C.c_double), ("ms_assemble", C.c_double),
                ("ms_schur", C.c_double), ("ms_factor", C.c_double), ("ms_pcg", C.c_double), ("ms_update", C.c_double),
                ("kernel_launches", C.c_int64), ("factors_linearized", C.c_int64),
                ("ms_class", C.c_double * 16), ("launches_class", C.c_int64 * 16), ("worst_pcg_rel_residual", C.c_double)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n not in ("reserved", "ms_class", "launches_class")}
        d["ms_class"] = dict(zip(KERNEL_CLASSES, list(self.ms_class)))
        d["launches_class"] = dict(zip(KERNEL_CLASSES, list(self.launches_class)))
        return d


class LmTry(C.Structure):                               # vus_lm_try
    _fields_ = [("lambda_", C.c_double), ("new_error", C.c_double), ("solved", C.c_int32), ("accepted", C.c_int32),
                ("pcg_iterations", C.c_int32), ("reserved", C.c_int32)]


class ComponentResult(C.Structure):                     # vus_component_result
    _fields_ = [("iterations", C.c_int32), ("inner_iterations", C.c_int32), ("initial_error", C.c_double),
                ("final_error", C.c_double), ("final_lambda", C.c_double)]


COMM_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64)     # vus_comm_fn
COMM_ALLREDUCE_SUM, COMM_HALO = 0, 1

EXPORTS = {
    "vus_default_lm_params": (None, [C.POINTER(LmParams)]),
    "vus_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "vus_destroy": (None, [C.c_void_p]),
    "vus_last_error": (C.c_char_p, [C.c_void_p]),
    "vus_set_variables": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_u64_p, C.c_void_p, C.c_int]),
    "vus_get_variables": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "vus_save_values": (C.c_int, [C.c_void_p]),
    "vus_restore_values": (C.c_int, [C.c_void_p]),
    "vus_add_factors": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_i32_p, C.c_void_p, C.c_void_p, c_i64_p, C.c_int]),
    "vus_set_calibration": (C.c_int, [C.c_void_p, c_double_p]),
    "vus_set_gravity": (C.c_int, [C.c_void_p, c_double_p]),
    "vus_set_lm_params": (C.c_int, [C.c_void_p, C.POINTER(LmParams)]),
    "vus_set_gtsam_build": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vus_set_partition": (C.c_int, [C.c_void_p, C.c_int64, c_i64_p]),
    "vus_set_partition_chain": (C.c_int, [C.c_void_p, C.c_int32, c_i64_p, c_i64_p, C.c_int32, C.c_int32]),
    "vus_set_comm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vus_set_comm_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "vus_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "vus_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "vus_set_halo": (C.c_int, [C.c_void_p, C.c_int32, c_i32_p, c_i64_p, c_i32_p, c_i64_p, c_i64_p]),
    "vus_set_components": (C.c_int, [C.c_void_p, C.c_int64, c_i64_p]),
    "vus_get_component_results": (C.c_int, [C.c_void_p, C.POINTER(ComponentResult)]),
    "vus_analyze": (C.c_int, [C.c_void_p]),
    "vus_get_layout": (C.c_int, [C.c_void_p, c_i64_p]),
    "vus_optimize": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(LmResult)]),
    "vus_get_trace": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(LmTry), c_i32_p]),
    "vus_error": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p]),
    "vus_factor_errors": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p]),
    "vus_linearize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_double_p, c_double_p]),
    "vus_solve_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, c_double_p, c_double_p, c_double_p, c_double_p, c_i32_p]),
    "vus_marginal_covariance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, c_i32_p, c_i32_p, c_double_p]),
    "vus_preintegrate_imu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, c_double_p,
                                       c_double_p, c_double_p, c_double_p, C.c_void_p, C.c_void_p, C.c_int]),
    "vus_backproject_stereo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, c_i32_p, C.c_void_p, C.c_void_p, C.c_int]),
    "vus_debug_band_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, c_double_p, c_double_p, c_double_p, C.c_int]),
    "vus_time_linearize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_double_p]),
}


def bind(path):
    """dlopen `path` and attach the prototypes of every symbol include/vus.h declares."""
    lib = C.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def load():
    """The product library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a). There is no CPU fallback for this path.")
        _lib = bind(LIB_PATH)
    return _lib

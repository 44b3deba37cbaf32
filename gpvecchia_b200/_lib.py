"""ctypes binding of libgpvecchia_b200.so (include/gpvecchia_b200.h).

The shared library is the product; this module only declares its C signatures.  There is no
Python/NumPy/PyTorch compute fallback: if the library is missing, import fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GPV_LIB_PATH: development override used by tools/kbench.py to compare kernel variants
LIB_PATH = os.environ.get("GPV_LIB_PATH") or os.path.join(_HERE, "libgpvecchia_b200.so")

GPV_OK = 0
GPV_ERR_ARG, GPV_ERR_CUDA, GPV_ERR_COVTYPE, GPV_ERR_NOMEM, GPV_ERR_UNSUPPORTED = 1, 2, 3, 4, 5
GPV_COND_RLOGICAL_I32, GPV_COND_F64 = 0, 1

# every symbol include/gpvecchia_b200.h declares (tests check the .so exports all of them)
EXPORTED = [
    "gpv_last_error", "gpv_version", "gpv_device_count", "gpv_host_alloc", "gpv_host_free", "gpv_release_cached", "gpv_create", "gpv_create_shard", "gpv_destroy",
    "gpv_set_revcond", "gpv_u_nzentries", "gpv_packed_len", "gpv_nuggets_read", "gpv_u_values_packed",
    "gpv_u_nzentries_mat", "gpv_u_values_packed_mat", "gpv_csc_dims", "gpv_u_sparsity", "gpv_u_csc_pattern", "gpv_u_values_csc",
    "gpv_multi_csc_dims", "gpv_multi_u_csc_pattern", "gpv_multi_u_values_csc",
    "gpv_loglik_numerator", "gpv_loglik_z", "gpv_set_scalar_nugget", "gpv_u_dev", "gpv_last_kernel_ms", "gpv_kernel_time_stats", "gpv_last_kernel_name",
    "gpv_launch_count", "gpv_U_NZentries", "gpv_MaternFun", "gpv_EsqeFun",
    "gpv_measure_fp64_peak", "gpv_measure_copy_bw", "gpv_harness_ordered_nn", "gpv_whichCondOnLatent",
    "gpv_ic0", "gpv_createUcppM", "gpv_createUcpp",
    "gpv_multi_create", "gpv_multi_destroy", "gpv_multi_num_devices", "gpv_multi_packed_len",
    "gpv_multi_row_cuts", "gpv_multi_set_revcond", "gpv_multi_u_values_packed",
    "gpv_multi_loglik_numerator", "gpv_multi_loglik_z", "gpv_set_last_error",
    "gpv_dist_unique_id", "gpv_dist_init", "gpv_dist_finalize", "gpv_loglik_z_dist",
]


class GpvError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"gpvecchia_b200 status {status}: {message}")
        self.status = status


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`make -C gpvecchia_b200/csrc -j8` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "gpvecchia_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    cp = C.c_char_p
    L.gpv_last_error.restype = cp
    L.gpv_version.restype = cp
    L.gpv_device_count.restype = i32
    L.gpv_host_alloc.argtypes = [C.c_size_t]
    L.gpv_host_alloc.restype = vp
    L.gpv_host_free.argtypes = [vp]
    L.gpv_host_free.restype = None
    L.gpv_release_cached.argtypes = []
    L.gpv_release_cached.restype = None
    L.gpv_create.argtypes = [C.POINTER(vp), i64, i32, i32, vp, vp, vp, i32, vp, i64, i64, i32]
    L.gpv_create.restype = i32
    L.gpv_create_shard.argtypes = [C.POINTER(vp), i64, i32, i32, vp, vp, vp, i32, vp, i64, i64, i32]
    L.gpv_create_shard.restype = i32
    L.gpv_kernel_time_stats.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(dbl)]
    L.gpv_kernel_time_stats.restype = i32
    L.gpv_destroy.argtypes = [vp]
    L.gpv_destroy.restype = None
    L.gpv_set_revcond.argtypes = [vp, vp, i32]
    L.gpv_set_revcond.restype = i32
    L.gpv_u_nzentries.argtypes = [vp, cp, vp, i32, vp, vp, i64, vp, vp, C.POINTER(i64), C.POINTER(i64)]
    L.gpv_u_nzentries.restype = i32
    L.gpv_packed_len.argtypes = [vp]
    L.gpv_packed_len.restype = i64
    L.gpv_nuggets_read.argtypes = [vp]
    L.gpv_nuggets_read.restype = i64
    L.gpv_u_values_packed.argtypes = [vp, cp, vp, i32, vp, vp, i64, i32, vp, C.POINTER(i64), C.POINTER(i64)]
    L.gpv_u_values_packed.restype = i32
    L.gpv_csc_dims.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    L.gpv_csc_dims.restype = i32
    L.gpv_u_sparsity.argtypes = [vp, vp, vp]
    L.gpv_u_sparsity.restype = i32
    L.gpv_u_csc_pattern.argtypes = [vp, vp, vp]
    L.gpv_u_csc_pattern.restype = i32
    L.gpv_u_values_csc.argtypes = [vp, cp, vp, i32, vp, vp, i64, vp, C.POINTER(i64), C.POINTER(i64)]
    L.gpv_u_values_csc.restype = i32
    L.gpv_multi_csc_dims.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    L.gpv_multi_csc_dims.restype = i32
    L.gpv_multi_u_csc_pattern.argtypes = [vp, vp, vp]
    L.gpv_multi_u_csc_pattern.restype = i32
    L.gpv_multi_u_values_csc.argtypes = [vp, cp, vp, i32, vp, vp, i64, vp, C.POINTER(i64), C.POINTER(i64)]
    L.gpv_multi_u_values_csc.restype = i32
    L.gpv_whichCondOnLatent.argtypes = [vp, i64, i32, i64, vp]
    L.gpv_whichCondOnLatent.restype = i32
    L.gpv_ic0.argtypes = [i64, vp, vp, i64, vp]
    L.gpv_ic0.restype = i32
    L.gpv_createUcppM.argtypes = [i64, vp, vp, i64, vp]
    L.gpv_createUcppM.restype = i32
    L.gpv_createUcpp.argtypes = [i64, i32, vp, vp, i64, vp, vp, vp, i32]
    L.gpv_createUcpp.restype = i32
    L.gpv_u_nzentries_mat.argtypes = [vp, vp, vp, i64, vp, vp, C.POINTER(i64), C.POINTER(i64)]
    L.gpv_u_nzentries_mat.restype = i32
    L.gpv_u_values_packed_mat.argtypes = [vp, vp, vp, i64, i32, vp, C.POINTER(i64), C.POINTER(i64)]
    L.gpv_u_values_packed_mat.restype = i32
    L.gpv_set_scalar_nugget.argtypes = [vp, C.c_double]
    L.gpv_set_scalar_nugget.restype = i32
    L.gpv_loglik_numerator.argtypes = [vp, cp, vp, i32, vp, vp, vp, i64, i64, i32, vp]
    L.gpv_loglik_numerator.restype = i32
    L.gpv_loglik_z.argtypes = [vp, cp, vp, i32, vp, vp, vp, i64, i32, vp]
    L.gpv_loglik_z.restype = i32
    L.gpv_u_dev.argtypes = [vp, cp, vp, i32, vp, vp, i32, vp, i64, vp, vp]
    L.gpv_u_dev.restype = i32
    L.gpv_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.gpv_last_kernel_ms.restype = i32
    L.gpv_last_kernel_name.argtypes = [vp]
    L.gpv_last_kernel_name.restype = cp
    L.gpv_launch_count.restype = i64
    L.gpv_U_NZentries.argtypes = [i32, i64, i64, i32, i32, vp, vp, vp, i32, vp, vp, cp, vp, i32, vp, vp,
                                  C.POINTER(i64), C.POINTER(i64), i32]
    L.gpv_U_NZentries.restype = i32
    L.gpv_MaternFun.argtypes = [vp, i64, vp, vp, i32]
    L.gpv_MaternFun.restype = i32
    L.gpv_EsqeFun.argtypes = [vp, i64, vp, vp, i32]
    L.gpv_EsqeFun.restype = i32
    L.gpv_measure_fp64_peak.argtypes = [i32, C.POINTER(dbl)]
    L.gpv_measure_fp64_peak.restype = i32
    L.gpv_measure_copy_bw.argtypes = [i32, C.POINTER(dbl)]
    L.gpv_measure_copy_bw.restype = i32
    L.gpv_harness_ordered_nn.argtypes = [i64, i32, i32, vp, i64, i64, vp, i32]
    L.gpv_harness_ordered_nn.restype = i32
    L.gpv_multi_create.argtypes = [C.POINTER(vp), i64, i32, i32, vp, vp, vp, i32, vp, vp, i32]
    L.gpv_multi_create.restype = i32
    L.gpv_multi_destroy.argtypes = [vp]
    L.gpv_multi_destroy.restype = None
    L.gpv_multi_num_devices.argtypes = [vp]
    L.gpv_multi_num_devices.restype = i32
    L.gpv_multi_packed_len.argtypes = [vp]
    L.gpv_multi_packed_len.restype = i64
    L.gpv_multi_row_cuts.argtypes = [vp, vp]
    L.gpv_multi_row_cuts.restype = None
    L.gpv_multi_set_revcond.argtypes = [vp, vp, i32]
    L.gpv_multi_set_revcond.restype = i32
    L.gpv_multi_u_values_packed.argtypes = [vp, cp, vp, i32, vp, vp, i64, i32, vp, C.POINTER(i64), C.POINTER(i64)]
    L.gpv_multi_u_values_packed.restype = i32
    L.gpv_multi_loglik_numerator.argtypes = [vp, cp, vp, i32, vp, vp, vp, i64, i64, vp]
    L.gpv_multi_loglik_numerator.restype = i32
    L.gpv_multi_loglik_z.argtypes = [vp, cp, vp, i32, vp, vp, vp, i64, vp]
    L.gpv_multi_loglik_z.restype = i32
    L.gpv_dist_unique_id.argtypes = [vp]
    L.gpv_dist_unique_id.restype = i32
    L.gpv_dist_init.argtypes = [vp, vp, i32, i32]
    L.gpv_dist_init.restype = i32
    L.gpv_dist_finalize.argtypes = [vp]
    L.gpv_dist_finalize.restype = i32
    L.gpv_loglik_z_dist.argtypes = [vp, cp, vp, i32, vp, vp, vp, vp, vp, vp]
    L.gpv_loglik_z_dist.restype = i32
    L.gpv_set_last_error.argtypes = [cp]
    L.gpv_set_last_error.restype = None
    # host-only self-test hooks of the general-nu machinery (single points; not a compute path)
    L.gpv_selftest_matern_general_host.argtypes = [dbl, dbl, dbl]
    L.gpv_selftest_matern_general_host.restype = dbl
    L.gpv_selftest_table_eval_host.argtypes = [dbl, dbl, dbl, dbl, dbl]
    L.gpv_selftest_table_eval_host.restype = dbl
    return L


lib = _load()


def check(status):
    if status != GPV_OK:
        raise GpvError(status, lib.gpv_last_error().decode(errors="replace"))

"""ctypes binding of libtbslas_b200.so (the C ABI in include/tbslas_b200.h).

The library is the product; this module only loads it and declares the signatures.
There is no CPU fallback: if the shared object is missing the import fails loudly, and
every compute entry point returns TBSLAS_ERR_CUDA when no sm_100 device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtbslas_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_COMM, ERR_UNSUPPORTED, ERR_NOMEM = range(6)
FREESPACE, PERIODIC = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
FIELD_STEADY, FIELD_SET4, FIELD_EXTRAP = 0, 1, 2
MAX_CHEB_DEG = 19

# every symbol include/tbslas_b200.h declares (tests check the export table against it)
SYMBOLS = [
    "tbslas_b200_init", "tbslas_b200_finalize", "tbslas_b200_set_stream",
    "tbslas_b200_synchronize", "tbslas_b200_set_time_combine", "tbslas_b200_cubic_time_weights", "tbslas_b200_set_tensor_grid", "tbslas_b200_last_grid_exceptions", "tbslas_b200_last_error", "tbslas_b200_version",
    "tbslas_b200_comm_unique_id", "tbslas_b200_comm_init", "tbslas_b200_comm_rank",
    "tbslas_b200_comm_last_exchange", "tbslas_b200_comm_set_exchange", "tbslas_b200_comm_set_mailbox",
    "tbslas_b200_comm_exchange_mode", "tbslas_b200_tree_update_coeff_async", "tbslas_b200_set_host_chunks", "tbslas_b200_set_virtual_arrival_points",
    "tbslas_b200_tree_create", "tbslas_b200_tree_create_replicated", "tbslas_b200_tree_update_coeff", "tbslas_b200_tree_get_coeff", "tbslas_b200_tree_destroy",
    "tbslas_b200_tree_info", "tbslas_b200_eval", "tbslas_b200_eval_set4",
    "tbslas_b200_eval_extrap", "tbslas_b200_eval_field", "tbslas_b200_traj_rk2",
    "tbslas_b200_semilag_rk2", "tbslas_b200_semilag_insitu", "tbslas_b200_semilag_insitu_dep", "tbslas_b200_set_pt2coeff", "tbslas_b200_has_pt2coeff",
    "tbslas_b200_tree_set_grid_values", "tbslas_b200_semilag_insitu_update", "tbslas_b200_cubic_eval", "tbslas_b200_grid_create", "tbslas_b200_grid_update", "tbslas_b200_grid_eval",
    "tbslas_b200_grid_destroy", "tbslas_b200_collect_grid_points",
    "tbslas_b200_new_nodes", "tbslas_b200_point_key", "tbslas_b200_owner_of_key",
    "tbslas_b200_partition_leaves", "tbslas_b200_partition_leaves_weighted",
    "tbslas_b200_tree_reshard", "tbslas_b200_tree_global_range", "tbslas_b200_tree_last_point_counts", "tbslas_b200_tree_tail_norm", "tbslas_b200_profile_enable", "tbslas_b200_profile_reset",
    "tbslas_b200_profile_num_stages", "tbslas_b200_profile_stage_name", "tbslas_b200_profile_reference_tag",
    "tbslas_b200_profile_get", "tbslas_b200_kernel_launches", "tbslas_b200_fp64_peak",
]


class Field(C.Structure):
    """struct tbslas_field"""
    _fields_ = [("kind", C.c_int), ("tree", C.c_void_p * 4), ("times", C.c_double * 4)]


class TbslasError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("TBSLAS_B200_LIB") or LIB_PATH  # A/B builds of the same library (build.py --variant)
    if not os.path.exists(path):
        raise TbslasError(
            "%s not found: build it with `python -m tbslas_b200.build` (there is no CPU "
            "fallback)" % path)
    L = C.CDLL(path)
    vp, dp, i32p = C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses (host or device)
    sz = C.c_size_t
    L.tbslas_b200_init.argtypes = [C.c_int, C.POINTER(vp)]
    L.tbslas_b200_finalize.argtypes = [vp]
    L.tbslas_b200_set_stream.argtypes = [vp, vp]
    L.tbslas_b200_synchronize.argtypes = [vp]
    L.tbslas_b200_set_time_combine.argtypes = [vp, C.c_int]
    L.tbslas_b200_cubic_time_weights.argtypes = [C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double)]
    L.tbslas_b200_set_tensor_grid.argtypes = [vp, C.c_int]
    L.tbslas_b200_last_grid_exceptions.argtypes = [vp, C.POINTER(sz)]
    L.tbslas_b200_last_error.argtypes = [vp]
    L.tbslas_b200_last_error.restype = C.c_char_p
    L.tbslas_b200_version.restype = C.c_char_p
    L.tbslas_b200_comm_unique_id.argtypes = [vp]
    L.tbslas_b200_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.tbslas_b200_comm_rank.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.tbslas_b200_comm_last_exchange.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
    L.tbslas_b200_comm_set_exchange.argtypes = [vp, C.c_int]
    L.tbslas_b200_comm_set_mailbox.argtypes = [vp, sz]
    L.tbslas_b200_comm_exchange_mode.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(sz)]
    L.tbslas_b200_set_host_chunks.argtypes = [vp, C.c_int]
    L.tbslas_b200_set_virtual_arrival_points.argtypes = [vp, C.c_int]
    L.tbslas_b200_tree_update_coeff_async.argtypes = [vp, dp, C.c_int]
    L.tbslas_b200_tree_create.argtypes = [vp, C.c_int, C.c_int, sz, dp, vp, dp, C.c_int,
                                          C.POINTER(vp)]
    L.tbslas_b200_tree_create_replicated.argtypes = L.tbslas_b200_tree_create.argtypes
    L.tbslas_b200_tree_update_coeff.argtypes = [vp, dp, C.c_int]
    L.tbslas_b200_tree_get_coeff.argtypes = [vp, dp, C.c_int]
    L.tbslas_b200_tree_destroy.argtypes = [vp]
    L.tbslas_b200_tree_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(sz)]
    L.tbslas_b200_eval.argtypes = [vp, C.c_int, dp, sz, dp, i32p, C.c_int]
    L.tbslas_b200_eval_set4.argtypes = [C.POINTER(vp), C.POINTER(C.c_double), C.c_double,
                                        C.c_int, dp, sz, dp, C.c_int]
    L.tbslas_b200_eval_extrap.argtypes = [vp, vp, C.c_int, dp, sz, dp, C.c_int]
    L.tbslas_b200_eval_field.argtypes = [C.POINTER(Field), C.c_double, C.c_int, dp, sz, dp,
                                         C.c_int]
    L.tbslas_b200_traj_rk2.argtypes = [C.POINTER(Field), C.POINTER(Field), C.c_int, dp, sz,
                                       C.c_double, C.c_double, C.c_int, dp, C.c_int]
    L.tbslas_b200_semilag_rk2.argtypes = [C.POINTER(Field), C.POINTER(Field), vp, C.c_int, dp,
                                          sz, C.c_int, C.c_double, C.c_int, dp, dp, C.c_int]
    L.tbslas_b200_semilag_insitu.argtypes = [C.POINTER(Field), C.POINTER(Field), vp, C.c_int,
                                             C.c_int, C.c_double, C.c_int, dp, C.c_int]
    L.tbslas_b200_semilag_insitu_dep.argtypes = [C.POINTER(Field), C.POINTER(Field), vp, C.c_int,
                                                 C.c_int, C.c_double, C.c_int, dp, dp, C.c_int]
    L.tbslas_b200_set_pt2coeff.argtypes = [vp, C.c_int, dp]
    L.tbslas_b200_has_pt2coeff.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
    L.tbslas_b200_tree_set_grid_values.argtypes = [vp, dp, C.c_int, C.c_int]
    L.tbslas_b200_semilag_insitu_update.argtypes = [C.POINTER(Field), C.POINTER(Field), vp, C.c_int,
                                                    C.c_int, C.c_double, C.c_int]
    L.tbslas_b200_cubic_eval.argtypes = [vp, dp, C.c_int, C.c_int, dp, sz, dp, C.c_int]
    L.tbslas_b200_grid_create.argtypes = [vp, dp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.tbslas_b200_grid_update.argtypes = [vp, dp, C.c_int]
    L.tbslas_b200_grid_eval.argtypes = [vp, dp, sz, dp, C.c_int]
    L.tbslas_b200_grid_destroy.argtypes = [vp]
    L.tbslas_b200_collect_grid_points.argtypes = [vp, dp, C.c_int]
    L.tbslas_b200_new_nodes.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.tbslas_b200_point_key.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
    L.tbslas_b200_point_key.restype = C.c_uint64
    L.tbslas_b200_owner_of_key.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
    L.tbslas_b200_partition_leaves.argtypes = [sz, C.c_int, C.POINTER(sz)]
    L.tbslas_b200_partition_leaves_weighted.argtypes = [sz, C.POINTER(C.c_double), C.c_int, C.POINTER(sz)]
    L.tbslas_b200_tree_reshard.argtypes = [vp, C.POINTER(sz)]
    L.tbslas_b200_tree_global_range.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
    L.tbslas_b200_tree_last_point_counts.argtypes = [vp, vp, C.c_int]
    L.tbslas_b200_tree_tail_norm.argtypes = [vp, dp, C.c_int]
    L.tbslas_b200_profile_enable.argtypes = [vp, C.c_int]
    L.tbslas_b200_profile_reset.argtypes = [vp]
    L.tbslas_b200_profile_stage_name.argtypes = [C.c_int]
    L.tbslas_b200_profile_stage_name.restype = C.c_char_p
    L.tbslas_b200_profile_reference_tag.argtypes = [C.c_int]
    L.tbslas_b200_profile_reference_tag.restype = C.c_char_p
    L.tbslas_b200_profile_get.argtypes = [vp, C.c_int, C.POINTER(C.c_double),
                                          C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
    L.tbslas_b200_kernel_launches.argtypes = [vp]
    L.tbslas_b200_kernel_launches.restype = C.c_longlong
    L.tbslas_b200_fp64_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    _lib = L
    return L

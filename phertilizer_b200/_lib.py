"""ctypes binding of the C-ABI library ``libphertilizer_b200.so`` (declared in include/phertilizer_b200.h).

There is no CPU fallback: if the shared library has not been built (``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C phertilizer_b200/csrc``) importing the product path raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libphertilizer_b200.so")

PH_OK = 0
ABI_VERSION = 9


class PhError(RuntimeError):
    pass


class PhInfo(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int32), ("n_snvs", C.c_int32), ("n_codes", C.c_int32),
        ("nnz", C.c_int64),
        ("dense_codes_by_snv", C.c_int32), ("dense_codes_by_cell", C.c_int32),
        ("device_bytes", C.c_int64),
        ("device", C.c_int32), ("sm_count", C.c_int32),
        ("group_by_snv", C.c_int32), ("group_by_cell", C.c_int32),
        ("labels_in_smem_by_snv", C.c_int32), ("labels_in_smem_by_cell", C.c_int32),
        ("chunks_by_snv", C.c_int32), ("chunks_by_cell", C.c_int32),
        ("stream_bytes_by_snv", C.c_int64), ("stream_bytes_by_cell", C.c_int64),
        ("stream4_bytes_by_snv", C.c_int64), ("stream4_bytes_by_cell", C.c_int64),
    ]


# every symbol include/phertilizer_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "ph_last_error", "ph_abi_version", "ph_kernel_launches", "ph_device_count", "ph_create", "ph_destroy", "ph_get_info", "ph_set_group",
    "ph_snv_table", "ph_assign_linear", "ph_assign_branching", "ph_cell_table", "ph_block_sums", "ph_tree_loglik",
    "ph_line_reduce", "ph_node_reduce", "ph_snv_node_table", "ph_cell_node_table", "ph_set_embedding", "ph_gauss_node_ll", "ph_pairwise_dist", "ph_pairwise_dist_gram",
    "ph_snv_partial", "ph_finish_linear", "ph_finish_branching",
    "ph_peer_handle_bytes", "ph_peer_create", "ph_peer_export", "ph_peer_connect", "ph_peer_assign_linear",
    "ph_peer_assign_branching", "ph_peer_assign_linear_snv", "ph_peer_assign_branching_snv", "ph_peer_timed_out",
    "ph_peer_debug", "ph_peer_destroy",
    "ph_affinity_create", "ph_affinity_rows", "ph_affinity_nearest", "ph_affinity_build", "ph_affinity_matmat", "ph_affinity_matrix",
    "ph_affinity_destroy", "ph_affinity_create_rows", "ph_affinity_stats_rows", "ph_affinity_build_rows", "ph_affinity_laplacian_rows",
    "ph_affinity_matmat_rows", "ph_affinity_bv_set", "ph_affinity_bv_get", "ph_affinity_bv_apply", "ph_affinity_bv_gram",
    "ph_affinity_bv_combine",
]

_lib = None
_vp = C.c_void_p


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PhError(
            f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built and phertilizer_b200 has no CPU "
            "fallback. Run `make -C phertilizer_b200/csrc` (or __graft_entry__.build()).")
    L = C.CDLL(LIB_PATH)
    L.ph_last_error.restype = C.c_char_p
    L.ph_abi_version.restype = C.c_int
    L.ph_kernel_launches.restype = C.c_int64
    L.ph_create.argtypes = [C.c_int32, C.c_int32, C.c_int64, _vp, _vp, _vp, C.c_int, C.c_int32, _vp, _vp, _vp, C.c_int,
                            C.POINTER(_vp)]
    L.ph_device_count.argtypes = [C.POINTER(C.c_int)]
    L.ph_destroy.argtypes = [_vp]
    L.ph_destroy.restype = None
    L.ph_get_info.argtypes = [_vp, C.POINTER(PhInfo)]
    L.ph_set_group.argtypes = [_vp, C.c_int, C.c_int]
    L.ph_snv_table.argtypes = [_vp, _vp, C.c_int32, _vp, C.c_int32, _vp, _vp, _vp, _vp, C.c_int, _vp]
    L.ph_cell_table.argtypes = L.ph_snv_table.argtypes
    L.ph_assign_linear.argtypes = [_vp, _vp, C.c_double, _vp, C.c_int32, _vp, _vp, C.c_int, _vp]
    L.ph_assign_branching.argtypes = [_vp, _vp, _vp, C.c_int32, _vp, _vp, C.c_int, _vp]
    L.ph_block_sums.argtypes = [_vp, _vp, C.c_int32, _vp, C.c_int32, _vp, _vp, _vp, _vp, C.c_int, _vp]
    L.ph_line_reduce.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp]
    L.ph_node_reduce.argtypes = [_vp, _vp, _vp, C.c_int32, C.c_int32, _vp]
    L.ph_tree_loglik.argtypes = [_vp, _vp, _vp, _vp, C.c_int32, _vp, _vp, C.c_int, _vp]
    L.ph_snv_node_table.argtypes = [_vp, _vp, _vp, C.c_int32, _vp, C.c_int32, _vp, C.c_int, _vp]
    L.ph_cell_node_table.argtypes = L.ph_snv_node_table.argtypes
    L.ph_snv_partial.argtypes = [_vp, _vp, C.c_int32, _vp, C.c_int32, _vp, C.c_int, _vp]
    L.ph_finish_linear.argtypes = [_vp, C.c_int32, C.c_double, _vp, _vp, _vp]
    L.ph_finish_branching.argtypes = [_vp, C.c_int32, _vp, _vp, _vp]
    L.ph_peer_create.argtypes = [_vp, C.c_int, C.c_int, C.c_int32, C.POINTER(_vp)]
    L.ph_peer_assign_linear_snv.argtypes = [_vp, _vp, C.c_double, C.c_int32, _vp, _vp, _vp]
    L.ph_peer_assign_branching_snv.argtypes = [_vp, _vp, C.c_int32, _vp, _vp, _vp]
    L.ph_peer_export.argtypes = [_vp, _vp]
    L.ph_peer_connect.argtypes = [_vp, _vp]
    L.ph_peer_assign_linear.argtypes = [_vp, _vp, C.c_double, _vp, C.c_int32, _vp, _vp, _vp]
    L.ph_peer_assign_branching.argtypes = [_vp, _vp, _vp, C.c_int32, _vp, _vp, _vp]
    L.ph_peer_timed_out.argtypes = [_vp, _vp]
    L.ph_peer_debug.argtypes = [_vp, _vp]
    L.ph_peer_destroy.argtypes = [_vp]
    L.ph_peer_destroy.restype = None
    L.ph_affinity_create.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int, C.POINTER(_vp)]
    L.ph_affinity_rows.argtypes = [_vp, _vp, C.c_int32, _vp]
    L.ph_affinity_nearest.argtypes = [_vp, _vp, C.c_int32, _vp, _vp, _vp]
    L.ph_affinity_build.argtypes = [_vp, _vp, C.c_int32, _vp, C.c_int32, C.c_int, C.c_double, C.c_int, _vp, _vp]
    L.ph_affinity_matmat.argtypes = [_vp, _vp, C.c_int32, _vp]
    L.ph_affinity_matrix.argtypes = [_vp, _vp]
    L.ph_affinity_create_rows.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int, C.c_int32, C.POINTER(_vp)]
    L.ph_affinity_stats_rows.argtypes = [_vp, _vp, C.c_int32, _vp, C.c_int32, C.c_int, C.c_int32, C.c_int32, _vp]
    L.ph_affinity_build_rows.argtypes = [_vp, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int, _vp, C.POINTER(C.c_int)]
    L.ph_affinity_laplacian_rows.argtypes = [_vp, _vp]
    L.ph_affinity_matmat_rows.argtypes = [_vp, _vp, C.c_int32, _vp]
    L.ph_affinity_bv_set.argtypes = [_vp, C.c_int, _vp]
    L.ph_affinity_bv_get.argtypes = [_vp, C.c_int, _vp]
    L.ph_affinity_bv_apply.argtypes = [_vp, C.c_int, C.c_int, _vp]
    L.ph_affinity_bv_gram.argtypes = [_vp, _vp, C.c_int32, _vp, C.c_int32, _vp]
    L.ph_affinity_bv_combine.argtypes = [_vp, C.c_int, _vp, _vp, C.c_int32]
    L.ph_affinity_destroy.argtypes = [_vp]
    L.ph_affinity_destroy.restype = None
    L.ph_set_embedding.argtypes = [_vp, _vp, C.c_int32]
    L.ph_gauss_node_ll.argtypes = [_vp, _vp, C.c_int32, _vp, C.c_int32, _vp, _vp, C.c_int, _vp]
    L.ph_pairwise_dist.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, C.c_int, C.c_int, _vp]
    L.ph_pairwise_dist_gram.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, C.c_int, _vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("ph_last_error", "ph_kernel_launches", "ph_destroy", "ph_peer_destroy", "ph_affinity_destroy"):
            fn.restype = C.c_int
    if L.ph_abi_version() != ABI_VERSION:
        raise PhError(f"ABI version mismatch: library {L.ph_abi_version()} != binding {ABI_VERSION}")
    _lib = L
    return L


def check(rc: int):
    if rc != PH_OK:
        raise PhError(f"phertilizer_b200 error {rc}: {lib().ph_last_error().decode()}")


def kernel_launches() -> int:
    return int(lib().ph_kernel_launches())

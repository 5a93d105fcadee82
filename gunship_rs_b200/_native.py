"""Loader + ctypes prototypes of libgunship_collision.so (include/gunship_collision.h).

There is no CPU fallback: if the CUDA library is missing this module raises at import of the
symbols, and gsc_create fails loudly when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GSC_LIB: development knob -- load another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("GSC_LIB") or os.path.join(_HERE, "libgunship_collision.so")

GSC_OK, GSC_ERR_INVALID, GSC_ERR_CUDA, GSC_ERR_NCCL, GSC_ERR_CAPACITY, GSC_ERR_RANGE, GSC_ERR_KIND = range(7)
GSC_SPHERE, GSC_BOX, GSC_MESH = 0, 1, 2
GSC_FLAG_TIMING, GSC_FLAG_GRAPH, GSC_FLAG_CONTACTS = 1, 2, 4
GSC_NUM_STAGES = 6
GSC_HALO_RECORD_BYTES = 96
GSC_MAX_PEERS = 16
GSC_IPC_EXPORT_BYTES = 256
STAGE_NAMES = ("bvh_update", "cell_keys", "sort", "cell_table", "pairs", "narrow")
(GSC_DBG_AABB, GSC_DBG_SORTED_KEYS, GSC_DBG_SORTED_IDX, GSC_DBG_CELL_START, GSC_DBG_OBB, GSC_DBG_ENTITY, GSC_DBG_KIND,
 GSC_DBG_OFFSET, GSC_DBG_SIZE) = range(9)
GSC_NO_PARENT = 0xFFFFFFFF

ERROR_NAMES = {GSC_ERR_INVALID: "GSC_ERR_INVALID", GSC_ERR_CUDA: "GSC_ERR_CUDA", GSC_ERR_NCCL: "GSC_ERR_NCCL",
               GSC_ERR_CAPACITY: "GSC_ERR_CAPACITY", GSC_ERR_RANGE: "GSC_ERR_RANGE", GSC_ERR_KIND: "GSC_ERR_KIND"}


class GscError(RuntimeError):
    """A non-zero status from the C ABI (the reference panics in the same situations)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


class GscConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_colliders", C.c_uint32), ("max_pairs", C.c_uint64),
                ("max_candidates", C.c_uint64), ("max_cells", C.c_uint64), ("flags", C.c_uint32),
                ("reserved", C.c_uint32)]


class GscStepOut(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("n_sphere_sphere", C.c_uint64), ("n_sphere_obb", C.c_uint64), ("n_obb_obb", C.c_uint64),
                ("n_cells", C.c_uint64), ("cell_size", C.c_float), ("n_wide", C.c_uint32),
                ("grid_origin", C.c_int32 * 3), ("grid_dims", C.c_int32 * 3), ("sort_passes", C.c_uint32),
                ("status_flags", C.c_uint32), ("stage_ms", C.c_float * GSC_NUM_STAGES), ("step_ms", C.c_float),
                ("d_pairs", C.c_void_p), ("d_contacts", C.c_void_p), ("n_ghost", C.c_uint32), ("reserved", C.c_uint32)]


class GscAdjacency(C.Structure):
    _fields_ = [("n_entities", C.c_uint64), ("n_entries", C.c_uint64), ("d_entity", C.c_void_p),
                ("d_offset", C.c_void_p), ("d_other", C.c_void_p)]


class GscBounds(C.Structure):
    _fields_ = [("longest_axis", C.c_float), ("world_min", C.c_float * 3), ("world_max", C.c_float * 3),
                ("status_flags", C.c_uint32), ("home_x_max", C.c_float)]


# every symbol include/gunship_collision.h declares
EXPORTS = (
    "gsc_create", "gsc_destroy", "gsc_last_error", "gsc_abi_version", "gsc_upload_colliders",
    "gsc_update_transforms", "gsc_update_transforms_packed", "gsc_bind_device_transforms", "gsc_step", "gsc_download_pairs", "gsc_debug_download",
    "gsc_test_sphere_sphere", "gsc_test_sphere_obb", "gsc_test_obb_obb", "gsc_test_aabb_aabb", "gsc_fnv1a_gridcell",
    "gsc_sort_pairs_u32", "gsc_phase_bounds", "gsc_set_global_bounds", "gsc_local_x_range", "gsc_halo_select",
    "gsc_halo_append", "gsc_phase_finish", "gsc_set_x_window", "gsc_ipc_export", "gsc_ipc_open_peer", "gsc_halo_push",
    "gsc_halo_sync", "gsc_halo_commit", "gsc_download_contacts", "gsc_contact_sphere_sphere", "gsc_contact_sphere_obb",
    "gsc_contact_obb_obb", "gsc_build_adjacency", "gsc_download_adjacency", "gsc_step_begin", "gsc_step_end",
    "gsc_append_colliders", "gsc_remove_colliders", "gsc_collider_count", "gsc_hierarchy_upload", "gsc_hierarchy_update",
    "gsc_hierarchy_download", "gsc_open_peer_local", "gsc_slab_init", "gsc_slab_step", "gsc_slab_step_begin",
    "gsc_group_create", "gsc_group_destroy", "gsc_group_last_error", "gsc_group_upload_colliders",
    "gsc_group_update_transforms", "gsc_group_step", "gsc_group_download_pairs", "gsc_group_rebalance", "gsc_group_reorder", "gsc_group_member",
    "gsc_map_transform_staging", "gsc_commit_transform_staging", "gsc_box_slots", "gsc_reorder_colliders", "gsc_set_flags",
)

_lib = None


def load():
    """dlopen the CUDA library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32p, u8p, f32p, u64p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    L.gsc_abi_version.restype = C.c_int
    L.gsc_last_error.restype = C.c_char_p
    L.gsc_last_error.argtypes = [vp]
    L.gsc_create.argtypes = [C.POINTER(GscConfig), C.POINTER(vp)]
    L.gsc_destroy.argtypes = [vp]
    L.gsc_destroy.restype = None
    L.gsc_upload_colliders.argtypes = [vp, C.c_uint32, u32p, u8p, f32p, f32p, u32p]
    L.gsc_update_transforms.argtypes = [vp, C.c_uint32, f32p, f32p, f32p]
    L.gsc_update_transforms_packed.argtypes = [vp, C.c_uint32, f32p, C.c_uint32, f32p, f32p]
    L.gsc_bind_device_transforms.argtypes = [vp, C.c_uint32, vp, vp, vp]
    L.gsc_step.argtypes = [vp, C.POINTER(GscStepOut)]
    L.gsc_step_begin.argtypes = [vp]
    L.gsc_step_end.argtypes = [vp, C.POINTER(GscStepOut)]
    L.gsc_append_colliders.argtypes = [vp, C.c_uint32, u32p, u8p, f32p, f32p]
    L.gsc_remove_colliders.argtypes = [vp, C.c_uint32, u32p]
    L.gsc_collider_count.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.gsc_hierarchy_upload.argtypes = [vp, C.c_uint32, C.c_uint32, u32p, u32p, C.c_uint32, u32p]
    L.gsc_hierarchy_update.argtypes = [vp, f32p, f32p, f32p]
    L.gsc_hierarchy_download.argtypes = [vp, f32p, f32p, f32p, f32p]
    L.gsc_download_pairs.argtypes = [vp, u32p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    L.gsc_debug_download.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.gsc_build_adjacency.argtypes = [vp, C.POINTER(GscAdjacency)]
    L.gsc_download_adjacency.argtypes = [vp, u32p, u32p, u32p, C.c_uint64, C.c_uint64]
    L.gsc_download_contacts.argtypes = [vp, f32p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    L.gsc_contact_sphere_sphere.argtypes = [C.c_int, C.c_size_t, f32p, f32p, f32p]
    L.gsc_contact_sphere_obb.argtypes = [C.c_int, C.c_size_t, f32p, f32p, u8p, f32p]
    L.gsc_contact_obb_obb.argtypes = [C.c_int, C.c_size_t, f32p, f32p, f32p]
    L.gsc_test_sphere_sphere.argtypes = [C.c_int, C.c_size_t, f32p, f32p, u8p]
    L.gsc_test_sphere_obb.argtypes = [C.c_int, C.c_size_t, f32p, f32p, u8p]
    L.gsc_test_obb_obb.argtypes = [C.c_int, C.c_size_t, f32p, f32p, u8p]
    L.gsc_test_aabb_aabb.argtypes = [C.c_int, C.c_size_t, f32p, f32p, u8p]
    L.gsc_fnv1a_gridcell.argtypes = [C.c_int, C.c_size_t, vp, u64p]
    L.gsc_sort_pairs_u32.argtypes = [C.c_int, C.c_size_t, u32p, u32p, C.c_int]
    L.gsc_phase_bounds.argtypes = [vp, C.POINTER(GscBounds)]
    L.gsc_set_global_bounds.argtypes = [vp, C.POINTER(GscBounds)]
    L.gsc_local_x_range.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.gsc_halo_select.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_uint32, C.POINTER(C.c_uint32)]
    L.gsc_halo_append.argtypes = [vp, C.c_uint32, vp]
    L.gsc_phase_finish.argtypes = [vp, C.POINTER(GscStepOut)]
    L.gsc_set_x_window.argtypes = [vp, C.c_int32, C.c_int32]
    L.gsc_ipc_export.argtypes = [vp, vp, C.c_size_t]
    L.gsc_ipc_open_peer.argtypes = [vp, C.c_uint32, vp, C.c_size_t]
    L.gsc_halo_push.argtypes = [vp, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32]
    L.gsc_halo_sync.argtypes = [vp]
    L.gsc_halo_commit.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.gsc_open_peer_local.argtypes = [vp, C.c_uint32, vp]
    L.gsc_slab_init.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.gsc_slab_step.argtypes = [vp, C.POINTER(GscStepOut)]
    L.gsc_slab_step_begin.argtypes = [vp]
    L.gsc_group_create.argtypes = [C.c_uint32, C.POINTER(C.c_int32), C.POINTER(GscConfig), C.POINTER(vp)]
    L.gsc_group_destroy.argtypes = [vp]
    L.gsc_group_destroy.restype = None
    L.gsc_group_last_error.argtypes = [vp]
    L.gsc_group_last_error.restype = C.c_char_p
    L.gsc_group_upload_colliders.argtypes = [vp, C.c_uint32, u32p, u8p, f32p, f32p, f32p]
    L.gsc_group_update_transforms.argtypes = [vp, C.c_uint32, f32p, f32p, f32p]
    L.gsc_group_step.argtypes = [vp, C.POINTER(GscStepOut)]
    L.gsc_group_download_pairs.argtypes = [vp, u32p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    L.gsc_group_rebalance.argtypes = [vp, C.c_uint32, f32p]
    L.gsc_group_reorder.argtypes = [vp]
    L.gsc_group_member.argtypes = [vp, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_uint32), u32p]
    L.gsc_map_transform_staging.argtypes = [vp, C.c_uint32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.gsc_commit_transform_staging.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
    L.gsc_box_slots.argtypes = [vp, u32p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.gsc_reorder_colliders.argtypes = [vp, u32p, C.c_uint32]
    L.gsc_set_flags.argtypes = [vp, C.c_uint32]
    for name in EXPORTS:
        fn = getattr(L, name)  # raises AttributeError if the .so does not export it
        if name not in ("gsc_destroy", "gsc_last_error", "gsc_group_destroy", "gsc_group_last_error"):
            fn.restype = C.c_int
    _lib = L
    return L

"""Thin object wrapper over the C ABI (one gsc_ctx per GPU).  Host arrays are numpy; nothing here
computes -- every call forwards to libgunship_collision.so and raises GscError on a non-zero status."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from ._native import GscError


def _ptr(a):
    return None if a is None else a.ctypes.data


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class CollisionContext:
    """gsc_ctx: device buffers + the per-frame step for up to `max_colliders` colliders on one GPU."""

    def __init__(self, max_colliders: int, device: int = 0, max_pairs: int = 0, max_candidates: int = 0,
                 max_cells: int = 0, timing: bool = False, contacts: bool = False, graph: bool = False):
        self._lib = N.load()
        cfg = N.GscConfig(device, max_colliders, max_pairs, max_candidates, max_cells,
                          (N.GSC_FLAG_TIMING if timing else 0) | (N.GSC_FLAG_CONTACTS if contacts else 0) |
                          (N.GSC_FLAG_GRAPH if graph else 0), 0)
        self.cap = max_colliders
        self._flags = cfg.flags
        h = C.c_void_p()
        rc = self._lib.gsc_create(C.byref(cfg), C.byref(h))
        if rc != N.GSC_OK:
            raise GscError(rc, self._lib.gsc_last_error(None).decode())
        self._h = h
        self.device = device
        self.n = 0
        self.last = None
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gsc_destroy(self._h)
            self._h = None

    def set_flags(self, timing: bool = False, graph: bool = False):
        """GSC_FLAG_TIMING / GSC_FLAG_GRAPH of the live context (per-stage times and graph replay exclude each other)"""
        keep = self._flags & N.GSC_FLAG_CONTACTS
        self._flags = keep | (N.GSC_FLAG_TIMING if timing else 0) | (N.GSC_FLAG_GRAPH if graph else 0)
        self._check(self._lib.gsc_set_flags(self._h, self._flags))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != N.GSC_OK:
            raise GscError(rc, self._lib.gsc_last_error(self._h).decode())

    # ColliderManager::assign over the dense array (mod.rs:221-223)
    def upload_colliders(self, entity, kind, offset, size, global_index=None):
        e, k = _c(entity, np.uint32), _c(kind, np.uint8)
        o, s, g = _c(offset, np.float32), _c(size, np.float32), _c(global_index, np.uint32)
        self.n = int(e.shape[0])
        self.n_box = int((k == N.GSC_BOX).sum())
        self._check(self._lib.gsc_upload_colliders(self._h, self.n, _ptr(e), _ptr(k), _ptr(o), _ptr(s), _ptr(g)))

    def update_transforms(self, pos, quat=None, scale=None):
        p, q, s = _c(pos, np.float32), _c(quat, np.float32), _c(scale, np.float32)
        self._keep = [p, q, s]  # stay alive until the step has consumed them
        self._check(self._lib.gsc_update_transforms(self._h, self.n, _ptr(p), _ptr(q), _ptr(s)))

    def update_transforms_packed(self, pos, box_quat=None, box_scale=None):
        """Rotation / scale for the boxes only, in collider order (gsc_update_transforms_packed)."""
        p, q, s = _c(pos, np.float32), _c(box_quat, np.float32), _c(box_scale, np.float32)
        self._keep = [p, q, s]
        n_box = self.n_box if q is None else int(q.shape[0])
        self._check(self._lib.gsc_update_transforms_packed(self._h, self.n, _ptr(p), n_box, _ptr(q), _ptr(s)))

    def update_transforms_packed_ptr(self, pos_ptr: int, n_box: int, box_quat_ptr: int = 0, box_scale_ptr: int = 0):
        self._check(self._lib.gsc_update_transforms_packed(self._h, self.n, pos_ptr or None, n_box, box_quat_ptr or None,
                                                           box_scale_ptr or None))

    def update_transforms_ptr(self, pos_ptr: int, quat_ptr: int = 0, scale_ptr: int = 0):
        """Host pointers (e.g. pinned torch tensors' data_ptr()); arrays must outlive the next step()."""
        self._check(self._lib.gsc_update_transforms(self._h, self.n, pos_ptr or None, quat_ptr or None, scale_ptr or None))

    def bind_device_transforms(self, d_pos: int, d_quat: int = 0, d_scale: int = 0):
        self._check(self._lib.gsc_bind_device_transforms(self._h, self.n, d_pos, d_quat or None, d_scale or None))

    # CollisionSystem::update (mod.rs:590-611)
    def step(self) -> N.GscStepOut:
        out = N.GscStepOut()
        rc = self._lib.gsc_step(self._h, C.byref(out))
        self.last = out
        self._check(rc)
        return out

    # the same step in two halves: the next frame's update_transforms* may be issued in between (gsc_step_begin)
    def step_begin(self):
        self._check(self._lib.gsc_step_begin(self._h))

    def step_end(self) -> N.GscStepOut:
        out = N.GscStepOut()
        rc = self._lib.gsc_step_end(self._h, C.byref(out))
        self.last = out
        self._check(rc)
        return out

    # lifecycle after the first upload: ColliderManager::assign / BoundingVolumeManager::destroy_immediate
    def append_colliders(self, entity, kind, offset, size):
        e, k = _c(entity, np.uint32), _c(kind, np.uint8)
        o, s = _c(offset, np.float32), _c(size, np.float32)
        self._check(self._lib.gsc_append_colliders(self._h, int(e.shape[0]), _ptr(e), _ptr(k), _ptr(o), _ptr(s)))
        self.n += int(e.shape[0])
        self.n_box += int((k == N.GSC_BOX).sum())

    def remove_colliders(self, index):
        """swap_remove the dense positions index[0], index[1], ... in that order (bounding_volume.rs:82-104)."""
        i = _c(index, np.uint32)
        self._check(self._lib.gsc_remove_colliders(self._h, int(i.shape[0]), _ptr(i)))
        self.n = self.collider_count()
        self.n_box = int((self.debug(N.GSC_DBG_KIND) == N.GSC_BOX).sum()) if self.n else 0

    def collider_count(self) -> int:
        n = C.c_uint32(0)
        self._check(self._lib.gsc_collider_count(self._h, C.byref(n)))
        return int(n.value)

    # TransformManager::update_transforms on the device (component/transform.rs:243-251, 588-608)
    def hierarchy_upload(self, row_begin, parent, collider_node):
        rb, p, cn = _c(row_begin, np.uint32), _c(parent, np.uint32), _c(collider_node, np.uint32)
        self.n_nodes = int(p.shape[0])
        self._check(self._lib.gsc_hierarchy_upload(self._h, self.n_nodes, int(rb.shape[0]) - 1, _ptr(rb), _ptr(p),
                                                   int(cn.shape[0]), _ptr(cn)))

    def hierarchy_update(self, local_pos=None, local_quat=None, local_scale=None):
        p, q, s = _c(local_pos, np.float32), _c(local_quat, np.float32), _c(local_scale, np.float32)
        self._check(self._lib.gsc_hierarchy_update(self._h, _ptr(p), _ptr(q), _ptr(s)))

    def hierarchy_download(self):
        """derived (pos[n,3], quat[n,4], scale[n,3], matrix[n,4,4]) of every node"""
        n = self.n_nodes
        pos, quat = np.zeros((n, 3), np.float32), np.zeros((n, 4), np.float32)
        scale, mat = np.zeros((n, 3), np.float32), np.zeros((n, 4, 4), np.float32)
        self._check(self._lib.gsc_hierarchy_download(self._h, _ptr(pos), _ptr(quat), _ptr(scale), _ptr(mat)))
        return pos, quat, scale, mat

    def pairs(self, sorted: bool = True) -> np.ndarray:
        """(n,2) uint32: (entity of later volume, entity of earlier volume) -- grid_collision.rs:119,497."""
        n = C.c_uint64(0)
        self._check(self._lib.gsc_download_pairs(self._h, None, 0, 0, C.byref(n)))
        out = np.zeros((int(n.value), 2), dtype=np.uint32)
        if n.value:
            self._check(self._lib.gsc_download_pairs(self._h, _ptr(out), n.value, 1 if sorted else 0, C.byref(n)))
        return out

    def contacts(self, sorted: bool = True) -> np.ndarray:
        """(n,4) float32: (normal.xyz, depth) per pair, same order as pairs(sorted) -- needs contacts=True."""
        n = C.c_uint64(0)
        self._check(self._lib.gsc_download_pairs(self._h, None, 0, 0, C.byref(n)))
        out = np.zeros((int(n.value), 4), dtype=np.float32)
        if n.value:
            self._check(self._lib.gsc_download_contacts(self._h, _ptr(out), n.value, 1 if sorted else 0, C.byref(n)))
        return out

    def adjacency(self):
        """CollisionCallbackManager::process_collisions' per-entity lists (mod.rs:678-708) as CSR:
        (entity[m], offset[m+1], other[2*n_pairs]); entity i collides with other[offset[i]:offset[i+1]]."""
        a = N.GscAdjacency()
        self._check(self._lib.gsc_build_adjacency(self._h, C.byref(a)))
        ent = np.zeros(int(a.n_entities), dtype=np.uint32)
        off = np.zeros(int(a.n_entities) + 1, dtype=np.uint32)
        oth = np.zeros(int(a.n_entries), dtype=np.uint32)
        self._check(self._lib.gsc_download_adjacency(self._h, _ptr(ent), _ptr(off), _ptr(oth), ent.shape[0], oth.shape[0]))
        return ent, off, oth

    def pairs_packed(self, sorted: bool = True) -> np.ndarray:
        p = self.pairs(sorted).astype(np.uint64)
        return (p[:, 0] << np.uint64(32)) | p[:, 1]

    def debug(self, what: int) -> np.ndarray:
        n = self.n
        if what == N.GSC_DBG_AABB:
            out = np.zeros((n, 6), dtype=np.float32)
        elif what == N.GSC_DBG_OBB:
            out = np.zeros((n, 16), dtype=np.float32)
        elif what == N.GSC_DBG_CELL_START:
            out = np.zeros(int(self.last.n_cells) + 1, dtype=np.uint32)
        elif what == N.GSC_DBG_KIND:
            out = np.zeros(n, dtype=np.uint8)
        elif what in (N.GSC_DBG_OFFSET, N.GSC_DBG_SIZE):
            out = np.zeros((n, 3), dtype=np.float32)
        else:
            out = np.zeros(n, dtype=np.uint32)
        self._check(self._lib.gsc_debug_download(self._h, what, _ptr(out), out.nbytes))
        return out

    # multi-GPU phases
    def phase_bounds(self) -> N.GscBounds:
        b = N.GscBounds()
        self._check(self._lib.gsc_phase_bounds(self._h, C.byref(b)))
        return b

    def set_global_bounds(self, b: N.GscBounds):
        self._check(self._lib.gsc_set_global_bounds(self._h, C.byref(b)))

    def set_x_window(self, x_lo: int, x_hi: int):
        self._check(self._lib.gsc_set_x_window(self._h, x_lo, x_hi))

    def local_x_range(self):
        lo, hi = C.c_int32(0), C.c_int32(0)
        self._check(self._lib.gsc_local_x_range(self._h, C.byref(lo), C.byref(hi)))
        return int(lo.value), int(hi.value)

    def halo_select(self, x_lo: int, x_hi: int, d_records: int, cap_records: int) -> int:
        """Pack local colliders with home x-cell in [x_lo, x_hi] into the DEVICE buffer at `d_records`.
        Returns the count; raises GscError(GSC_ERR_CAPACITY) with .needed set when the buffer is too small."""
        cnt = C.c_uint32(0)
        rc = self._lib.gsc_halo_select(self._h, x_lo, x_hi, d_records or None, cap_records, C.byref(cnt))
        if rc == N.GSC_ERR_CAPACITY:
            err = GscError(rc, self._lib.gsc_last_error(self._h).decode())
            err.needed = int(cnt.value)
            raise err
        self._check(rc)
        return int(cnt.value)

    def halo_append(self, count: int, d_records: int):
        self._check(self._lib.gsc_halo_append(self._h, count, d_records or None))

    # halo push over peer memory (GPUs of one box)
    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(N.GSC_IPC_EXPORT_BYTES)
        self._check(self._lib.gsc_ipc_export(self._h, buf, N.GSC_IPC_EXPORT_BYTES))
        return buf.raw

    def ipc_open_peer(self, peer: int, blob: bytes):
        buf = C.create_string_buffer(blob, N.GSC_IPC_EXPORT_BYTES)
        self._check(self._lib.gsc_ipc_open_peer(self._h, peer, buf, N.GSC_IPC_EXPORT_BYTES))

    def halo_push(self, peer: int, x_lo: int, x_hi: int, peer_n_local: int):
        self._check(self._lib.gsc_halo_push(self._h, peer, x_lo, x_hi, peer_n_local))

    def halo_sync(self):
        self._check(self._lib.gsc_halo_sync(self._h))

    def halo_commit(self) -> int:
        n = C.c_uint32(0)
        self._check(self._lib.gsc_halo_commit(self._h, C.byref(n)))
        return int(n.value)

    def phase_finish(self) -> N.GscStepOut:
        out = N.GscStepOut()
        rc = self._lib.gsc_phase_finish(self._h, C.byref(out))
        self.last = out
        self._check(rc)
        return out

    # the same exchange driven from the device: one call per step and rank, one host wait (gsc_slab_*)
    def open_peer_local(self, peer: int, other: "CollisionContext"):
        self._check(self._lib.gsc_open_peer_local(self._h, peer, other._h))

    def slab_init(self, rank: int, world: int):
        self._check(self._lib.gsc_slab_init(self._h, rank, world))

    def slab_step(self) -> N.GscStepOut:
        out = N.GscStepOut()
        rc = self._lib.gsc_slab_step(self._h, C.byref(out))
        self.last = out
        self._check(rc)
        return out

    def slab_step_begin(self):
        self._check(self._lib.gsc_slab_step_begin(self._h))

    # pinned transform staging (gsc_map_transform_staging): numpy views of the library's pinned banks
    def map_transform_staging(self, bank: int, cap: int):
        """(pos[cap,3], box_quat[cap,4], box_scale[cap,3]) float32 views of pinned bank `bank`; cap = max_colliders"""
        ptrs = [C.c_void_p() for _ in range(3)]
        self._check(self._lib.gsc_map_transform_staging(self._h, bank, *[C.byref(p) for p in ptrs]))
        out = []
        for p, w in zip(ptrs, (3, 4, 3)):
            buf = (C.c_float * (cap * w)).from_address(p.value)
            out.append(np.frombuffer(buf, dtype=np.float32).reshape(cap, w))
        return tuple(out)

    def commit_transform_staging(self, bank: int, n_box: int, with_rotation: bool = True, with_scale: bool = True):
        self._check(self._lib.gsc_commit_transform_staging(self._h, bank, self.n, n_box, int(with_rotation), int(with_scale)))

    def reorder_colliders(self) -> np.ndarray:
        """Permute the dense arrays into the cell order of the last step (gsc_reorder_colliders); returns order[slot] =
        previous dense index: hand over transforms[order] from now on."""
        order = np.zeros(max(self.n, 1), dtype=np.uint32)
        self._check(self._lib.gsc_reorder_colliders(self._h, _ptr(order), order.shape[0]))
        return order[: self.n]

    def box_slots(self) -> np.ndarray:
        out = np.zeros(max(self.n, 1), dtype=np.uint32)
        nb = C.c_uint32(0)
        self._check(self._lib.gsc_box_slots(self._h, _ptr(out), out.shape[0], C.byref(nb)))
        return out[: self.n]


class CollisionGroup:
    """gsc_group: all GPUs of a box behind one handle (one context per device, x-slab partition, device-driven halo
    exchange).  Colliders and transforms are given in the engine's dense order; pairs() is the union of the members'
    lists == the single-GPU pair set."""

    def __init__(self, devices, max_colliders_per_member: int, max_pairs: int = 0, max_candidates: int = 0, max_cells: int = 0,
                 timing: bool = False):
        self._lib = N.load()
        cfg = N.GscConfig(0, max_colliders_per_member, max_pairs, max_candidates, max_cells, N.GSC_FLAG_TIMING if timing else 0, 0)
        dev = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        rc = self._lib.gsc_group_create(len(devices), dev, C.byref(cfg), C.byref(h))
        if rc != N.GSC_OK:
            raise GscError(rc, self._lib.gsc_last_error(None).decode())
        self._h = h
        self.world = len(devices)
        self.n = 0
        self.last = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gsc_group_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != N.GSC_OK:
            raise GscError(rc, self._lib.gsc_group_last_error(self._h).decode())

    def upload_colliders(self, entity, kind, offset, size, pos):
        e, k = _c(entity, np.uint32), _c(kind, np.uint8)
        o, s, p = _c(offset, np.float32), _c(size, np.float32), _c(pos, np.float32)
        self.n = int(e.shape[0])
        self._check(self._lib.gsc_group_upload_colliders(self._h, self.n, _ptr(e), _ptr(k), _ptr(o), _ptr(s), _ptr(p)))

    def update_transforms(self, pos, quat=None, scale=None):
        p, q, s = _c(pos, np.float32), _c(quat, np.float32), _c(scale, np.float32)
        self._check(self._lib.gsc_group_update_transforms(self._h, self.n, _ptr(p), _ptr(q), _ptr(s)))

    def step(self) -> N.GscStepOut:
        out = N.GscStepOut()
        rc = self._lib.gsc_group_step(self._h, C.byref(out))
        self.last = out
        self._check(rc)
        return out

    def pairs(self, sorted: bool = True) -> np.ndarray:
        n = C.c_uint64(0)
        self._check(self._lib.gsc_group_download_pairs(self._h, None, 0, 0, C.byref(n)))
        out = np.zeros((int(n.value), 2), dtype=np.uint32)
        if n.value:
            self._check(self._lib.gsc_group_download_pairs(self._h, _ptr(out), n.value, 1 if sorted else 0, C.byref(n)))
        return out

    def pairs_packed(self, sorted: bool = True) -> np.ndarray:
        p = self.pairs(sorted).astype(np.uint64)
        return (p[:, 0] << np.uint64(32)) | p[:, 1]

    def rebalance(self, pos):
        """re-cut the slabs from the current positions (dense order) and the last step's work; the next
        update_transforms must carry rotation and scale again"""
        p = _c(pos, np.float32)
        self._check(self._lib.gsc_group_rebalance(self._h, int(p.shape[0]), _ptr(p)))

    def reorder(self):
        """every member keeps its colliders in the cell order of the last step (temporal coherence); the next
        update_transforms must carry rotation and scale again"""
        self._check(self._lib.gsc_group_reorder(self._h))

    def member_counts(self):
        out = []
        for i in range(self.world):
            n = C.c_uint32(0)
            self._check(self._lib.gsc_group_member(self._h, i, None, C.byref(n), None))
            out.append(int(n.value))
        return out


# ---- micro-bench batches (benches/collision.rs, benches/grid_collision.rs) ----

def _batch(fn_name, a, b, a_dtype, out_dtype, device=0):
    L = N.load()
    a = np.ascontiguousarray(a, dtype=a_dtype)
    n = a.shape[0]
    out = np.zeros(n, dtype=out_dtype)
    if b is not None:
        b = np.ascontiguousarray(b, dtype=a_dtype)
        rc = getattr(L, fn_name)(device, n, _ptr(a), _ptr(b), _ptr(out))
    else:
        rc = getattr(L, fn_name)(device, n, _ptr(a), _ptr(out))
    if rc != N.GSC_OK:
        raise GscError(rc, L.gsc_last_error(None).decode())
    return out


def test_sphere_sphere(a4, b4, device=0):
    return _batch("gsc_test_sphere_sphere", a4, b4, np.float32, np.uint8, device)


def test_sphere_obb(s4, obb16, device=0):
    return _batch("gsc_test_sphere_obb", s4, obb16, np.float32, np.uint8, device)


def test_obb_obb(a16, b16, device=0):
    return _batch("gsc_test_obb_obb", a16, b16, np.float32, np.uint8, device)


def contact_sphere_sphere(hi4, lo4, device=0):
    L = N.load()
    a, b = np.ascontiguousarray(hi4, dtype=np.float32), np.ascontiguousarray(lo4, dtype=np.float32)
    out = np.zeros((a.shape[0], 4), dtype=np.float32)
    rc = L.gsc_contact_sphere_sphere(device, a.shape[0], _ptr(a), _ptr(b), _ptr(out))
    if rc != N.GSC_OK:
        raise GscError(rc, L.gsc_last_error(None).decode())
    return out


def contact_sphere_obb(s4, obb16, sphere_is_hi, device=0):
    L = N.load()
    a, b = np.ascontiguousarray(s4, dtype=np.float32), np.ascontiguousarray(obb16, dtype=np.float32)
    f = np.ascontiguousarray(sphere_is_hi, dtype=np.uint8)
    out = np.zeros((a.shape[0], 4), dtype=np.float32)
    rc = L.gsc_contact_sphere_obb(device, a.shape[0], _ptr(a), _ptr(b), _ptr(f), _ptr(out))
    if rc != N.GSC_OK:
        raise GscError(rc, L.gsc_last_error(None).decode())
    return out


def contact_obb_obb(hi16, lo16, device=0):
    L = N.load()
    a, b = np.ascontiguousarray(hi16, dtype=np.float32), np.ascontiguousarray(lo16, dtype=np.float32)
    out = np.zeros((a.shape[0], 4), dtype=np.float32)
    rc = L.gsc_contact_obb_obb(device, a.shape[0], _ptr(a), _ptr(b), _ptr(out))
    if rc != N.GSC_OK:
        raise GscError(rc, L.gsc_last_error(None).decode())
    return out


def test_aabb_aabb(a6, b6, device=0):
    return _batch("gsc_test_aabb_aabb", a6, b6, np.float32, np.uint8, device)


def fnv1a_gridcell(cells3, device=0):
    return _batch("gsc_fnv1a_gridcell", cells3, None, np.int16, np.uint64, device)


def sort_pairs_u32(keys, vals, key_bits=32, device=0):
    L = N.load()
    k = np.array(keys, dtype=np.uint32, copy=True)
    v = np.array(vals, dtype=np.uint32, copy=True)
    rc = L.gsc_sort_pairs_u32(device, k.shape[0], _ptr(k), _ptr(v), key_bits)
    if rc != N.GSC_OK:
        raise GscError(rc, L.gsc_last_error(None).decode())
    return k, v

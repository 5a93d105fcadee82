"""ctypes binding of oracle/libgunship_oracle.so -- the CHECKER, for tests / smoke / cpu_baseline only.

Nothing under gunship_rs_b200/ imports this module (the product path never touches the oracle).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libgunship_oracle.so")

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i16p = C.POINTER(C.c_int16)


def build_oracle() -> str:
    src = os.path.join(_ROOT, "oracle", "gunship_oracle.c")
    if (not os.path.exists(_SO)) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.orc_system_new.restype = C.c_void_p
        L.orc_system_new.argtypes = [C.c_int, C.c_int]
        L.orc_system_free.argtypes = [C.c_void_p]
        L.orc_bvh_update.restype = C.c_int
        L.orc_bvh_update.argtypes = [C.c_void_p, C.c_uint32, _u32p, _u8p, _f32p, _f32p, _f32p, _f32p, _f32p]
        L.orc_longest_axis.restype = C.c_float
        L.orc_longest_axis.argtypes = [C.c_void_p]
        L.orc_num_volumes.restype = C.c_uint32
        L.orc_num_volumes.argtypes = [C.c_void_p]
        L.orc_components.restype = C.c_void_p
        L.orc_components.argtypes = [C.c_void_p]
        L.orc_grid_update.restype = C.c_uint64
        L.orc_grid_update.argtypes = [C.c_void_p]
        L.orc_collisions_sorted.restype = C.c_uint64
        L.orc_collisions_sorted.argtypes = [C.c_void_p, _u64p, C.c_uint64]
        L.orc_brute_force_sorted.restype = C.c_uint64
        L.orc_brute_force_sorted.argtypes = [C.c_void_p, _u64p, C.c_uint64]
        for name in ("orc_last_candidates", "orc_last_cell_entries"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("orc_last_wide_volumes", "orc_last_out_of_range"):
            getattr(L, name).restype = C.c_uint32
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_quat_to_mat3.argtypes = [_f32p, _f32p]
        L.orc_mat3_rotation.argtypes = [C.c_float, C.c_float, C.c_float, _f32p]
        L.orc_batch_sphere_sphere.argtypes = [C.c_size_t, _f32p, _f32p, _u8p]
        L.orc_batch_sphere_obb.argtypes = [C.c_size_t, _f32p, _f32p, _u8p]
        L.orc_batch_obb_obb.argtypes = [C.c_size_t, _f32p, _f32p, _u8p]
        L.orc_batch_aabb_aabb.argtypes = [C.c_size_t, _f32p, _f32p, _u8p]
        L.orc_batch_fnv1a_gridcell.argtypes = [C.c_size_t, _i16p, _u64p]
        L.orc_batch_contact_sphere_sphere.argtypes = [C.c_size_t, _f32p, _f32p, _f32p]
        L.orc_batch_contact_sphere_obb.argtypes = [C.c_size_t, _f32p, _f32p, _u8p, _f32p]
        L.orc_batch_contact_obb_obb.argtypes = [C.c_size_t, _f32p, _f32p, _f32p]
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# numpy view of orc_bound_volume (104 B): entity, aabb(min xyzw, max xyzw), tag, 16 payload floats
BOUND_VOLUME = np.dtype([("entity", "<u4"), ("aabb_min", "<f4", 4), ("aabb_max", "<f4", 4),
                         ("tag", "<u4"), ("payload", "<f4", 16)])
assert BOUND_VOLUME.itemsize == 104


class OracleSystem:
    """CollisionSystem of the reference restated on the CPU (bvh_update + GridCollisionSystem::update)."""

    def __init__(self, work_units: int = 8, workers: int = 8):
        self._h = lib().orc_system_new(work_units, workers)
        if not self._h:
            raise ValueError("unsupported work unit / worker count")
        self.cores = workers

    def close(self):
        if self._h:
            lib().orc_system_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bvh_update(self, entity, kind, offset, size, pos, quat, scale) -> None:
        self._keep = [_c(entity, np.uint32), _c(kind, np.uint8), _c(offset, np.float32), _c(size, np.float32),
                      _c(pos, np.float32), _c(quat, np.float32), _c(scale, np.float32)]
        e, k, o, s, p, q, sc = self._keep
        rc = lib().orc_bvh_update(self._h, e.shape[0], _p(e, _u32p), _p(k, _u8p), _p(o, _f32p), _p(s, _f32p),
                                  _p(p, _f32p), _p(q, _f32p), _p(sc, _f32p))
        if rc != 0:
            raise NotImplementedError("Collider::Mesh is unimplemented!() in the reference (mod.rs:331)")

    def grid_update(self) -> int:
        return int(lib().orc_grid_update(self._h))

    @property
    def longest_axis(self) -> float:
        return float(lib().orc_longest_axis(self._h))

    def volumes(self) -> np.ndarray:
        n = lib().orc_num_volumes(self._h)
        ptr = lib().orc_components(self._h)
        if n == 0:
            return np.zeros(0, dtype=BOUND_VOLUME)
        buf = (C.c_uint8 * (n * 104)).from_address(ptr)
        return np.frombuffer(buf, dtype=BOUND_VOLUME, count=n).copy()

    def collisions_sorted(self) -> np.ndarray:
        n = int(lib().orc_collisions_sorted(self._h, None, 0))
        out = np.zeros(max(n, 1), dtype=np.uint64)
        lib().orc_collisions_sorted(self._h, _p(out, _u64p), n)
        return out[:n]

    def brute_force_sorted(self) -> np.ndarray:
        n = int(lib().orc_brute_force_sorted(self._h, None, 0))
        out = np.zeros(max(n, 1), dtype=np.uint64)
        lib().orc_brute_force_sorted(self._h, _p(out, _u64p), n)
        return out[:n]

    def process_collisions(self):
        """(entity, offset, other) CSR of CollisionCallbackManager::process_collisions (mod.rs:678-708)."""
        L = lib()
        L.orc_process_collisions.restype = C.c_uint64
        L.orc_process_collisions.argtypes = [C.c_void_p, _u32p, _u32p, _u32p, C.c_uint64]
        h = int(L.orc_collisions_sorted(self._h, None, 0))
        other = np.zeros(max(2 * h, 1), dtype=np.uint32)
        ent = np.zeros(max(2 * h, 1), dtype=np.uint32)
        off = np.zeros(2 * h + 2, dtype=np.uint32)
        n = int(L.orc_process_collisions(self._h, _p(ent, _u32p), _p(off, _u32p), _p(other, _u32p), 2 * h + 1))
        return ent[:n].copy(), off[:n + 1].copy(), other[:2 * h].copy()

    def stats(self) -> dict:
        L = lib()
        return {"candidates": int(L.orc_last_candidates(self._h)), "cell_entries": int(L.orc_last_cell_entries(self._h)),
                "wide": int(L.orc_last_wide_volumes(self._h)), "out_of_range": int(L.orc_last_out_of_range(self._h))}

    def step(self, scene, pos=None, quat=None) -> np.ndarray:
        """One CollisionSystem::update over a scenes.Scene; returns the sorted packed pair set."""
        self.bvh_update(scene.entity, scene.kind, scene.offset, scene.size,
                        scene.pos if pos is None else pos, scene.quat if quat is None else quat, scene.scale)
        self.grid_update()
        return self.collisions_sorted()


def quat_to_mat3(q) -> np.ndarray:
    q = _c(q, np.float32)
    out = np.zeros(9, dtype=np.float32)
    lib().orc_quat_to_mat3(_p(q, _f32p), _p(out, _f32p))
    return out.reshape(3, 3)


def mat3_rotation(x, y, z) -> np.ndarray:
    out = np.zeros(9, dtype=np.float32)
    lib().orc_mat3_rotation(x, y, z, _p(out, _f32p))
    return out.reshape(3, 3)


def batch_sphere_sphere(a, b) -> np.ndarray:
    a, b = _c(a, np.float32), _c(b, np.float32)
    out = np.zeros(a.shape[0], dtype=np.uint8)
    lib().orc_batch_sphere_sphere(a.shape[0], _p(a, _f32p), _p(b, _f32p), _p(out, _u8p))
    return out


def batch_sphere_obb(s, o) -> np.ndarray:
    s, o = _c(s, np.float32), _c(o, np.float32)
    out = np.zeros(s.shape[0], dtype=np.uint8)
    lib().orc_batch_sphere_obb(s.shape[0], _p(s, _f32p), _p(o, _f32p), _p(out, _u8p))
    return out


def batch_obb_obb(a, b) -> np.ndarray:
    a, b = _c(a, np.float32), _c(b, np.float32)
    out = np.zeros(a.shape[0], dtype=np.uint8)
    lib().orc_batch_obb_obb(a.shape[0], _p(a, _f32p), _p(b, _f32p), _p(out, _u8p))
    return out


def batch_contact_sphere_sphere(hi, lo) -> np.ndarray:
    hi, lo = _c(hi, np.float32), _c(lo, np.float32)
    out = np.zeros((hi.shape[0], 4), dtype=np.float32)
    lib().orc_batch_contact_sphere_sphere(hi.shape[0], _p(hi, _f32p), _p(lo, _f32p), _p(out, _f32p))
    return out


def batch_contact_sphere_obb(s, o, sphere_is_hi) -> np.ndarray:
    s, o, f = _c(s, np.float32), _c(o, np.float32), _c(sphere_is_hi, np.uint8)
    out = np.zeros((s.shape[0], 4), dtype=np.float32)
    lib().orc_batch_contact_sphere_obb(s.shape[0], _p(s, _f32p), _p(o, _f32p), _p(f, _u8p), _p(out, _f32p))
    return out


def batch_contact_obb_obb(hi, lo) -> np.ndarray:
    hi, lo = _c(hi, np.float32), _c(lo, np.float32)
    out = np.zeros((hi.shape[0], 4), dtype=np.float32)
    lib().orc_batch_contact_obb_obb(hi.shape[0], _p(hi, _f32p), _p(lo, _f32p), _p(out, _f32p))
    return out


def batch_aabb_aabb(a, b) -> np.ndarray:
    a, b = _c(a, np.float32), _c(b, np.float32)
    out = np.zeros(a.shape[0], dtype=np.uint8)
    lib().orc_batch_aabb_aabb(a.shape[0], _p(a, _f32p), _p(b, _f32p), _p(out, _u8p))
    return out


def batch_fnv1a_gridcell(cells) -> np.ndarray:
    cells = _c(cells, np.int16)
    out = np.zeros(cells.shape[0], dtype=np.uint64)
    lib().orc_batch_fnv1a_gridcell(cells.shape[0], _p(cells, _i16p), _p(out, _u64p))
    return out


def obb16(center, orientation, half) -> np.ndarray:
    """Pack an OBB as the 16-float record the batch helpers take."""
    out = np.zeros(16, dtype=np.float32)
    out[0:3] = center
    out[3:12] = np.asarray(orientation, dtype=np.float32).reshape(9)
    out[12:15] = half
    return out


def transform_update(row_begin, parent, lpos, lquat, lscale):
    """TransformManager::update_transforms restated (component/transform.rs:243-251, 588-608): derived
    (pos[n,3], quat[n,4], scale[n,3], matrix[n,4,4]) of every node, rows in order."""
    L = lib()
    L.orc_transform_update.restype = None
    L.orc_transform_update.argtypes = [C.c_uint32, _u32p, _u32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p]
    rb, p = _c(row_begin, np.uint32), _c(parent, np.uint32)
    lp, lq, ls = _c(lpos, np.float32), _c(lquat, np.float32), _c(lscale, np.float32)
    n = p.shape[0]
    dpos, dquat = np.zeros((n, 3), np.float32), np.zeros((n, 4), np.float32)
    dscale, dmat = np.zeros((n, 3), np.float32), np.zeros((n, 4, 4), np.float32)
    L.orc_transform_update(rb.shape[0] - 1, _p(rb, _u32p), _p(p, _u32p), _p(lp, _f32p), _p(lq, _f32p), _p(ls, _f32p),
                           _p(dpos, _f32p), _p(dquat, _f32p), _p(dscale, _f32p), _p(dmat, _f32p))
    return dpos, dquat, dscale, dmat


def swap_remove_sequence(order, index):
    """A sequence of BoundingVolumeManager::destroy_immediate calls (bounding_volume.rs:82-104) on the dense array
    `order`; returns the surviving array."""
    L = lib()
    L.orc_swap_remove_sequence.restype = C.c_uint32
    L.orc_swap_remove_sequence.argtypes = [_u32p, C.c_uint32, _u32p, C.c_uint32]
    o = np.array(order, dtype=np.uint32, copy=True)
    i = _c(index, np.uint32)
    n = int(L.orc_swap_remove_sequence(_p(o, _u32p), o.shape[0], _p(i, _u32p), i.shape[0]))
    if n == 0xFFFFFFFF:
        raise IndexError("swap_remove index out of range")
    return o[:n]

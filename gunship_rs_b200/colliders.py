"""Host-side mirror of the reference's collider bookkeeping for the device-resident dense arrays.

`ColliderSet` keeps what `ColliderManager` + `BoundingVolumeManager` keep on the host in the reference -- the
entity -> dense index map and the order of the dense arrays -- and forwards every change to the device through the
C ABI (`gsc_append_colliders`, `gsc_remove_colliders`).  Nothing here computes collisions.

Reference semantics mirrored (file:line relative to the reference):
  * assign appends at the end of the dense arrays and panics on a duplicate entity
    (src/component/struct_component_manager.rs:83-96; BoundingVolumeManager pushes a volume when bvh_update first
    meets the entity, collider/bounding_volume.rs:394-407)
  * destroy only MARKS the entity (collider/mod.rs:282-285); the cleanup loop at the end of
    CollisionSystem::update (mod.rs:605-609) drains the marks and calls
    BoundingVolumeManager::destroy_immediate (bounding_volume.rs:82-104): a swap_remove -- the last element moves
    into the hole and its index entry is rewritten.  The INDEX order orients every output pair
    (grid_collision.rs:440,497), so it must change exactly as in the reference.
"""
from __future__ import annotations

import numpy as np

from . import _native as N


class ColliderSet:
    def __init__(self, ctx):
        self.ctx = ctx
        self.entities: list[int] = []      # dense order (BoundingVolumeManager.entities)
        self.indices: dict[int, int] = {}  # entity -> dense index (BoundingVolumeManager.indices)
        self.kind: list[int] = []
        self.offset: list = []
        self.size: list = []
        self._uploaded = 0                 # how many of the dense elements the device already has
        self._marked: list[int] = []       # marked_for_destroy, in marking order

    def __len__(self):
        return len(self.entities)

    # ColliderManager::assign (mod.rs:221-223)
    def assign(self, entity: int, kind: int, offset, size):
        if entity in self.indices:
            raise AssertionError(f"Component already assign to entity {entity}")  # struct_component_manager.rs:84-87
        if kind not in (N.GSC_SPHERE, N.GSC_BOX):
            raise NotImplementedError("Collider::Mesh is unimplemented!() in the reference (mod.rs:331)")
        self.indices[entity] = len(self.entities)
        self.entities.append(int(entity))
        self.kind.append(int(kind))
        self.offset.append(tuple(float(x) for x in offset))
        self.size.append(tuple(float(x) for x in size))

    # ComponentManager::destroy (mod.rs:282-285): mark only
    def destroy(self, entity: int):
        if entity in self.indices and entity not in self._marked:
            self._marked.append(int(entity))

    def cleanup(self):
        """The cleanup loop at the END of CollisionSystem::update (mod.rs:605-609): drain the marks, one
        destroy_immediate each.  (A marked collider still takes part in the step before it, as in the reference.)
        Returns the removal index sequence that was applied to the device arrays."""
        self.push_assigned()  # a mark may refer to a collider the device has not seen yet
        seq = []
        for entity in self._marked:
            index = self.indices.pop(entity)              # bounding_volume.rs:84
            seq.append(index)
            last = len(self.entities) - 1
            for arr in (self.entities, self.kind, self.offset, self.size):
                arr[index] = arr[last]                    # Vec::swap_remove (:86, :103)
                arr.pop()
            if index != len(self.entities):               # :91-94
                self.indices[self.entities[index]] = index
        self._marked.clear()
        if seq:
            self.ctx.remove_colliders(np.asarray(seq, dtype=np.uint32))
        self._uploaded = len(self.entities)
        return seq

    def push_assigned(self):
        """What bvh_update does at the START of the step for entities it has not met yet
        (bounding_volume.rs:394-407): their volumes go to the end of the dense arrays."""
        lo, hi = self._uploaded, len(self.entities)
        if hi == lo:
            return
        ent = np.asarray(self.entities[lo:hi], dtype=np.uint32)
        kind = np.asarray(self.kind[lo:hi], dtype=np.uint8)
        off = np.asarray(self.offset[lo:hi], dtype=np.float32).reshape(-1, 3)
        size = np.asarray(self.size[lo:hi], dtype=np.float32).reshape(-1, 3)
        if lo == 0:
            self.ctx.upload_colliders(ent, kind, off, size)
        else:
            self.ctx.append_colliders(ent, kind, off, size)
        self._uploaded = hi

    def arrays(self):
        """(entity u32[n], kind u8[n], offset f32[n,3], size f32[n,3]) in dense order"""
        n = len(self.entities)
        return (np.asarray(self.entities, dtype=np.uint32), np.asarray(self.kind, dtype=np.uint8),
                np.asarray(self.offset, dtype=np.float32).reshape(n, 3), np.asarray(self.size, dtype=np.float32).reshape(n, 3))

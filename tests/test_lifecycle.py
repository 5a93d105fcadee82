"""Collider lifecycle: assign appends to the dense arrays, destroy is a swap_remove
(bounding_volume.rs:82-104) -- the index order, which orients every output pair, must change as in the reference."""
import numpy as np
import pytest

import oracle_lib as O
from gunship_rs_b200 import _native as N
from gunship_rs_b200.colliders import ColliderSet


def _python_destroy_immediate(entities, indices, entity):
    """BoundingVolumeManager::destroy_immediate line by line (bounding_volume.rs:82-104) on Python lists"""
    index = indices.pop(entity, None)          # :84
    if index is None:
        return None
    last = entities.pop()                      # Vec::swap_remove(index) :86
    if index < len(entities):
        removed, entities[index] = entities[index], last
    else:
        removed = last
    assert removed == entity                   # :87
    if index != len(entities):                 # :91-94
        indices[entities[index]] = index
    return index


def test_oracle_swap_remove_sequence_matches_line_by_line_restatement():
    rng = np.random.default_rng(3)
    for n, k in [(1, 1), (2, 2), (10, 4), (257, 200), (1000, 999), (1000, 1000)]:
        entities = list(range(1, n + 1))
        indices = {e: i for i, e in enumerate(entities)}
        victims = rng.permutation(entities)[:k]
        seq = [_python_destroy_immediate(entities, indices, int(e)) for e in victims]
        got = O.swap_remove_sequence(np.arange(1, n + 1, dtype=np.uint32), seq)
        assert got.tolist() == entities
        assert all(indices[e] == i for i, e in enumerate(entities))
    with pytest.raises(IndexError):
        O.swap_remove_sequence(np.arange(4, dtype=np.uint32), [1, 3])  # only 3 are left when index 3 is asked for


class _RecordingCtx:
    """stands in for CollisionContext on the CPU: replays what the C ABI is told onto plain arrays"""

    def __init__(self):
        self.entity = np.zeros(0, np.uint32)
        self.calls = []

    def upload_colliders(self, e, k, o, s):
        self.entity = np.array(e, np.uint32)
        self.calls.append(("upload", len(e)))

    def append_colliders(self, e, k, o, s):
        self.entity = np.concatenate([self.entity, np.asarray(e, np.uint32)])
        self.calls.append(("append", len(e)))

    def remove_colliders(self, index):
        self.entity = O.swap_remove_sequence(self.entity, index)
        self.calls.append(("remove", len(index)))


def test_collider_set_mirrors_reference_order():
    ctx = _RecordingCtx()
    cs = ColliderSet(ctx)
    for e in range(1, 7):
        cs.assign(e, N.GSC_SPHERE, (0, 0, 0), (1, 0, 0))
    with pytest.raises(AssertionError):
        cs.assign(3, N.GSC_SPHERE, (0, 0, 0), (1, 0, 0))       # struct_component_manager.rs:84-87
    with pytest.raises(NotImplementedError):
        cs.assign(99, N.GSC_MESH, (0, 0, 0), (1, 0, 0))        # mod.rs:331
    cs.push_assigned()
    assert ctx.calls == [("upload", 6)]
    cs.destroy(2)
    cs.destroy(6)
    cs.destroy(2)            # marking twice is one mark (EntitySet)
    cs.assign(7, N.GSC_BOX, (0, 0, 0), (1, 1, 1))
    # the step of this frame still sees 2 and 6, and meets 7 for the first time (bvh_update pushes it)
    cs.push_assigned()
    assert ctx.entity.tolist() == [1, 2, 3, 4, 5, 6, 7]
    seq = cs.cleanup()       # end of CollisionSystem::update: remove 2 (7 moves in), then 6 (last after the pop? no: index 5)
    assert seq == [1, 5]
    assert cs.entities == [1, 7, 3, 4, 5] and ctx.entity.tolist() == cs.entities
    assert cs.indices == {1: 0, 7: 1, 3: 2, 4: 3, 5: 4}
    assert cs.kind == [0, 1, 0, 0, 0]


@pytest.mark.gpu
def test_device_arrays_and_pairs_follow_the_reference_through_assign_and_destroy():
    from gunship_rs_b200 import context as G
    from gunship_rs_b200 import scenes
    scene = scenes.by_config(3, 6000)
    rng = np.random.default_rng(17)
    n0 = 4000
    with G.CollisionContext(8192) as ctx:
        cs = ColliderSet(ctx)
        tr = {}  # entity -> (pos, quat, scale): transforms travel with the entity, not with the dense index
        for i in range(scene.n):
            tr[int(scene.entity[i])] = (scene.pos[i], scene.quat[i], scene.scale[i])
        for i in range(n0):
            cs.assign(int(scene.entity[i]), int(scene.kind[i]), scene.offset[i], scene.size[i])
        nxt = n0
        sysm = O.OracleSystem(8, 4)
        for frame in range(5):
            cs.push_assigned()
            ent, kind, off, size = cs.arrays()
            assert ctx.collider_count() == len(cs)
            assert np.array_equal(ctx.debug(N.GSC_DBG_ENTITY), ent)
            assert np.array_equal(ctx.debug(N.GSC_DBG_KIND), kind)
            assert np.array_equal(ctx.debug(N.GSC_DBG_OFFSET), off)
            assert np.array_equal(ctx.debug(N.GSC_DBG_SIZE), size)
            pos = np.stack([tr[int(e)][0] for e in ent]).astype(np.float32)
            quat = np.stack([tr[int(e)][1] for e in ent]).astype(np.float32)
            scale = np.stack([tr[int(e)][2] for e in ent]).astype(np.float32)
            if frame % 2:
                ctx.update_transforms_packed(pos, quat[kind == 1], scale[kind == 1])
            else:
                ctx.update_transforms(pos, quat, scale)
            ctx.step()
            sysm.bvh_update(ent, kind, off, size, pos, quat, scale)
            sysm.grid_update()
            want = sysm.collisions_sorted()
            assert want.shape[0] > 100
            assert np.array_equal(ctx.pairs_packed(), want), f"frame {frame}"
            # this frame's game logic: destroy some, assign some; the cleanup runs at the end of the update
            victims = rng.choice(ent, size=300 + 100 * frame, replace=False)
            for e in victims:
                cs.destroy(int(e))
            seq = cs.cleanup()
            assert len(seq) == victims.shape[0]
            for _ in range(350):
                if nxt < scene.n:
                    cs.assign(int(scene.entity[nxt]), int(scene.kind[nxt]), scene.offset[nxt], scene.size[nxt])
                    nxt += 1
        sysm.close()
        # removing everything, and error paths
        cs.push_assigned()
        for e in list(cs.entities):
            cs.destroy(e)
        cs.cleanup()
        assert ctx.collider_count() == 0 and len(cs) == 0
        with pytest.raises(N.GscError) as ei:
            ctx.remove_colliders([0])
        assert ei.value.code == N.GSC_ERR_INVALID

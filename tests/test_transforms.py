"""Transform derivation: TransformManager::update_transforms / TransformData::update
(component/transform.rs:243-251, 588-608) -- oracle properties on the CPU, bit-exact CUDA parity on the GPU.
The reference has no test of this code: parity unpinned beyond the restated arithmetic."""
import numpy as np
import pytest

import oracle_lib as O
from gunship_rs_b200 import _native as N

NOP = N.GSC_NO_PARENT


def _forest(rng, n_roots, depth, fan):
    """rows of a random forest: (row_begin, parent) with nodes numbered row by row"""
    rows = [n_roots]
    parent = [np.full(n_roots, NOP, dtype=np.uint32)]
    begin = [0, n_roots]
    for d in range(1, depth):
        cnt = max(1, int(rows[-1] * fan))
        lo = begin[-2]
        parent.append(rng.integers(lo, lo + rows[-1], size=cnt).astype(np.uint32))
        rows.append(cnt)
        begin.append(begin[-1] + cnt)
    return np.asarray(begin, dtype=np.uint32), np.concatenate(parent)


def _locals(rng, n):
    pos = rng.uniform(-5, 5, size=(n, 3)).astype(np.float32)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    scale = rng.uniform(0.5, 2.0, size=(n, 3)).astype(np.float32)
    return pos, q.astype(np.float32), scale


def test_roots_derive_to_their_local_values():
    """with the identity dummy parent the derived values equal the local ones
    (the products with identity only add exact zeros)"""
    rng = np.random.default_rng(1)
    n = 1000
    pos, quat, scale = _locals(rng, n)
    rb = np.array([0, n], dtype=np.uint32)
    parent = np.full(n, NOP, dtype=np.uint32)
    dpos, dquat, dscale, dmat = O.transform_update(rb, parent, pos, quat, scale)
    assert np.array_equal(dpos, pos) and np.array_equal(dquat, quat) and np.array_equal(dscale, scale)
    assert np.array_equal(dmat[:, 3], np.tile(np.array([0, 0, 0, 1], np.float32), (n, 1)))
    assert np.array_equal(dmat[:, :3, 3], pos)


def test_known_answers():
    ident = [0, 0, 0, 1]
    s = np.float32(np.sqrt(0.5))
    rb = np.array([0, 1, 2, 3], dtype=np.uint32)
    parent = np.array([NOP, 0, 1], dtype=np.uint32)
    # root: translated (1,2,3), scaled 2; child: rotated 90 degrees about z, at local (1,0,0); grandchild at local (1,0,0)
    pos = np.array([[1, 2, 3], [1, 0, 0], [1, 0, 0]], dtype=np.float32)
    quat = np.array([ident, [0, 0, s, s], ident], dtype=np.float32)
    scale = np.array([[2, 2, 2], [1, 1, 1], [1, 1, 1]], dtype=np.float32)
    dpos, dquat, dscale, dmat = O.transform_update(rb, parent, pos, quat, scale)
    assert np.allclose(dpos[0], [1, 2, 3]) and np.allclose(dpos[1], [3, 2, 3])       # 1 + 2 * 1
    assert np.allclose(dpos[2], [3, 4, 3], atol=1e-6)                                 # x axis of the child points along +y
    assert np.allclose(dquat[2], [0, 0, s, s]) and np.allclose(dscale[2], [2, 2, 2])
    # the derived matrix maps the local origin to the derived position and carries the accumulated scale
    assert np.allclose(np.linalg.norm(dmat[2, :3, :3], axis=0), [2, 2, 2], atol=1e-6)


def test_derived_matrix_equals_float64_composition():
    rng = np.random.default_rng(2)
    rb, parent = _forest(rng, 50, 4, 1.5)
    n = parent.shape[0]
    pos, quat, scale = _locals(rng, n)
    dpos, dquat, dscale, dmat = O.transform_update(rb, parent, pos, quat, scale)
    ref = np.zeros((n, 4, 4))
    for i in range(n):
        x, y, z, w = quat[i].astype(np.float64)
        R = np.array([[w*w+x*x-y*y-z*z, 2*x*y-2*w*z, 2*x*z+2*w*y], [2*x*y+2*w*z, w*w-x*x+y*y-z*z, 2*y*z-2*w*x],
                      [2*x*z-2*w*y, 2*y*z+2*w*x, w*w-x*x-y*y+z*z]])
        L = np.eye(4)
        L[:3, :3] = R * scale[i].astype(np.float64)[None, :]
        L[:3, 3] = pos[i]
        ref[i] = L if parent[i] == NOP else ref[parent[i]] @ L
    assert np.allclose(dmat, ref, rtol=1e-4, atol=1e-4)
    assert np.array_equal(dpos, dmat[:, :3, 3])


@pytest.mark.gpu
def test_cuda_hierarchy_is_bit_exact_and_feeds_the_step():
    from gunship_rs_b200 import context as G
    from gunship_rs_b200 import scenes
    rng = np.random.default_rng(5)
    scene = scenes.by_config(3, 20000)
    n = scene.n
    # every collider hangs under a small random hierarchy: 400 roots, 3 further levels; colliders read random nodes
    rb, parent = _forest(rng, 400, 4, 3.0)
    nodes = parent.shape[0]
    lpos, lquat, lscale = _locals(rng, nodes)
    lpos *= 4.0
    lscale = rng.uniform(0.8, 1.25, size=(nodes, 3)).astype(np.float32)
    node = rng.integers(0, nodes, size=n).astype(np.uint32)
    with G.CollisionContext(n) as ctx:
        ctx.upload_colliders(scene.entity, scene.kind, scene.offset, scene.size)
        ctx.hierarchy_upload(rb, parent, node)
        sysm = O.OracleSystem(8, 4)
        for frame in range(3):
            if frame:  # move every node a little; rotation and scale are kept from the previous frame on frame 2
                lpos = (lpos + rng.uniform(-0.2, 0.2, size=lpos.shape)).astype(np.float32)
            if frame == 1:
                q = lquat + rng.normal(scale=0.05, size=lquat.shape)
                lquat = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
            ctx.hierarchy_update(lpos, lquat if frame < 2 else None, lscale if frame < 2 else None)
            want = O.transform_update(rb, parent, lpos, lquat, lscale)
            got = ctx.hierarchy_download()
            for g, w, name in zip(got, want, ("position", "rotation", "scale", "matrix")):
                assert np.array_equal(g.view(np.uint32), w.view(np.uint32)), f"{name} differs in frame {frame}"
            out = ctx.step()
            dpos, dquat, dscale, _ = want
            sysm.bvh_update(scene.entity, scene.kind, scene.offset, scene.size, dpos[node], dquat[node], dscale[node])
            sysm.grid_update()
            pairs = sysm.collisions_sorted()
            assert pairs.shape[0] > 50
            assert out.n_pairs == pairs.shape[0] and np.array_equal(ctx.pairs_packed(), pairs)
        sysm.close()
        # a NEW hierarchy (same size: the buffers are reused) must not inherit the previous one's local transforms:
        # "NULL keeps the previous values" is refused until all three arrays were sent again
        ctx.hierarchy_upload(rb, parent, node)
        with pytest.raises(N.GscError):
            ctx.hierarchy_update(lpos, None, None)
        ctx.hierarchy_update(lpos, lquat, lscale)
        # error paths: a parent outside the previous row; an update after the collider set changed
        bad = parent.copy()
        bad[-1] = 0
        with pytest.raises(N.GscError):
            ctx.hierarchy_upload(rb, bad, node)
        ctx.remove_colliders([0])
        with pytest.raises(N.GscError):
            ctx.hierarchy_update(lpos, lquat, lscale)


@pytest.mark.gpu
def test_step_begin_end_overlaps_the_next_upload_and_matches_step():
    from gunship_rs_b200 import context as G
    from gunship_rs_b200 import scenes
    scene = scenes.by_config(4, 200000)
    frames = [scene.frame(f) for f in range(4)]
    box = scene.kind == 1
    with G.CollisionContext(scene.n) as ctx:
        ctx.upload_colliders(scene.entity, scene.kind, scene.offset, scene.size)
        want = []
        for pos, quat in frames:
            ctx.update_transforms(pos, quat, scene.scale)
            ctx.step()
            want.append(ctx.pairs_packed())
        assert len({w.shape[0] for w in want}) > 1  # the frames really differ
        # pipelined: frame f+1 is handed over while step f is in flight
        pos, quat = frames[0]
        ctx.update_transforms_packed(pos, quat[box], scene.scale[box])
        ctx.step_begin()
        with pytest.raises(N.GscError):
            ctx.pairs()          # results are not available while the step is in flight
        with pytest.raises(N.GscError):
            ctx.step_begin()
        for f in range(len(frames)):
            if f + 1 < len(frames):
                npos, nquat = frames[f + 1]
                ctx.update_transforms_packed(npos, nquat[box])
            out = ctx.step_end()
            assert out.n_pairs == want[f].shape[0]
            assert np.array_equal(ctx.pairs_packed(), want[f]), f"frame {f}"
            if f + 1 < len(frames):
                ctx.step_begin()
        with pytest.raises(N.GscError):
            ctx.step_end()

"""gsc_group_* (one process, one context per GPU, device-driven slab exchange) and the round-2 host-facing additions
(pinned staging, CUDA-graph replay) through the C ABI, against the oracle.  Members of a group may share a GPU (the
group orders their phases by stream events), so the N-member paths run on a one-GPU box too; with more GPUs the members
spread over them and the halo really crosses NVLink."""
import numpy as np
import pytest

import oracle_lib as O
from gunship_rs_b200 import _native as N
from gunship_rs_b200 import context as G
from gunship_rs_b200 import scenes

pytestmark = pytest.mark.gpu


def _devices(world):
    import torch
    nd = torch.cuda.device_count()
    return [i % nd for i in range(world)]


@pytest.mark.parametrize("cfg,n,world", [(4, 60000, 2), (3, 40000, 3), (2, 50000, 4), (4, 200000, 8), (4, 1 << 21, 4)])
def test_group_union_is_the_single_gpu_pair_set(cfg, n, world):
    s = scenes.by_config(cfg, n)
    oracle = O.OracleSystem()
    cap = int(n / world * 1.6) + (1 << 16)
    with G.CollisionGroup(_devices(world), cap) as grp:
        grp.upload_colliders(s.entity, s.kind, s.offset, s.size, s.pos)
        counts = grp.member_counts()
        assert sum(counts) == n and max(counts) - min(counts) <= max(2, n // 1000)   # count-balanced slabs
        for f in (0, 1, 3):
            pos, quat = s.frame(f)
            grp.update_transforms(pos, quat, s.scale)
            out = grp.step()
            want = oracle.step(s, pos, quat)
            got = grp.pairs_packed(sorted=True)
            assert out.n_pairs == len(want), (cfg, n, world, f, out.n_pairs, len(want))
            assert np.array_equal(got, want)                                         # union == oracle, sorted, no duplicates
            assert np.float32(out.cell_size) == np.float32(oracle.longest_axis)
            assert out.n_wide == 0 and 0 < out.n_ghost < n
            # unsorted download: the same set
            assert np.array_equal(np.sort(grp.pairs_packed(sorted=False)), want)
            if f == 1:
                grp.reorder()       # members switch to the cell order of this frame; the dense interface does not change
                with pytest.raises(G.GscError):
                    grp.step()      # (transforms have to be handed over again, in full)
    oracle.close()


def test_group_of_one_equals_plain_step():
    s = scenes.by_config(3, 30000)
    oracle = O.OracleSystem()
    with G.CollisionGroup([0], s.n) as grp:
        grp.upload_colliders(s.entity, s.kind, s.offset, s.size, s.pos)
        grp.update_transforms(s.pos, s.quat, s.scale)
        out = grp.step()
        assert out.n_ghost == 0 and np.array_equal(grp.pairs_packed(), oracle.step(s))
    oracle.close()


def test_migrants_and_rebalance_over_fifty_frames():
    """Colliders drift in +x across several slabs; ownership is static between re-balances (the halo grows, the union
    stays exact) and gsc_group_rebalance re-cuts the slabs: migrants change owner, the counts even out again."""
    n, world = 40000, 4
    s = scenes.by_config(4, n)
    side = float(s.meta["side"])
    rng = np.random.default_rng(7)
    speed = rng.uniform(0.0, side / 60.0, size=n).astype(np.float32)     # up to a slab width in ~15 frames
    speed[rng.random(n) < 0.5] *= -1.0
    oracle = O.OracleSystem()
    cap = int(n / world * 3.0) + (1 << 16)
    ghosts_before = ghosts_after = 0
    with G.CollisionGroup(_devices(world), cap) as grp:
        grp.upload_colliders(s.entity, s.kind, s.offset, s.size, s.pos)
        for f in range(50):
            pos, quat = s.frame(f % 7)
            pos = pos.copy()
            pos[:, 0] += speed * f
            grp.update_transforms(pos, quat, s.scale)
            out = grp.step()
            if f % 5 == 0 or f == 49:
                want = oracle.step(s, pos, quat)
                assert out.n_pairs == len(want) and np.array_equal(grp.pairs_packed(sorted=True), want), f
            if f % 10 == 9:
                ghosts_before = out.n_ghost
                grp.rebalance(pos)
                counts = grp.member_counts()
                assert sum(counts) == n
                with pytest.raises(G.GscError):
                    grp.update_transforms(pos)                               # the new owners need rotation and scale again
                grp.update_transforms(pos, quat, s.scale)
                out2 = grp.step()                                            # same frame, new ownership: same set
                want = oracle.step(s, pos, quat)
                assert np.array_equal(grp.pairs_packed(sorted=True), want), f
                ghosts_after = out2.n_ghost
                assert ghosts_after < ghosts_before                            # the overlap of the drifted slabs is gone
                assert max(counts) < 0.5 * n
    oracle.close()


def test_pinned_staging_and_box_slots():
    """gsc_map_transform_staging / gsc_commit_transform_staging: the host marshals into the library's pinned banks"""
    s = scenes.by_config(3, 25000)
    oracle = O.OracleSystem()
    box = s.kind == 1
    with G.CollisionContext(s.n) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        slot = ctx.box_slots()
        assert np.array_equal(slot, np.concatenate([[0], np.cumsum(box)[:-1]]).astype(np.uint32))
        for f, bank in ((0, 0), (1, 1), (2, 0)):
            pos, quat = s.frame(f)
            p, q, sc = ctx.map_transform_staging(bank, ctx.cap)
            p[: s.n] = pos
            q[: ctx.n_box] = quat[box]
            sc[: ctx.n_box] = s.scale[box]
            ctx.commit_transform_staging(bank, ctx.n_box, True, f == 0)
            ctx.step()
            assert np.array_equal(ctx.pairs_packed(sorted=True), oracle.step(s, pos, quat))
    oracle.close()


def test_graph_replay_gives_the_same_pair_sets():
    """GSC_FLAG_GRAPH: stages 1-5 replayed from a captured CUDA graph; re-captured when the collider count changes"""
    s = scenes.by_config(4, 30000)
    oracle = O.OracleSystem()
    with G.CollisionContext(s.n, graph=True) as ctx:
        m = 20000
        ctx.upload_colliders(s.entity[:m], s.kind[:m], s.offset[:m], s.size[:m])
        sub = scenes.Scene("sub", s.entity[:m], s.kind[:m], s.offset[:m], s.size[:m], s.pos[:m], s.quat[:m], s.scale[:m])
        for f in (0, 1, 2):
            pos, quat = s.frame(f)
            ctx.update_transforms(pos[:m], quat[:m], s.scale[:m])
            ctx.step()
            assert np.array_equal(ctx.pairs_packed(sorted=True), oracle.step(sub, pos[:m], quat[:m])), f
        ctx.append_colliders(s.entity[m:], s.kind[m:], s.offset[m:], s.size[m:])      # new launch geometry
        for f in (3, 4):
            pos, quat = s.frame(f)
            ctx.update_transforms(pos, quat, s.scale)
            ctx.step()
            assert np.array_equal(ctx.pairs_packed(sorted=True), oracle.step(s, pos, quat)), f
    oracle.close()


def test_null_rotation_is_refused_after_the_collider_set_changed():
    """'NULL keeps the previous values' must not reuse rotation / scale arrays that belong to another
    collider set or layout (packed, or pre-swap_remove order)."""
    s = scenes.by_config(3, 4000)
    box = s.kind == 1
    oracle = O.OracleSystem()
    with G.CollisionContext(s.n) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        with pytest.raises(G.GscError):
            ctx.update_transforms(s.pos)                     # boxes, and rotation / scale were never sent
        ctx.update_transforms_packed(s.pos, s.quat[box], s.scale[box])
        ctx.step()
        ctx.remove_colliders(np.array([5, 17, 0], dtype=np.uint32))
        keep = O.swap_remove_sequence(np.arange(s.n, dtype=np.uint32), [5, 17, 0])
        with pytest.raises(G.GscError) as ei:
            ctx.update_transforms(s.pos[keep])               # NULL rotation after a remove: refused, not silently stale
        assert ei.value.code == N.GSC_ERR_INVALID
        with pytest.raises(G.GscError):
            ctx.update_transforms_packed(s.pos[keep])
        ctx.update_transforms(s.pos[keep], s.quat[keep], s.scale[keep])
        ctx.step()
        sub = scenes.Scene("sub", s.entity[keep], s.kind[keep], s.offset[keep], s.size[keep], s.pos[keep], s.quat[keep], s.scale[keep])
        assert np.array_equal(ctx.pairs_packed(sorted=True), oracle.step(sub))
        ctx.update_transforms(s.pos[keep])                    # now NULL keeps the values just sent
        ctx.step()
        assert np.array_equal(ctx.pairs_packed(sorted=True), oracle.step(sub))
    oracle.close()


def test_reordered_colliders_give_the_same_pair_sets():
    """gsc_reorder_colliders (temporal coherence): the dense arrays are permuted into the cell order of the
    last step and the host hands over transforms in slot order; every collider keeps its global index, so the pair set
    (orientation included) does not change."""
    s = scenes.by_config(4, 60000)
    rng = np.random.default_rng(5)
    s.entity = (rng.permutation(s.n) + 11).astype(np.uint32)      # ids unrelated to the index order
    oracle = O.OracleSystem()
    with G.CollisionContext(s.n) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        pos, quat = s.frame(0)
        ctx.update_transforms(pos, quat, s.scale)
        ctx.step()
        assert np.array_equal(ctx.pairs_packed(sorted=True), oracle.step(s, pos, quat))
        order = ctx.reorder_colliders()
        assert np.array_equal(np.sort(order), np.arange(s.n, dtype=np.uint32))
        assert np.array_equal(ctx.debug(N.GSC_DBG_ENTITY), s.entity[order])
        with pytest.raises(G.GscError):
            ctx.step()                                            # transforms must be handed over again, in slot order
        with pytest.raises(G.GscError):
            ctx.update_transforms(pos[order])                     # ... with rotation and scale
        for f in (0, 1, 4):
            pos, quat = s.frame(f)
            ctx.update_transforms(pos[order], quat[order], s.scale[order])
            out = ctx.step()
            assert np.array_equal(ctx.pairs_packed(sorted=True), oracle.step(s, pos, quat)), f
            # the colliders now arrive nearly cell-ordered: the sorted index stream is close to the identity
            idx = ctx.debug(N.GSC_DBG_SORTED_IDX).astype(np.int64)
            assert np.median(np.abs(idx - np.arange(s.n))) < s.n / 50
        # a second reorder composes
        order2 = ctx.reorder_colliders()
        total = order[order2]
        pos, quat = s.frame(5)
        ctx.update_transforms(pos[total], quat[total], s.scale[total])
        ctx.step()
        assert np.array_equal(ctx.pairs_packed(sorted=True), oracle.step(s, pos, quat))
        with pytest.raises(G.GscError):
            ctx.append_colliders(s.entity[:1], s.kind[:1], s.offset[:1], s.size[:1])
    oracle.close()


def test_wide_volume_on_a_slab_border_is_refused_not_dropped():
    """Volumes that rounding makes span 3 cells are matched by k_wide_pairs against LOCAL volumes only; in a multi-GPU
    step a wide volume on a slab border would silently lose its cross-slab pairs -- the step fails with
    GSC_ERR_RANGE instead.  (Single context: the corner-cell rule is followed exactly, test_gpu_parity.)"""
    rng = np.random.default_rng(0)
    n = 6000
    s = scenes.bench_scene(n)
    k = rng.integers(-6, 6, size=(n, 3)).astype(np.float32)
    pos = k + np.float32(0.5)
    big = np.arange(n) % 2 == 0
    for d in range(3):
        pos[big, d] = np.nextafter(pos[big, d], np.float32(-100), dtype=np.float32)
    s.pos = pos.astype(np.float32)
    s.size[~big, 0] = 0.2
    with G.CollisionGroup(_devices(2), n) as grp:
        grp.upload_colliders(s.entity, s.kind, s.offset, s.size, s.pos)
        grp.update_transforms(s.pos, s.quat, s.scale)
        with pytest.raises(G.GscError) as ei:
            grp.step()
        assert ei.value.code == N.GSC_ERR_RANGE and "slab border" in str(ei.value)

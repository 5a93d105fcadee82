"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.
Integer / index / pair-set work is compared bit-exact."""
import numpy as np
import pytest

import oracle_lib as O
from gunship_rs_b200 import _native as N
from gunship_rs_b200 import context as G
from gunship_rs_b200 import scenes

pytestmark = pytest.mark.gpu


def _rand_obb16(rng, n, spread=3.0):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    out = np.zeros((n, 16), dtype=np.float32)
    out[:, 0:3] = rng.uniform(-spread, spread, size=(n, 3))
    for i in range(n):
        out[i, 3:12] = O.quat_to_mat3(q[i].astype(np.float32)).reshape(9)
    out[:, 12:15] = rng.uniform(0.25, 2.0, size=(n, 3))
    return out


def test_narrowphase_batches_match_oracle():
    rng = np.random.default_rng(11)
    n = 20000
    a = np.concatenate([rng.uniform(-2, 2, size=(n, 3)), rng.uniform(0.3, 1.5, size=(n, 1))], axis=1).astype(np.float32)
    b = np.concatenate([rng.uniform(-2, 2, size=(n, 3)), rng.uniform(0.3, 1.5, size=(n, 1))], axis=1).astype(np.float32)
    # exact-touch and identity cases of the reference KATs (mod.rs:720-750)
    a[:4] = [[0, 0, 0, 1], [2, 2, 2, 1], [1.7, 1.7, 1.7, 0.3], [2, 0, 0, 1]]
    b[:4] = [[2, 0, 0, 1], [1.7, 1.7, 1.7, 0.3], [0, 0, 0, 1], [2, 0, 0, 1]]
    want = O.batch_sphere_sphere(a, b)
    assert 0.1 < want.mean() < 0.9
    assert np.array_equal(G.test_sphere_sphere(a, b), want)

    oa, ob = _rand_obb16(rng, n), _rand_obb16(rng, n)
    want = O.batch_obb_obb(oa, ob)
    assert 0.1 < want.mean() < 0.9
    assert np.array_equal(G.test_obb_obb(oa, ob), want)
    assert np.array_equal(G.test_obb_obb(ob, oa), O.batch_obb_obb(ob, oa))

    want = O.batch_sphere_obb(a, oa)
    assert 0.05 < want.mean() < 0.95
    assert np.array_equal(G.test_sphere_obb(a, oa), want)

    x = rng.uniform(-1, 1, size=(n, 6)).astype(np.float32)
    y = rng.uniform(-1, 1, size=(n, 6)).astype(np.float32)
    x[:, 3:] = x[:, :3] + np.abs(x[:, 3:])
    y[:, 3:] = y[:, :3] + np.abs(y[:, 3:])
    y[:10, :3] = x[:10, 3:]  # exact touch on every axis
    assert np.array_equal(G.test_aabb_aabb(x, y), O.batch_aabb_aabb(x, y))


def test_obb_kats_on_gpu():
    ident = np.eye(3, dtype=np.float32)
    unit = O.obb16([0, 0, 0], ident, [0.5, 0.5, 0.5])
    rot = O.mat3_rotation(0.0, 0.0, np.float32(0.25) * np.float32(np.pi))
    rot_z = O.obb16([1, 0, 0], rot, [0.5, 0.5, 0.5])
    far = O.obb16([1.21, 0, 0], rot, [0.5, 0.5, 0.5])
    a = np.stack([unit, unit, rot_z, unit, far])
    b = np.stack([unit, rot_z, unit, far, unit])
    assert np.array_equal(G.test_obb_obb(a, b), np.array([1, 1, 1, 0, 0], dtype=np.uint8))  # mod.rs:779-780


def test_fnv1a_gridcell_matches_oracle():
    rng = np.random.default_rng(5)
    cells = rng.integers(-32768, 32767, size=(5000, 3)).astype(np.int16)
    assert np.array_equal(G.fnv1a_gridcell(cells), O.batch_fnv1a_gridcell(cells))


@pytest.mark.parametrize("n,bits", [(1, 8), (255, 8), (4096, 13), (4097, 16), (100003, 22), (1 << 20, 24), (300000, 32)])
def test_radix_sort_is_stable_and_sorted(n, bits):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << bits, size=n, dtype=np.uint64).astype(np.uint32)
    if n > 1000:
        keys[: n // 3] = keys[0]  # heavy duplicates exercise stability
    vals = np.arange(n, dtype=np.uint32)
    k, v = G.sort_pairs_u32(keys, vals, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, vals[order])


def _run_scene(ctx, oracle, s, frames=(0,)):
    ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
    outs = []
    for f in frames:
        pos, quat = s.frame(f)
        ctx.update_transforms(pos, quat, s.scale)
        out = ctx.step()
        want = oracle.step(s, pos, quat)
        got = ctx.pairs_packed(sorted=True)
        assert out.n_pairs == len(want), (s.name, f, out.n_pairs, len(want))
        assert np.array_equal(got, want), (s.name, f)
        assert np.float32(out.cell_size) == np.float32(oracle.longest_axis)
        assert out.n_wide == 0 and oracle.stats()["wide"] == 0
        outs.append(out)
    return outs


@pytest.mark.parametrize("cfg,n", [(1, 1000), (2, 20000), (3, 20000), (4, 20000), (2, 300000), (3, 300000), (4, 300000)])
def test_step_pair_set_bit_exact(cfg, n):
    s = scenes.by_config(cfg, n)
    oracle = O.OracleSystem()
    with G.CollisionContext(n) as ctx:
        outs = _run_scene(ctx, oracle, s, frames=(0, 1, 5))
        if cfg >= 3:
            assert outs[0].n_sphere_obb > 0 and outs[0].n_obb_obb > 0
    oracle.close()


def test_stage_outputs_match_oracle():
    """AABBs, OBBs, cell size bit-for-bit; keys sorted; cell table consistent with the sorted keys."""
    s = scenes.by_config(3, 50000)
    oracle = O.OracleSystem()
    oracle.step(s)
    vol = oracle.volumes()
    with G.CollisionContext(s.n) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        ctx.update_transforms(s.pos, s.quat, s.scale)
        out = ctx.step()
        aabb = ctx.debug(N.GSC_DBG_AABB)
        assert np.array_equal(aabb[:, :3].view(np.uint32), vol["aabb_min"][:, :3].view(np.uint32))
        assert np.array_equal(aabb[:, 3:].view(np.uint32), vol["aabb_max"][:, :3].view(np.uint32))
        obb = ctx.debug(N.GSC_DBG_OBB)
        box = s.kind == 1
        # oracle payload of a box: centre xyzw, orientation 9, half 3
        pay = vol["payload"][box]
        assert np.array_equal(obb[box, 0:3].view(np.uint32), pay[:, 0:3].view(np.uint32))
        assert np.array_equal(obb[box, 3:12].view(np.uint32), pay[:, 4:13].view(np.uint32))
        assert np.array_equal(obb[box, 12:15].view(np.uint32), pay[:, 13:16].view(np.uint32))
        assert np.array_equal(obb[box, 15].view(np.uint32), s.entity[box])
        # cell keys: home cell = floor(aabb.min / cell) (grid_collision.rs:329-335), x major
        cell = np.float32(out.cell_size)
        home = np.floor(aabb[:, :3] / cell).astype(np.int64) - np.array(list(out.grid_origin), dtype=np.int64)
        dims = np.array(list(out.grid_dims), dtype=np.int64)
        want_keys = ((home[:, 0] * dims[1] + home[:, 1]) * dims[2] + home[:, 2]).astype(np.uint32)
        sk, si = ctx.debug(N.GSC_DBG_SORTED_KEYS), ctx.debug(N.GSC_DBG_SORTED_IDX)
        order = np.argsort(want_keys, kind="stable")
        assert np.array_equal(sk, want_keys[order]) and np.array_equal(si, order.astype(np.uint32))
        start = ctx.debug(N.GSC_DBG_CELL_START)
        assert np.array_equal(start, np.searchsorted(sk, np.arange(out.n_cells + 1), side="left").astype(np.uint32))
    oracle.close()


def test_wide_volumes_follow_corner_cell_rule():
    """Volumes that rounding makes span 3 cells are entered in their corner cells only
    (grid_collision.rs:447-480): the oracle then differs from brute force and the GPU must follow the oracle."""
    rng = np.random.default_rng(0)
    n = 6000
    s = scenes.bench_scene(n)
    k = rng.integers(-6, 6, size=(n, 3)).astype(np.float32)
    pos = k + np.float32(0.5)
    big = np.arange(n) % 2 == 0
    for d in range(3):
        pos[big, d] = np.nextafter(pos[big, d], np.float32(-100), dtype=np.float32)
    s.pos = pos.astype(np.float32)
    s.size[~big, 0] = 0.2  # small spheres strictly inside one cell
    oracle = O.OracleSystem()
    want = oracle.step(s)
    assert oracle.stats()["wide"] > 0
    assert not np.array_equal(want, oracle.brute_force_sorted())  # the rule matters in this scene
    with G.CollisionContext(n, max_pairs=len(want) + 1000) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        ctx.update_transforms(s.pos, s.quat, s.scale)
        out = ctx.step()
        assert out.n_wide > 0
        assert np.array_equal(ctx.pairs_packed(), want)
    oracle.close()


def test_entity_ids_need_not_follow_index_order():
    """Tuple orientation follows the index in BoundingVolumeManager.components, not the id (grid_collision.rs:440)."""
    s = scenes.by_config(3, 5000)
    rng = np.random.default_rng(3)
    s.entity = (rng.permutation(s.n) + 7).astype(np.uint32)
    oracle = O.OracleSystem()
    with G.CollisionContext(s.n) as ctx:
        _run_scene(ctx, oracle, s)
    oracle.close()


def test_empty_single_and_errors():
    with G.CollisionContext(16) as ctx:
        e = scenes.bench_scene(1)
        ctx.upload_colliders(e.entity[:0], e.kind[:0], e.offset[:0], e.size[:0])
        ctx.update_transforms(e.pos[:0], e.quat[:0], e.scale[:0])
        assert ctx.step().n_pairs == 0 and ctx.pairs().shape == (0, 2)
        ctx.upload_colliders(e.entity, e.kind, e.offset, e.size)
        ctx.update_transforms(e.pos, e.quat, e.scale)
        assert ctx.step().n_pairs == 0
        # Mesh -> GSC_ERR_KIND (unimplemented!() mod.rs:331)
        s = scenes.bench_scene(8)
        s.kind[2] = scenes.KIND_MESH
        with pytest.raises(G.GscError) as ei:
            ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        assert ei.value.code == N.GSC_ERR_KIND
        # NaN position -> GSC_ERR_RANGE
        s = scenes.bench_scene(8)
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        bad = s.pos.copy()
        bad[1, 0] = np.nan
        ctx.update_transforms(bad, s.quat, s.scale)
        with pytest.raises(G.GscError) as ei:
            ctx.step()
        assert ei.value.code == N.GSC_ERR_RANGE
        # cell coordinate outside i16 -> GSC_ERR_RANGE (grid_collision.rs:530-535)
        far = s.pos.copy()
        far[0, 0] = 1.0e6
        ctx.update_transforms(far, s.quat, s.scale)
        with pytest.raises(G.GscError) as ei:
            ctx.step()
        assert ei.value.code == N.GSC_ERR_RANGE
        # zero-size colliders -> cell size 0 -> GSC_ERR_RANGE
        z = scenes.bench_scene(8)
        z.size[:] = 0
        ctx.upload_colliders(z.entity, z.kind, z.offset, z.size)
        ctx.update_transforms(z.pos, z.quat, z.scale)
        with pytest.raises(G.GscError) as ei:
            ctx.step()
        assert ei.value.code == N.GSC_ERR_RANGE
        # too many colliders for the context
        big = scenes.bench_scene(17)
        with pytest.raises(G.GscError) as ei:
            ctx.upload_colliders(big.entity, big.kind, big.offset, big.size)
        assert ei.value.code == N.GSC_ERR_CAPACITY
    # pair buffer overflow reports the needed size
    s = scenes.bench_scene(1000)
    with G.CollisionContext(1000, max_pairs=100) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        ctx.update_transforms(s.pos, s.quat, s.scale)
        with pytest.raises(G.GscError) as ei:
            ctx.step()
        assert ei.value.code == N.GSC_ERR_CAPACITY and ctx.last.n_pairs > 10000


def test_large_scene_properties():
    """At a size the oracle would take long: size-independent properties of the pair set."""
    n = 1 << 21
    s = scenes.by_config(2, n)
    with G.CollisionContext(n, timing=True) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        ctx.update_transforms(s.pos, s.quat, s.scale)
        out = ctx.step()
        p = ctx.pairs(sorted=True)
        assert out.n_pairs == p.shape[0] > 0
        packed = (p[:, 0].astype(np.uint64) << np.uint64(32)) | p[:, 1]
        assert np.all(packed[1:] > packed[:-1])          # strictly ascending == sorted and duplicate free
        assert np.all(p[:, 0] > p[:, 1])                 # (later, earlier); ids are monotone in index here
        # every reported pair really collides (sphere-sphere, mod.rs:383-389, in f64 with a loose margin)
        a, b = p[:, 0].astype(np.int64) - 1, p[:, 1].astype(np.int64) - 1
        d = np.linalg.norm(s.pos[a].astype(np.float64) - s.pos[b].astype(np.float64), axis=1)
        assert np.all(d <= s.size[a, 0].astype(np.float64) + s.size[b, 0] + 1e-3)
        # idempotence: a second step over the same inputs gives the same set
        out2 = ctx.step()
        assert out2.n_pairs == out.n_pairs and np.array_equal(ctx.pairs(sorted=True), p)
        # a sample of colliders against brute force over all others
        rng = np.random.default_rng(1)
        sample = rng.choice(n, size=64, replace=False)
        for i in sample:
            dd = np.linalg.norm(s.pos.astype(np.float64) - s.pos[i].astype(np.float64), axis=1)
            rr = s.size[:, 0].astype(np.float64) + float(s.size[i, 0])
            sure = np.nonzero(dd < rr - 1e-3)[0]
            sure = sure[sure != i]
            mine = set(p[p[:, 0] == i + 1, 1].tolist()) | set(p[p[:, 1] == i + 1, 0].tolist())
            assert set((sure + 1).tolist()) <= mine


def test_packed_transforms_give_the_same_pair_set():
    """gsc_update_transforms_packed: rotation / scale for the boxes only (spheres ignore them, mod.rs:313-317)"""
    s = scenes.by_config(3, 30000)
    oracle = O.OracleSystem()
    box = s.kind == 1
    with G.CollisionContext(s.n) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        for f in (0, 2):
            pos, quat = s.frame(f)
            ctx.update_transforms_packed(pos, quat[box], s.scale[box] if f == 0 else None)
            ctx.step()
            assert np.array_equal(ctx.pairs_packed(sorted=True), oracle.step(s, pos, quat))
        # back to the dense form, then packed again
        pos, quat = s.frame(3)
        ctx.update_transforms(pos, quat, s.scale)
        ctx.step()
        assert np.array_equal(ctx.pairs_packed(sorted=True), oracle.step(s, pos, quat))
        with pytest.raises(G.GscError):
            ctx.update_transforms_packed(pos)          # the first packed update after a dense one needs rot + scale
        with pytest.raises(G.GscError):
            ctx.update_transforms_packed(pos, quat[box][:-1], s.scale[box][:-1])   # wrong box count
    oracle.close()


def test_box_aabb_without_the_vertex_walk_is_bit_exact_also_for_degenerate_boxes():
    """k_bvh_update computes a box AABB as centre -+ (|m0 hx| + |m1 hy|) + |m2 hz| per axis instead of walking the 8
    vertices (bounding_volume.rs:161-327); zero-extent axes take the literal path.  Bits must match the oracle's
    vertex walk, signs of zero included."""
    rng = np.random.default_rng(23)
    n = 40000
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[: n // 8] = [0, 0, 0, 1]                         # axis-aligned: exact zeros in the matrix
    q[n // 8: n // 4] = [np.sqrt(0.5), 0, 0, np.sqrt(0.5)]
    scale = rng.uniform(0.5, 3.5, size=(n, 3)).astype(np.float32)
    scale[rng.random((n, 3)) < 0.15] = 0.0              # flat / line / point boxes
    pos = rng.uniform(-40, 40, size=(n, 3)).astype(np.float32)
    pos[rng.random((n, 3)) < 0.1] = 0.0
    pos[rng.random((n, 3)) < 0.05] = -0.0
    pos[:, 0] += np.where(rng.random(n) < 0.05, np.float32(1e6), np.float32(0))  # large offsets: heavy rounding
    kind = np.ones(n, dtype=np.uint8)
    entity = np.arange(1, n + 1, dtype=np.uint32)
    offset = np.zeros((n, 3), dtype=np.float32)
    offset[rng.random((n, 3)) < 0.3] = -0.0
    size = np.ones((n, 3), dtype=np.float32)
    size[0] = 0.0
    sysm = O.OracleSystem(8, 4)
    sysm.bvh_update(entity, kind, offset, size, pos, q.astype(np.float32), scale)
    vol = sysm.volumes()
    with G.CollisionContext(n, max_cells=1 << 27) as ctx:
        ctx.upload_colliders(entity, kind, offset, size)
        ctx.update_transforms(pos, q.astype(np.float32), scale)
        try:
            ctx.step()
        except N.GscError as e:  # the grid may not fit (1e6 offsets): the AABBs were still produced by stage 0
            assert e.code in (N.GSC_ERR_CAPACITY, N.GSC_ERR_RANGE)
        aabb = ctx.debug(N.GSC_DBG_AABB)
        assert np.array_equal(aabb[:, :3].view(np.uint32), vol["aabb_min"][:, :3].view(np.uint32))
        assert np.array_equal(aabb[:, 3:].view(np.uint32), vol["aabb_max"][:, :3].view(np.uint32))
    sysm.close()

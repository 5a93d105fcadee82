"""Golden fixtures (tests/golden/collision_golden.npz, made by tests/golden/make_golden.py): the oracle must
reproduce them on CPU, the CUDA path must reproduce them on the GPU -- bit-exact (pair sets, AABBs, cell size,
narrowphase booleans, FNV hashes)."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
from gunship_rs_b200 import scenes

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as MG  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "collision_golden.npz"))


@pytest.mark.parametrize("cfg,n,frames", MG.SCENES)
def test_oracle_reproduces_golden(cfg, n, frames):
    s = scenes.by_config(cfg, n)
    sysm = O.OracleSystem()
    for f in frames:
        pos, quat = s.frame(f)
        assert np.array_equal(sysm.step(s, pos, quat), GOLD[f"pairs_c{cfg}_n{n}_f{f}"])
        assert np.float32(sysm.longest_axis) == GOLD[f"cell_c{cfg}_n{n}_f{f}"][0]
        v = sysm.volumes()
        aabb = np.concatenate([v["aabb_min"][:, :3], v["aabb_max"][:, :3]], axis=1)
        assert np.array_equal(aabb.view(np.uint32), GOLD[f"aabb_c{cfg}_n{n}_f{f}"].view(np.uint32))
    sysm.close()


def test_oracle_narrowphase_reproduces_golden():
    a, b, oa, ob = MG.narrowphase_inputs()
    assert np.array_equal(np.packbits(O.batch_sphere_sphere(a, b)), GOLD["np_sphere_sphere"])
    assert np.array_equal(np.packbits(O.batch_sphere_obb(a, oa)), GOLD["np_sphere_obb"])
    assert np.array_equal(np.packbits(O.batch_obb_obb(oa, ob)), GOLD["np_obb_obb"])
    assert np.array_equal(np.packbits(O.batch_obb_obb(ob, oa)), GOLD["np_obb_obb_swapped"])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,n,frames", MG.SCENES)
def test_cuda_step_reproduces_golden(cfg, n, frames):
    from gunship_rs_b200 import _native as N
    from gunship_rs_b200 import context as G
    s = scenes.by_config(cfg, n)
    with G.CollisionContext(n) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        for f in frames:
            pos, quat = s.frame(f)
            ctx.update_transforms(pos, quat, s.scale)
            out = ctx.step()
            assert np.array_equal(ctx.pairs_packed(sorted=True), GOLD[f"pairs_c{cfg}_n{n}_f{f}"])
            assert np.float32(out.cell_size) == GOLD[f"cell_c{cfg}_n{n}_f{f}"][0]
            assert np.array_equal(ctx.debug(N.GSC_DBG_AABB).view(np.uint32), GOLD[f"aabb_c{cfg}_n{n}_f{f}"].view(np.uint32))


@pytest.mark.gpu
def test_cuda_narrowphase_reproduces_golden():
    from gunship_rs_b200 import context as G
    a, b, oa, ob = MG.narrowphase_inputs()
    assert np.array_equal(np.packbits(G.test_sphere_sphere(a, b)), GOLD["np_sphere_sphere"])
    assert np.array_equal(np.packbits(G.test_sphere_obb(a, oa)), GOLD["np_sphere_obb"])
    assert np.array_equal(np.packbits(G.test_obb_obb(oa, ob)), GOLD["np_obb_obb"])
    assert np.array_equal(np.packbits(G.test_obb_obb(ob, oa)), GOLD["np_obb_obb_swapped"])
    cells = np.random.default_rng(9).integers(-32768, 32767, size=(512, 3)).astype(np.int16)
    assert np.array_equal(G.fnv1a_gridcell(cells), GOLD["fnv_gridcell"])


TGOLD = np.load(os.path.join(HERE, "golden", "transform_golden.npz"))


def test_oracle_transform_and_lifecycle_reproduce_golden():
    got = O.transform_update(TGOLD["row_begin"], TGOLD["parent"], TGOLD["lpos"], TGOLD["lquat"], TGOLD["lscale"])
    for g, name in zip(got, ("dpos", "dquat", "dscale", "dmat")):
        assert np.array_equal(g.view(np.uint32), TGOLD[name].view(np.uint32)), name
    after = O.swap_remove_sequence(np.arange(1, 501, dtype=np.uint32), TGOLD["remove_seq"])
    assert np.array_equal(after, TGOLD["order_after"])


@pytest.mark.gpu
def test_cuda_transform_and_lifecycle_reproduce_golden():
    from gunship_rs_b200 import _native as N
    from gunship_rs_b200 import context as G
    n_nodes = TGOLD["parent"].shape[0]
    n = 500
    with G.CollisionContext(n) as ctx:
        ent = np.arange(1, n + 1, dtype=np.uint32)
        ctx.upload_colliders(ent, np.zeros(n, np.uint8), np.zeros((n, 3), np.float32), np.ones((n, 3), np.float32))
        ctx.hierarchy_upload(TGOLD["row_begin"], TGOLD["parent"], (np.arange(n) * 3 % n_nodes).astype(np.uint32))
        ctx.hierarchy_update(TGOLD["lpos"], TGOLD["lquat"], TGOLD["lscale"])
        for g, name in zip(ctx.hierarchy_download(), ("dpos", "dquat", "dscale", "dmat")):
            assert np.array_equal(g.view(np.uint32), TGOLD[name].view(np.uint32)), name
        ctx.remove_colliders(TGOLD["remove_seq"])
        assert np.array_equal(ctx.debug(N.GSC_DBG_ENTITY), TGOLD["order_after"])

"""Contacts (unit normal lo->hi, depth).  The reference has none -- PARITY UNPINNED: the definition
lives in include/gunship_collision.h, the oracle restates it on the CPU and the CUDA path must agree with the
oracle within 1e-5 relative (north_star).  The CPU tests check the definition itself on analytic cases."""
import numpy as np
import pytest

import oracle_lib as O
from gunship_rs_b200 import scenes

RTOL = 1e-5


def _obb(center, quat, half):
    return O.obb16(center, O.quat_to_mat3(np.asarray(quat, dtype=np.float32)), half)


def test_oracle_sphere_sphere_contact_analytic():
    hi = np.array([[3.0, 0, 0, 1.0], [0, 0, 0, 1.0], [0, 2.0, 0, 1.5]], dtype=np.float32)
    lo = np.array([[0.0, 0, 0, 2.5], [0, 0, 0, 1.0], [0, 0, 0, 1.0]], dtype=np.float32)
    c = O.batch_contact_sphere_sphere(hi, lo)
    assert np.allclose(c[0], [1, 0, 0, 0.5]) and np.allclose(c[1], [1, 0, 0, 2.0]) and np.allclose(c[2], [0, 1, 0, 0.5])


def test_oracle_sphere_obb_contact_analytic():
    box = np.stack([_obb([0, 0, 0], [0, 0, 0, 1], [1, 2, 3])] * 3)
    s = np.array([[1.5, 0, 0, 1.0],      # outside the +x face by 0.5: depth 0.5, normal +x (box -> sphere)
                  [0.5, 0, 0, 0.25],     # centre inside: nearest face +x at distance 0.5 -> depth 0.75
                  [0, -2.5, 0, 1.0]], dtype=np.float32)
    hi = np.array([1, 1, 0], dtype=np.uint8)
    c = O.batch_contact_sphere_obb(s, box, hi)
    assert np.allclose(c[0], [1, 0, 0, 0.5]) and np.allclose(c[1], [1, 0, 0, 0.75])
    assert np.allclose(c[2], [0, 1, 0, 0.5])  # sphere is lo: normal flips to point sphere -> box


def test_oracle_obb_obb_contact_analytic():
    a = _obb([0, 0, 0], [0, 0, 0, 1], [1, 1, 1])[None]
    b = _obb([1.5, 0.2, 0.1], [0, 0, 0, 1], [1, 1, 1])[None]
    c = O.batch_contact_obb_obb(a, b)[0]   # hi = a at the origin, lo = b at +x: normal lo -> hi is -x, overlap 0.5
    assert np.allclose(c, [-1, 0, 0, 0.5], atol=2e-6)
    q = np.array([0, 0, np.sin(np.pi / 8), np.cos(np.pi / 8)], dtype=np.float32)   # 45 deg about z
    b2 = _obb([0, 2.2, 0], q, [1, 1, 1])[None]
    c2 = O.batch_contact_obb_obb(a, b2)[0]
    assert np.allclose(c2[:3], [0, -1, 0], atol=1e-5) and abs(c2[3] - (1 + np.sqrt(2) - 2.2)) < 1e-5
    assert abs(np.linalg.norm(c2[:3]) - 1) < 1e-5


def _rand_inputs(seed=3, n=20000):
    rng = np.random.default_rng(seed)
    sph = lambda: np.concatenate([rng.uniform(-1.5, 1.5, (n, 3)), rng.uniform(0.3, 1.5, (n, 1))], axis=1).astype(np.float32)

    def obb():
        q = rng.normal(size=(n, 4))
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        out = np.zeros((n, 16), dtype=np.float32)
        out[:, 0:3] = rng.uniform(-2, 2, (n, 3))
        for i in range(n):
            out[i, 3:12] = O.quat_to_mat3(q[i].astype(np.float32)).reshape(9)
        out[:, 12:15] = rng.uniform(0.25, 2.0, (n, 3))
        return out
    return sph(), sph(), obb(), obb(), (rng.uniform(size=n) < 0.5).astype(np.uint8)


def test_oracle_contacts_are_unit_and_separate():
    """moving hi by depth * normal (plus a hair) separates every colliding random pair along the chosen axis"""
    a, b, oa, ob, flag = _rand_inputs(n=4000)
    hit = O.batch_sphere_sphere(a, b) != 0
    c = O.batch_contact_sphere_sphere(a, b)[hit]
    assert np.allclose(np.linalg.norm(c[:, :3], axis=1), 1, atol=1e-5) and np.all(c[:, 3] >= -1e-5)
    moved = a[hit].copy()
    moved[:, :3] += c[:, :3] * (c[:, 3:4] + 1e-3)
    assert not O.batch_sphere_sphere(moved, b[hit]).any()
    hit = O.batch_obb_obb(oa, ob) != 0
    c = O.batch_contact_obb_obb(oa, ob)[hit]
    assert np.allclose(np.linalg.norm(c[:, :3], axis=1), 1, atol=1e-5) and np.all(c[:, 3] >= -1e-5)
    moved = oa[hit].copy()
    moved[:, :3] += c[:, :3] * (c[:, 3:4] + 1e-3)
    assert not O.batch_obb_obb(moved, ob[hit]).any()
    hit = O.batch_sphere_obb(a, oa) != 0
    c = O.batch_contact_sphere_obb(a, oa, np.ones(len(a), np.uint8))[hit]
    moved = a[hit].copy()
    moved[:, :3] += c[:, :3] * (c[:, 3:4] + 1e-3)
    assert not O.batch_sphere_obb(moved, oa[hit]).any()


def _close(got, want):
    return np.allclose(got, want, rtol=RTOL, atol=RTOL)


@pytest.mark.gpu
def test_cuda_contact_batches_match_oracle():
    from gunship_rs_b200 import context as G
    a, b, oa, ob, flag = _rand_inputs()
    assert _close(G.contact_sphere_sphere(a, b), O.batch_contact_sphere_sphere(a, b))
    assert _close(G.contact_sphere_obb(a, oa, flag), O.batch_contact_sphere_obb(a, oa, flag))
    got, want = G.contact_obb_obb(oa, ob), O.batch_contact_obb_obb(oa, ob)
    assert _close(got, want)
    # same operation order on both sides: in fact bit-identical
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,n", [(1, 1000), (3, 40000), (4, 40000)])
def test_cuda_step_contacts_match_oracle(cfg, n):
    """contacts of a whole step: pair i of the sorted pair list carries the oracle's contact of that pair"""
    from gunship_rs_b200 import context as G
    s = scenes.by_config(cfg, n)
    oracle = O.OracleSystem()
    pos, quat = s.frame(1)
    want_pairs = oracle.step(s, pos, quat)
    vol = oracle.volumes()
    with G.CollisionContext(n, contacts=True) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        ctx.update_transforms(pos, quat, s.scale)
        ctx.step()
        assert np.array_equal(ctx.pairs_packed(sorted=True), want_pairs)
        got = ctx.contacts(sorted=True)
    # entity -> index (ids are 1..n in index order in these scenes)
    hi = (want_pairs >> np.uint64(32)).astype(np.int64) - 1
    lo = (want_pairs & np.uint64(0xFFFFFFFF)).astype(np.int64) - 1
    pay, tag = vol["payload"], vol["tag"]
    sph = lambda i: np.ascontiguousarray(np.concatenate([pay[i, 0:3], pay[i, 4:5]], axis=1))
    obb = lambda i: np.ascontiguousarray(np.concatenate([pay[i, 0:3], pay[i, 4:16], np.zeros((len(i), 1), np.float32)], axis=1))
    want = np.zeros((len(want_pairs), 4), dtype=np.float32)
    hs, ls = tag[hi] == 0, tag[lo] == 0
    m = hs & ls
    want[m] = O.batch_contact_sphere_sphere(sph(hi[m]), sph(lo[m]))
    m = hs & ~ls
    want[m] = O.batch_contact_sphere_obb(sph(hi[m]), obb(lo[m]), np.ones(m.sum(), np.uint8))
    m = ~hs & ls
    want[m] = O.batch_contact_sphere_obb(sph(lo[m]), obb(hi[m]), np.zeros(m.sum(), np.uint8))
    m = ~hs & ~ls
    want[m] = O.batch_contact_obb_obb(obb(hi[m]), obb(lo[m]))
    assert got.shape == want.shape and _close(got, want)
    assert np.allclose(np.linalg.norm(got[:, :3], axis=1), 1, atol=1e-4)
    oracle.close()

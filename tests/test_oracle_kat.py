"""Pins the CPU oracle against every known-answer test the reference holds for the collision path
 and against the brute-force second oracle.  CPU only."""
import math

import numpy as np
import pytest

import oracle_lib as O
from gunship_rs_b200 import scenes


def _sph(x, y, z, r):
    return np.array([x, y, z, r], dtype=np.float32)


# src/component/collider/mod.rs:720-750 (sphere_sphere_tests)
UNIT = _sph(0, 0, 0, 1.0)
FAR_UNIT = _sph(2, 2, 2, 1.0)
ODD_RADIUS = _sph(1.7, 1.7, 1.7, 0.3)
UNIT_EDGE = _sph(2, 0, 0, 1.0)
SPHERE_KATS = [
    (UNIT, UNIT, 1), (FAR_UNIT, FAR_UNIT, 1), (ODD_RADIUS, ODD_RADIUS, 1), (UNIT_EDGE, UNIT_EDGE, 1),
    (UNIT, FAR_UNIT, 0), (UNIT, ODD_RADIUS, 0), (FAR_UNIT, ODD_RADIUS, 1), (UNIT, UNIT_EDGE, 1),
]


def test_sphere_sphere_kats_both_orders():
    a = np.stack([k[0] for k in SPHERE_KATS])
    b = np.stack([k[1] for k in SPHERE_KATS])
    want = np.array([k[2] for k in SPHERE_KATS], dtype=np.uint8)
    assert np.array_equal(O.batch_sphere_sphere(a, b), want)
    assert np.array_equal(O.batch_sphere_sphere(b, a), want)  # macro asserts both orders (mod.rs:713-718)


def _obb_kats():
    # mod.rs:768-781 (obb_obb_tests)
    ident = np.eye(3, dtype=np.float32)
    unit = O.obb16([0, 0, 0], ident, [0.5, 0.5, 0.5])
    rot = O.mat3_rotation(0.0, 0.0, np.float32(0.25) * np.float32(math.pi))
    rot_z = O.obb16([1, 0, 0], rot, [0.5, 0.5, 0.5])
    return unit, rot_z, rot


def test_obb_obb_kats_both_orders():
    unit, rot_z, _ = _obb_kats()
    a = np.stack([unit, unit, rot_z])
    b = np.stack([unit, rot_z, unit])
    assert np.array_equal(O.batch_obb_obb(a, b), np.array([1, 1, 1], dtype=np.uint8))


def test_obb_obb_extra_negatives():
    # not from the reference (re-derived): centre 1.2 -> hit, 1.21 / 3.0 -> miss
    unit, _, rot = _obb_kats()
    cases = [(1.2, 1), (1.21, 0), (3.0, 0)]
    a = np.stack([unit] * len(cases))
    b = np.stack([O.obb16([c, 0, 0], rot, [0.5, 0.5, 0.5]) for c, _ in cases])
    want = np.array([w for _, w in cases], dtype=np.uint8)
    assert np.array_equal(O.batch_obb_obb(a, b), want)
    assert np.array_equal(O.batch_obb_obb(b, a), want)


def _axis_angle_quat(axis, angle):
    # orientation.rs:35-42: Quaternion::new(axis * sin(angle/2), cos(angle/2)).normalized()
    h = np.float32(angle) * np.float32(0.5)
    q = np.zeros(4, dtype=np.float64)
    q[axis] = math.sin(h)
    q[3] = math.cos(h)
    q /= np.linalg.norm(q)
    return q.astype(np.float32)


@pytest.mark.parametrize("angle", [math.pi, math.pi * 0.5, 0.5])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_quaternion_to_matrix_kats(axis, angle):
    # lib/polygon_math/src/test/quaternion_test.rs:20-32: axis_angle(..).as_matrix() == rotation(..), eps 1e-6
    # (Matrix4 PartialEq is an is_zero() compare per element, matrix.rs)
    e = [0.0, 0.0, 0.0]
    e[axis] = angle
    want = O.mat3_rotation(np.float32(e[0]), np.float32(e[1]), np.float32(e[2]))
    got = O.quat_to_mat3(_axis_angle_quat(axis, angle))
    assert np.all(np.abs(got - want) < 1e-6 + 1e-7), (got, want)


def test_quaternion_identity():
    assert np.array_equal(O.quat_to_mat3([0, 0, 0, 1]), np.eye(3, dtype=np.float32))


def test_fnv1a_known_vectors():
    # FNV-1a 64 published test vectors: "" -> basis, "a" -> af63dc4c8601ec8c
    L = O.lib()
    import ctypes as C
    L.orc_fnv1a.restype = C.c_uint64
    L.orc_fnv1a.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64]
    basis = 0xcbf29ce484222325
    assert L.orc_fnv1a(b"", 0, basis) == basis
    assert L.orc_fnv1a(b"a", 1, basis) == 0xaf63dc4c8601ec8c
    assert L.orc_fnv1a(b"foobar", 6, basis) == 0x85944171f73967e8
    # GridCell hashes its three i16 as little-endian bytes (grid_collision.rs:523-528)
    cell = np.array([[1, -2, 300]], dtype=np.int16)
    assert int(O.batch_fnv1a_gridcell(cell)[0]) == L.orc_fnv1a(cell.tobytes(), 6, basis)


def test_aabb_inclusive_touch():
    # bounding_volume.rs:411-419: touching counts as overlapping
    a = np.array([[0, 0, 0, 1, 1, 1]], dtype=np.float32)
    b = np.array([[1, 1, 1, 2, 2, 2]], dtype=np.float32)
    c = np.array([[1.0000001, 0, 0, 2, 1, 1]], dtype=np.float32)
    assert O.batch_aabb_aabb(a, b)[0] == 1
    assert O.batch_aabb_aabb(a, c)[0] == 0


def test_sphere_obb_strict():
    # mod.rs:393: dist_sqr < r*r, strict -- exact touch is NOT a hit (unlike sphere-sphere)
    box = O.obb16([0, 0, 0], np.eye(3), [0.5, 0.5, 0.5])[None]
    touch = np.array([[1.5, 0, 0, 1.0]], dtype=np.float32)
    inside = np.array([[1.4990, 0, 0, 1.0]], dtype=np.float32)
    assert O.batch_sphere_obb(touch, box)[0] == 0
    assert O.batch_sphere_obb(inside, box)[0] == 1


def test_bench_scene_matches_survey_count():
    # ~14.4k hits for 1000 r=0.5 spheres in 10x10, cell ~= 1.0000002
    s = scenes.bench_scene()
    sysm = O.OracleSystem()
    pairs = sysm.step(s)
    assert 13000 < len(pairs) < 16000
    assert abs(sysm.longest_axis - 1.0) < 1e-6
    st = sysm.stats()
    assert st["wide"] == 0 and st["out_of_range"] == 0
    # tuple orientation: (later index, earlier index) == (larger id, smaller id) here
    assert np.all((pairs >> np.uint64(32)) > (pairs & np.uint64(0xFFFFFFFF)))
    sysm.close()


@pytest.mark.parametrize("cfg,n", [(1, 1000), (2, 6000), (3, 6000), (4, 6000)])
@pytest.mark.parametrize("units,workers", [(8, 8), (8, 3), (2, 2), (4, 1)])
def test_grid_equals_brute_force(cfg, n, units, workers):
    """grid + octant pipeline == {(i,j): i>j, aabb & narrowphase} for any work-unit split
    that covers space (1 unit covers only the negative octant in the reference, so it is excluded)."""
    s = scenes.by_config(cfg, n)
    sysm = O.OracleSystem(units, workers)
    for f in (0, 2):
        pos, quat = s.frame(f)
        got = sysm.step(s, pos, quat)
        want = sysm.brute_force_sorted()
        assert sysm.stats()["wide"] == 0
        assert np.array_equal(got, want)
    sysm.close()


def test_volume_layout_and_aabb():
    s = scenes.by_config(3, 500)
    sysm = O.OracleSystem()
    sysm.step(s)
    v = sysm.volumes()
    assert v.shape[0] == 500 and np.array_equal(v["entity"], s.entity)
    sph = s.kind == 0
    # sphere AABB = centre -/+ r (bounding_volume.rs:150-159)
    c = (s.pos + s.offset)[sph]
    r = s.size[sph, 0:1]
    assert np.array_equal(v["aabb_min"][sph, :3], c - r)
    assert np.array_equal(v["aabb_max"][sph, :3], c + r)
    # box AABB contains its centre and is no larger than the diagonal
    box = ~sph
    assert np.all(v["aabb_min"][box, :3] <= s.pos[box]) and np.all(v["aabb_max"][box, :3] >= s.pos[box])
    ext = (v["aabb_max"][:, :3] - v["aabb_min"][:, :3]).max()
    assert sysm.longest_axis == ext
    sysm.close()


def test_mesh_is_rejected():
    s = scenes.bench_scene(10)
    s.kind[3] = scenes.KIND_MESH
    sysm = O.OracleSystem()
    with pytest.raises(NotImplementedError):
        sysm.step(s)
    sysm.close()


def test_empty_and_single():
    sysm = O.OracleSystem()
    s = scenes.bench_scene(1)
    assert len(sysm.step(s)) == 0
    e = scenes.bench_scene(1)
    for name in ("entity", "kind", "offset", "size", "pos", "quat", "scale"):
        setattr(e, name, getattr(e, name)[:0])
    assert len(sysm.step(e)) == 0
    sysm.close()

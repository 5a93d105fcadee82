"""The compiled host side above the C ABI (host/gunship_collision.hpp): the C++ mirror of the reference's collider
interface -- Collider / ColliderManager / CollisionSystem / GridCollisionSystem.collisions / callbacks / destroy --
built with g++ against libgunship_collision.so.  On the CPU: it compiles, links and fails loudly without a device.
On the GPU: the reference's own unit tests (mod.rs:711-782) and a three-frame scene with destroy / late assign /
callbacks, compared with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from gunship_rs_b200 import _native as N
from gunship_rs_b200 import scenes
from gunship_rs_b200.colliders import ColliderSet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, name, src_dir=("tests", "cpp")):
    exe = tmp_path / name
    libdir = os.path.join(ROOT, "gunship_rs_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "host"), os.path.join(ROOT, *src_dir, name + ".cpp"), "-o", str(exe),
                           "-L", libdir, "-lgunship_collision", f"-Wl,-rpath,{libdir}"])
    return str(exe)


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_host_mirror_builds_and_panics_without_a_device(tmp_path):
    kat = _build(tmp_path, "collider_kat")
    _build(tmp_path, "scene_runner")
    _build(tmp_path, "group_runner")
    _build(tmp_path, "grid_collision", ("benches",))  # the C++ counterpart of benches/grid_collision.rs
    if _has_gpu():
        pytest.skip("a CUDA device is present: the GPU tests run the binaries")
    r = subprocess.run([kat], capture_output=True, text=True)
    assert r.returncode == 3 and "PANIC" in r.stdout  # gunship::Panic, as the reference would abort; no CPU fallback


@pytest.mark.gpu
def test_reference_unit_tests_through_the_cpp_host(tmp_path):
    r = subprocess.run([_build(tmp_path, "collider_kat")], capture_output=True, text=True)
    assert r.returncode == 0 and "collider_kat: ok" in r.stdout, r.stdout + r.stderr


class _NoDevice:
    def upload_colliders(self, *a): pass
    def append_colliders(self, *a): pass
    def remove_colliders(self, *a): pass


@pytest.mark.gpu
def test_scene_through_the_cpp_host_matches_the_oracle(tmp_path):
    n, n_late, n_frames, stride = 3000, 300, 3, 7
    s = scenes.by_config(3, n + n_late)
    frames = [s.frame(f) for f in range(n_frames)]
    scene_file, out_file = tmp_path / "scene.bin", tmp_path / "out.bin"
    with open(scene_file, "wb") as f:
        f.write(struct.pack("<4I", n, n_late, n_frames, stride))
        for i in range(n + n_late):
            f.write(struct.pack("<2I6f", int(s.entity[i]), int(s.kind[i]), *s.offset[i], *s.size[i]))
        for pos, quat in frames:
            f.write(np.concatenate([pos, quat, s.scale], axis=1).astype("<f4").tobytes())
    r = subprocess.run([_build(tmp_path, "scene_runner"), str(scene_file), str(out_file)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr

    # what the reference would have produced: the oracle over the dense order as assign / destroy evolve it
    row = {int(e): i for i, e in enumerate(s.entity)}
    cs = ColliderSet(_NoDevice())
    for i in range(n):
        cs.assign(int(s.entity[i]), int(s.kind[i]), s.offset[i], s.size[i])
    has_cb = {int(s.entity[i]) for i in range(0, n, 3)}
    sysm = O.OracleSystem(8, 4)
    want_pairs, calls, others, checksum = [], 0, 0, 0
    for fr in range(n_frames):
        if fr == 1:
            for i in range(0, n, stride):
                cs.destroy(int(s.entity[i]))
            for i in range(n, n + n_late):
                cs.assign(int(s.entity[i]), int(s.kind[i]), s.offset[i], s.size[i])
        ent, kind, off, size = cs.arrays()
        idx = np.array([row[int(e)] for e in ent])
        pos, quat = frames[fr]
        sysm.bvh_update(ent, kind, off, size, pos[idx], quat[idx], s.scale[idx])
        sysm.grid_update()
        pairs = sysm.collisions_sorted()
        want_pairs.append(pairs)
        e, o_, oth = sysm.process_collisions()
        for k in range(e.shape[0]):
            if int(e[k]) in has_cb:
                lst = oth[o_[k]:o_[k + 1]].astype(np.uint64)
                calls += 1
                others += int(lst.shape[0])
                checksum = (checksum + int(e[k]) * 1000003 * int(lst.shape[0]) + int(lst.sum())) % (1 << 64)
        for ent_removed in [int(x) for x in cs._marked]:
            has_cb.discard(ent_removed)  # unregister_all at the cleanup (mod.rs:606)
        cs.cleanup()
    sysm.close()

    raw = open(out_file, "rb").read()
    at = 0
    for fr in range(n_frames):
        (np_,) = struct.unpack_from("<Q", raw, at)
        at += 8
        got = np.frombuffer(raw, dtype="<u8", count=np_, offset=at)
        at += 8 * np_
        assert want_pairs[fr].shape[0] > 100
        assert np.array_equal(got, want_pairs[fr]), f"frame {fr}"
    got_calls, got_others, got_sum = struct.unpack_from("<3Q", raw, at)
    assert (got_calls, got_others, got_sum) == (calls, others, checksum)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_group_of_gpus_through_the_cpp_host_matches_the_oracle(tmp_path, world):
    """gunship::GroupGridCollisionSystem (gsc_group_* behind the compiled host mirror): the union of the slabs' lists is
    the oracle's set on every frame, also after a re-balance."""
    import torch
    n, n_frames = 50000, 4
    s = scenes.by_config(4, n)
    frames = [s.frame(f) for f in range(n_frames)]
    scene_file, out_file = tmp_path / "gscene.bin", tmp_path / "gout.bin"
    with open(scene_file, "wb") as f:
        f.write(struct.pack("<2I", n, n_frames))
        for i in range(n):
            f.write(struct.pack("<2I6f", int(s.entity[i]), int(s.kind[i]), *s.offset[i], *s.size[i]))
        for pos, quat in frames:
            f.write(np.concatenate([pos, quat, s.scale], axis=1).astype("<f4").tobytes())
    devices = [str(i % torch.cuda.device_count()) for i in range(world)]
    r = subprocess.run([_build(tmp_path, "group_runner"), str(scene_file), str(out_file)] + devices, capture_output=True, text=True)
    assert r.returncode == 0 and "group_runner: ok" in r.stdout, r.stdout + r.stderr
    raw = open(out_file, "rb").read()
    at = 0
    oracle = O.OracleSystem()
    for fr in range(n_frames):
        np_, ghosts = struct.unpack_from("<2Q", raw, at)
        at += 16
        got = np.frombuffer(raw, dtype="<u8", count=np_, offset=at)
        at += 8 * np_
        want = oracle.step(s, *frames[fr])
        assert np.array_equal(got, want), f"frame {fr}"
        assert 0 < ghosts < n
    oracle.close()

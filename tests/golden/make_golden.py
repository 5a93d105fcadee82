#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ from the CPU oracle (oracle/gunship_oracle.c).

    python tests/golden/make_golden.py

The reference has no fixtures for this path and cannot be run here, so these are snapshots
of the oracle -- taken once, committed -- that pin BOTH the oracle and the CUDA path against silent drift:
tests/test_golden.py checks the oracle against them on CPU and the CUDA path against them on the GPU.
Inputs are not stored: they come from the seeded generators in gunship_rs_b200/scenes.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O  # noqa: E402
from gunship_rs_b200 import scenes  # noqa: E402

# (config, n, frames)
SCENES = [(1, 1000, (0,)), (2, 5000, (0, 3)), (3, 5000, (0, 3)), (4, 5000, (0, 3))]


def narrowphase_inputs(seed=7, n=4096):
    rng = np.random.default_rng(seed)
    sph = lambda: np.concatenate([rng.uniform(-2, 2, (n, 3)), rng.uniform(0.3, 1.5, (n, 1))], axis=1).astype(np.float32)
    a, b = sph(), sph()

    def obb():
        q = rng.normal(size=(n, 4))
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        out = np.zeros((n, 16), dtype=np.float32)
        out[:, 0:3] = rng.uniform(-3, 3, (n, 3))
        for i in range(n):
            out[i, 3:12] = O.quat_to_mat3(q[i].astype(np.float32)).reshape(9)
        out[:, 12:15] = rng.uniform(0.25, 2.0, (n, 3))
        return out
    return a, b, obb(), obb()


def main():
    out = {}
    sysm = O.OracleSystem()
    for cfg, n, frames in SCENES:
        s = scenes.by_config(cfg, n)
        for f in frames:
            pos, quat = s.frame(f)
            pairs = sysm.step(s, pos, quat)
            out[f"pairs_c{cfg}_n{n}_f{f}"] = pairs
            out[f"cell_c{cfg}_n{n}_f{f}"] = np.array([sysm.longest_axis], dtype=np.float32)
            v = sysm.volumes()
            out[f"aabb_c{cfg}_n{n}_f{f}"] = np.concatenate([v["aabb_min"][:, :3], v["aabb_max"][:, :3]], axis=1)
    sysm.close()
    a, b, oa, ob = narrowphase_inputs()
    out["np_sphere_sphere"] = np.packbits(O.batch_sphere_sphere(a, b))
    out["np_sphere_obb"] = np.packbits(O.batch_sphere_obb(a, oa))
    out["np_obb_obb"] = np.packbits(O.batch_obb_obb(oa, ob))
    out["np_obb_obb_swapped"] = np.packbits(O.batch_obb_obb(ob, oa))
    cells = np.random.default_rng(9).integers(-32768, 32767, size=(512, 3)).astype(np.int16)
    out["fnv_gridcell"] = O.batch_fnv1a_gridcell(cells)
    path = os.path.join(HERE, "collision_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if k.startswith("pairs")})


if __name__ == "__main__":
    main()

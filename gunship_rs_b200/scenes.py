"""Deterministic synthetic collider fields (host side, numpy only).

The five workloads of BASELINE.json `configs`.  Scenes are generated once on the
host from a counter-based PRNG (SplitMix64) so the CPU oracle and the CUDA path are fed identical
bytes.  Shapes follow the reference's own scenes:

* config 1 -- benches/grid_collision.rs:69-81: 1000 spheres r=0.5, x,y = u*10-5, z=0.
* configs 2-5 -- examples/benchmarks/circle_collision.rs:83-127: boxes `widths=(1,1,1)` with
  scale U[0.5,3.5)^3 and free rotation, spheres r U[0.5,1.0); per-frame motion as in :209-271.

Nothing here runs on the GPU; it only produces inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

KIND_SPHERE = 0  # Collider::Sphere  (mod.rs:153-159)
KIND_BOX = 1     # Collider::Box     (mod.rs:176-179)
KIND_MESH = 2    # Collider::Mesh    (mod.rs:182) -- unimplemented!() in the reference

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(z: np.ndarray) -> np.ndarray:
    """SplitMix64 finaliser on a uint64 array (wrapping arithmetic)."""
    z = z.astype(np.uint64, copy=True)
    z ^= z >> np.uint64(30)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


def u01(seed: int, stream: int, n: int, start: int = 0) -> np.ndarray:
    """n float64 values in [0,1) with 24 significant bits (exactly representable in f32)."""
    with np.errstate(over="ignore"):
        base = _mix(np.array([(seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF],
                             dtype=np.uint64))[0]
        ctr = np.arange(start, start + n, dtype=np.uint64)
        h = _mix(base + (ctr + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
    return (h >> np.uint64(40)).astype(np.float64) * (1.0 / (1 << 24))


def _gauss(seed: int, stream: int, n: int) -> np.ndarray:
    a = u01(seed, stream, n)
    b = u01(seed, stream + 1, n)
    return np.sqrt(-2.0 * np.log(1.0 - a)) * np.cos(2.0 * np.pi * b)


@dataclass
class Scene:
    """SoA collider field: static collider part + root transforms (derived == local)."""
    name: str
    entity: np.ndarray   # u32[N], ids from 1 (src/old/ecs.rs:32,40-41)
    kind: np.ndarray     # u8[N]
    offset: np.ndarray   # f32[N,3]  Collider::{Sphere,Box}.offset
    size: np.ndarray     # f32[N,3]  sphere: [radius,0,0]; box: widths
    pos: np.ndarray      # f32[N,3]  position_derived
    quat: np.ndarray     # f32[N,4]  rotation_derived as (x,y,z,w)
    scale: np.ndarray    # f32[N,3]  scale_derived
    cell_hint: float = 0.0  # the cell size the generator aimed for (the step recomputes the real one)
    meta: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return int(self.entity.shape[0])

    def frame(self, f: int, dt: float = 1.0 / 60.0):
        """Transforms of frame f (f=0 is the base pose): per-frame jitter of SURVEY §8d config 3/4.

        position = base + U[-0.1,0.1)*cell per axis (fresh per frame);
        rotation = base * q(euler rates * f * dt), rates U[0.01,0.15)*pi per axis
        (circle_collision.rs:104-107, 263-271).  Returns (pos, quat) as new f32 arrays.
        """
        if f == 0:
            return self.pos.copy(), self.quat.copy()
        seed = int(self.meta.get("seed", 0))
        n = self.n
        cell = float(self.cell_hint)
        jit = np.stack([u01(seed, 1000 + 3 * f + a, n) for a in range(3)], axis=1)
        pos = (self.pos.astype(np.float64) + (jit * 0.2 - 0.1) * cell).astype(np.float32)
        rates = np.stack([u01(seed, 900 + a, n) for a in range(3)], axis=1) * (0.14 * np.pi) + 0.01 * np.pi
        ang = rates * (f * dt)
        q = self.quat.astype(np.float64)
        for a in range(3):  # from_eulers = axis_angle(x) + axis_angle(y) + axis_angle(z) (orientation.rs:76-80)
            h = 0.5 * ang[:, a]
            dq = np.zeros((n, 4))
            dq[:, a] = np.sin(h)
            dq[:, 3] = np.cos(h)
            q = _quat_mul(q, dq)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        return pos, q.astype(np.float32)


def _quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Hamilton product, (x,y,z,w) layout (quaternion.rs:161-170)."""
    av, aw = a[:, :3], a[:, 3:4]
    bv, bw = b[:, :3], b[:, 3:4]
    v = np.cross(av, bv) + bw * av + aw * bv
    w = aw * bw - np.sum(av * bv, axis=1, keepdims=True)
    return np.concatenate([v, w], axis=1)


def _ids(n: int) -> np.ndarray:
    return np.arange(1, n + 1, dtype=np.uint32)


def _identity_quat(n: int) -> np.ndarray:
    q = np.zeros((n, 4), dtype=np.float32)
    q[:, 3] = 1.0
    return q


def bench_scene(n: int = 1000, seed: int = 1) -> Scene:
    """Config 1: the scene of benches/grid_collision.rs:69-81 (the reference is unseeded)."""
    x = (u01(seed, 0, n).astype(np.float32) * np.float32(10.0) - np.float32(5.0)).astype(np.float32)
    y = (u01(seed, 1, n).astype(np.float32) * np.float32(10.0) - np.float32(5.0)).astype(np.float32)
    pos = np.stack([x, y, np.zeros(n, dtype=np.float32)], axis=1)
    size = np.zeros((n, 3), dtype=np.float32)
    size[:, 0] = 0.5
    return Scene("bench_1k_spheres", _ids(n), np.zeros(n, dtype=np.uint8), np.zeros((n, 3), dtype=np.float32),
                 size, pos, _identity_quat(n), np.ones((n, 3), dtype=np.float32), cell_hint=1.0,
                 meta={"seed": seed, "config": 1})


def _sphere_radii(seed: int, n: int) -> np.ndarray:
    return (0.5 + 0.5 * u01(seed, 10, n)).astype(np.float32)  # circle_collision.rs:126


def sphere_field(n: int = 1 << 20, seed: int = 2, rho: float = 0.5) -> Scene:
    """Configs 2 and 5: uniform-random spheres, cell = 2*max r, rho colliders per cell volume."""
    r = _sphere_radii(seed, n)
    cell = 2.0  # 2 * sup(r)
    side = cell * (n / rho) ** (1.0 / 3.0)
    pos = np.stack([(u01(seed, a, n) - 0.5) * side for a in range(3)], axis=1).astype(np.float32)
    size = np.zeros((n, 3), dtype=np.float32)
    size[:, 0] = r
    return Scene(f"spheres_{n}", _ids(n), np.zeros(n, dtype=np.uint8), np.zeros((n, 3), dtype=np.float32),
                 size, pos, _identity_quat(n), np.ones((n, 3), dtype=np.float32), cell_hint=cell,
                 meta={"seed": seed, "rho": rho, "side": side})


def _mixed_shapes(seed: int, n: int):
    """50/50 spheres and boxes interleaved by a random bit (so index order is not sorted by kind)."""
    kind = (u01(seed, 20, n) < 0.5).astype(np.uint8)  # 1 == box
    size = np.zeros((n, 3), dtype=np.float32)
    size[:, 0] = _sphere_radii(seed, n)
    box = kind == KIND_BOX
    size[box] = 1.0  # widths = Vector3::one()  (circle_collision.rs:98-101)
    scale = np.ones((n, 3), dtype=np.float32)
    sc = np.stack([0.5 + 3.0 * u01(seed, 30 + a, n) for a in range(3)], axis=1).astype(np.float32)
    scale[box] = sc[box]  # circle_collision.rs:83-87
    q = np.stack([_gauss(seed, 40 + 2 * a, n) for a in range(4)], axis=1)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    quat = _identity_quat(n)
    quat[box] = q[box].astype(np.float32)
    # largest possible AABB extent: a 3.5^3 box on its diagonal
    cell = float(np.sqrt(3.0) * 3.5)
    return kind, size, scale, quat, cell


def mixed_field(n: int = 1 << 22, seed: int = 3, rho: float = 0.5) -> Scene:
    """Config 3: 50% spheres + 50% oriented boxes, uniform density rho per cell volume."""
    kind, size, scale, quat, cell = _mixed_shapes(seed, n)
    side = cell * (n / rho) ** (1.0 / 3.0)
    pos = np.stack([(u01(seed, a, n) - 0.5) * side for a in range(3)], axis=1).astype(np.float32)
    return Scene(f"mixed_{n}", _ids(n), kind, np.zeros((n, 3), dtype=np.float32), size, pos, quat, scale,
                 cell_hint=cell, meta={"seed": seed, "rho": rho, "side": side})


def clustered_field(n: int = 1 << 24, seed: int = 4, rho: float = 0.25, blobs: int = 64,
                    blob_fraction: float = 0.8, peak_per_cell: float = 8.0) -> Scene:
    """Config 4: 80% of the colliders in `blobs` Gaussian blobs (peak <= peak_per_cell colliders per
    cell), 20% uniform background, mean density rho per cell volume."""
    kind, size, scale, quat, cell = _mixed_shapes(seed, n)
    side = cell * (n / rho) ** (1.0 / 3.0)
    n_blob = int(n * blob_fraction)
    per_blob = max(1, n_blob // blobs)
    background = rho * (1.0 - blob_fraction)
    # peak of an isotropic Gaussian of M points: M / ((2 pi)^1.5 sigma^3) per unit volume
    sigma_cells = (per_blob / ((2.0 * np.pi) ** 1.5 * max(peak_per_cell - background, 1e-3))) ** (1.0 / 3.0)
    sigma = sigma_cells * cell
    centres = np.stack([(u01(seed, 50 + a, blobs) - 0.5) * (side - 6.0 * sigma) for a in range(3)], axis=1)
    pos = np.stack([(u01(seed, a, n) - 0.5) * side for a in range(3)], axis=1)
    which = np.minimum((u01(seed, 60, n) * blobs).astype(np.int64), blobs - 1)
    in_blob = u01(seed, 61, n) < blob_fraction
    g = np.stack([_gauss(seed, 70 + 2 * a, n) for a in range(3)], axis=1) * sigma
    blob_pos = centres[which] + g
    half = 0.5 * side
    blob_pos = np.clip(blob_pos, -half, half)
    pos[in_blob] = blob_pos[in_blob]
    return Scene(f"clustered_{n}", _ids(n), kind, np.zeros((n, 3), dtype=np.float32), size,
                 pos.astype(np.float32), quat, scale, cell_hint=cell,
                 meta={"seed": seed, "rho": rho, "side": side, "sigma": sigma, "blobs": blobs})


def by_config(cfg: int, n: int | None = None) -> Scene:
    """Scene of BASELINE.json configs[cfg-1]; `n` overrides the collider count (same generator)."""
    if cfg == 1:
        return bench_scene(n or 1000, seed=1)
    if cfg == 2:
        return sphere_field(n or (1 << 20), seed=2)
    if cfg == 3:
        return mixed_field(n or (1 << 22), seed=3)
    if cfg == 4:
        return clustered_field(n or (1 << 24), seed=4)
    if cfg == 5:
        return sphere_field(n or (1 << 26), seed=5)
    raise ValueError(f"unknown config {cfg}")

"""Estimate (CPU, numpy) of how much of the cell order survives from one frame to the next under the bench's
per-frame jitter (configs[3] generator): the lever of temporal coherence.
Run from the repo root: python tools/estimate_temporal_coherence.py"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import oracle_lib as O  # noqa: E402
from gunship_rs_b200 import scenes  # noqa: E402

n = 400000
s = scenes.by_config(4, n)
sysm = O.OracleSystem(8, 8)


def order(frame):
    pos, quat = s.frame(frame)
    sysm.bvh_update(s.entity, s.kind, s.offset, s.size, pos, quat, s.scale)
    v = sysm.volumes()
    mn = v["aabb_min"][:, :3]
    cell = np.float32(sysm.longest_axis)
    c = np.floor(mn / cell).astype(np.int64)
    c -= c.min(axis=0) - 1
    d = c.max(axis=0) + 2
    key = (c[:, 0] * d[1] + c[:, 1]) * d[2] + c[:, 2]
    perm = np.argsort(key, kind="stable")
    rank = np.empty(n, dtype=np.int64)
    rank[perm] = np.arange(n)
    return key, rank, cell


k0, r0, cell0 = order(1)
k1, r1, cell1 = order(2)
moved = (k0 != k1).mean() if cell0 == cell1 else float("nan")
disp = np.abs(r1 - r0)
print(f"cell size frame 1 / 2: {cell0} / {cell1}")
print(f"colliders whose home cell changed: {100 * moved:.1f}%")
print("rank displacement percentiles (50, 90, 99, max):", np.percentile(disp, [50, 90, 99]).astype(int), disp.max())
for w in (32, 256, 4096):
    print(f"  within +-{w} positions: {100 * (disp <= w).mean():.1f}%")

#!/usr/bin/env python
"""Small target for compute-sanitizer (memcheck / racecheck): smoke() + a two-member group step (halo push through the
k_slab_* kernels, members on one GPU, ordered by stream events) + the tiled pairs kernel, each checked against the oracle.
  compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import __graft_entry__ as E
import oracle_lib as O
from gunship_rs_b200 import context as G
from gunship_rs_b200 import scenes

E.smoke()
s = scenes.by_config(4, 20000)
oracle = O.OracleSystem()
want = oracle.step(s)
with G.CollisionGroup([0, 0], 30000) as grp:
    grp.upload_colliders(s.entity, s.kind, s.offset, s.size, s.pos)
    grp.update_transforms(s.pos, s.quat, s.scale)
    out = grp.step()
    assert np.array_equal(grp.pairs_packed(sorted=True), want)
    grp.reorder()
    grp.update_transforms(s.pos, s.quat, s.scale)
    grp.step()
    assert np.array_equal(grp.pairs_packed(sorted=True), want)
    print("group of 2: ok,", out.n_ghost, "ghosts")
if os.environ.get("GSC_PAIRS_TILED") == "1":
    print("(tiled pairs kernel was used)")
with G.CollisionContext(s.n, contacts=True, graph=True) as ctx:
    ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
    ctx.update_transforms(s.pos, s.quat, s.scale)
    ctx.step()
    assert np.array_equal(ctx.pairs_packed(sorted=True), want)
    ctx.adjacency()
    print("contacts + graph + adjacency: ok")
oracle.close()
print("sanitize target: ok")

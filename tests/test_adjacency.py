"""The consumer of the pair set: CollisionCallbackManager::process_collisions (mod.rs:678-708) turns the pairs into
per-entity lists.  Oracle restatement on CPU, CUDA CSR build against it on the GPU (bit-exact: integer work)."""
import numpy as np
import pytest

import oracle_lib as O
from gunship_rs_b200 import scenes


def test_oracle_adjacency_is_the_symmetric_closure():
    s = scenes.by_config(3, 3000)
    sysm = O.OracleSystem()
    pairs = sysm.step(s)
    ent, off, oth = sysm.process_collisions()
    assert off[0] == 0 and off[-1] == 2 * len(pairs) == len(oth) and np.all(np.diff(ent.astype(np.int64)) > 0)
    got = set()
    for i, e in enumerate(ent):
        lst = oth[off[i]:off[i + 1]]
        assert len(lst) > 0 and np.all(np.diff(lst.astype(np.int64)) > 0)
        got.update((int(e), int(o)) for o in lst)
    want = set()
    for p in pairs:
        a, b = int(p >> np.uint64(32)), int(p & np.uint64(0xFFFFFFFF))
        want.add((a, b))
        want.add((b, a))
    assert got == want
    sysm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,n", [(1, 1000), (2, 50000), (4, 200000)])
def test_cuda_adjacency_matches_oracle(cfg, n):
    from gunship_rs_b200 import context as G
    s = scenes.by_config(cfg, n)
    rng = np.random.default_rng(cfg)
    s.entity = (rng.permutation(s.n) * 3 + 5).astype(np.uint32)   # ids unrelated to the index order
    sysm = O.OracleSystem()
    sysm.step(s)
    want = sysm.process_collisions()
    with G.CollisionContext(n) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        ctx.update_transforms(s.pos, s.quat, s.scale)
        ctx.step()
        got = ctx.adjacency()
        again = ctx.adjacency()                                    # cached until the next step
    for g, w, a in zip(got, want, again):
        assert np.array_equal(g, w) and np.array_equal(a, w)
    sysm.close()


@pytest.mark.gpu
def test_cuda_adjacency_of_an_empty_pair_set():
    from gunship_rs_b200 import context as G
    s = scenes.bench_scene(1)
    with G.CollisionContext(4) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        ctx.update_transforms(s.pos, s.quat, s.scale)
        ctx.step()
        ent, off, oth = ctx.adjacency()
        assert len(ent) == 0 and len(oth) == 0 and off.tolist() == [0]

"""Full-size parity gate: the CUDA pair set against the CPU oracle at the BASELINE.json sizes (configs 2, 3, 4 =
1M / 4M / 16M colliders), bit-exact, plus the radix-sort path those sizes take (25..27-bit cell keys = three passes of
9-bit digits).  BASELINE.md section 2: "bit-exact vs grid_ref on every config".  The oracle (8 work units on 8 threads,
grid_collision.rs:104-105) needs ~1 s at 1M and ~6-10 s at 16M, so these stay in the per-round GPU suite."""
import numpy as np
import pytest

import oracle_lib as O
from gunship_rs_b200 import context as G
from gunship_rs_b200 import scenes

pytestmark = pytest.mark.gpu


def fingerprint(packed: np.ndarray):
    """(count, order-independent checksum): what bench.py prints as config.frame0_pair_set"""
    return int(packed.shape[0]), f"{int(np.sum(packed, dtype=np.uint64)):016x}"


@pytest.mark.parametrize("cfg,n,frames", [(2, 1 << 20, (0, 1)), (3, 1 << 22, (0, 2)), (4, 1 << 24, (0,))])
def test_pair_set_bit_exact_at_baseline_sizes(cfg, n, frames):
    s = scenes.by_config(cfg, n)
    oracle = O.OracleSystem()
    with G.CollisionContext(n, timing=True) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        for f in frames:
            pos, quat = s.frame(f)
            ctx.update_transforms(pos, quat, s.scale)
            out = ctx.step()
            want = oracle.step(s, pos, quat)
            got = ctx.pairs_packed(sorted=True)
            assert out.n_pairs == len(want), (cfg, n, f, out.n_pairs, len(want))
            assert np.array_equal(got, want), (cfg, n, f)
            assert np.float32(out.cell_size) == np.float32(oracle.longest_axis)
            assert out.n_wide == 0 and oracle.stats()["wide"] == 0
            # the fingerprint bench.py prints is a function of the set only
            assert fingerprint(ctx.pairs_packed(sorted=False)) == fingerprint(want)
            if cfg == 4:
                assert out.sort_passes == 3 and 25 <= int(np.ceil(np.log2(out.n_cells + 1))) <= 27  # the 9-bit digit path
    oracle.close()


@pytest.mark.parametrize("n,bits", [(1 << 20, 25), ((1 << 21) + 77, 26), (1 << 22, 27), (1 << 24, 27), (1 << 20, 18), (3 << 19, 9), (1 << 20, 28)])
def test_radix_sort_nine_bit_digits(n, bits):
    """key_bits 25..27 -> 3 passes of 9 bits (512 bins, packed u16 counters); 18 -> 2 x 9; 9 -> one pass; 28 -> 4 x 7"""
    rng = np.random.default_rng(bits * 1000003 + n)
    keys = rng.integers(0, 1 << bits, size=n, dtype=np.uint64).astype(np.uint32)
    keys[: n // 4] = keys[0]                      # a heavy duplicate run: stability + u16 counter pressure
    keys[n // 4: n // 2] &= np.uint32(0x1FF)      # many keys that differ in the lowest digit only
    vals = np.arange(n, dtype=np.uint32)
    k, v = G.sort_pairs_u32(keys, vals, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, order.astype(np.uint32))

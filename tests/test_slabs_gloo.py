"""N>1 host logic on CPU: world_size-2 (and 3) gloo process groups drive gunship_rs_b200.slabs.SlabDriver with
the CPU stand-in backend; the union of the ranks' pair lists must be the oracle's single-process pair set."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cfg, n, frames, ownership, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from gunship_rs_b200 import scenes, slabs
    from slab_cpu_backend import CpuSlabBackend
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    s = scenes.by_config(cfg, n)
    if ownership == "slabs":
        owner = slabs.partition_x_slabs(s.pos[:, 0], world)
    elif ownership == "scattered":
        owner = (np.arange(s.n) % world).astype(np.int32)
    else:  # one rank owns nothing
        owner = np.zeros(s.n, dtype=np.int32)
    mine = np.nonzero(owner == rank)[0]
    be = CpuSlabBackend(s.entity[mine], s.kind[mine], s.offset[mine], s.size[mine], mine)
    drv = slabs.SlabDriver(be)
    # pipelined host loop: the next frame is handed over from the after_bounds hook of the step in flight (the
    # frame's transforms have been consumed by then), as bench.py's multi-GPU e2e loop does
    hooks = []

    def hand_over(f):
        hooks.append(f)
        pos, quat = s.frame(f)
        be.set_frame(pos[mine], quat[mine], s.scale[mine])

    hand_over(frames[0])
    for k, f in enumerate(frames):
        nxt = frames[k + 1] if k + 1 < len(frames) else None
        out = drv.step(after_bounds=(lambda: hand_over(nxt)) if nxt is not None else None)
        np.save(os.path.join(out_dir, f"p_{f}_{rank}.npy"), out.pairs)
        np.save(os.path.join(out_dir, f"h_{f}_{rank}.npy"), np.array([drv.last_halo_in, drv.last_halo_out, len(mine)]))
    assert hooks == list(frames)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,cfg,n,ownership", [(2, 4, 4000, "slabs"), (2, 3, 3000, "scattered"), (3, 2, 4000, "slabs"),
                                                   (2, 1, 1000, "empty")])
def test_union_of_rank_pairs_is_the_single_process_set(tmp_path, world, cfg, n, ownership):
    import torch.multiprocessing as mp
    import oracle_lib as O
    from gunship_rs_b200 import scenes
    frames = (0, 2)
    mp.spawn(_worker, args=(world, _free_port(), cfg, n, frames, ownership, str(tmp_path)), nprocs=world, join=True)
    s = scenes.by_config(cfg, n)
    oracle = O.OracleSystem()
    for f in frames:
        pos, quat = s.frame(f)
        want = oracle.step(s, pos, quat)
        parts = [np.load(tmp_path / f"p_{f}_{r}.npy") for r in range(world)]
        got = np.sort(np.concatenate(parts))
        assert len(got) == len(want) and np.array_equal(got, want)   # complete and duplicate free
        halo = [np.load(tmp_path / f"h_{f}_{r}.npy") for r in range(world)]
        assert sum(h[0] for h in halo) == sum(h[1] for h in halo)     # everything sent was received
        if ownership == "slabs":
            assert all(h[0] < 0.6 * max(h[2], 1) for h in halo)       # the halo is a boundary layer, not the world
            assert halo[-1][0] > 0
        if ownership == "empty":
            assert len(parts[1]) == 0 and halo[1][0] == 0
    oracle.close()


def test_partition_x_slabs_is_balanced_and_ordered():
    from gunship_rs_b200 import slabs
    x = np.random.default_rng(0).normal(size=10001).astype(np.float32)
    for world in (1, 2, 3, 8):
        owner = slabs.partition_x_slabs(x, world)
        counts = np.bincount(owner, minlength=world)
        assert counts.max() - counts.min() <= 1
        for r in range(world - 1):
            assert x[owner == r].max() <= x[owner == r + 1].min()
    assert slabs.needed_range(5, 9) == (4, 9)


def test_weighted_slab_partition_balances_the_estimated_work():
    """partition_x_slabs(x, world, weight): cuts at equal cumulative weight along x; pair_work_estimate: 26 + 9 x cell
    occupancy (what bench.py starts its slab calibration from)"""
    import numpy as np
    from gunship_rs_b200 import scenes, slabs
    s = scenes.by_config(4, 60000)
    w = slabs.pair_work_estimate(s.pos, s.cell_hint)
    assert w.shape == (s.n,) and w.min() >= 35.0          # a collider always shares its own cell with itself
    # occupancy really is the number of colliders with the same home cell
    cell = np.floor(s.pos.astype(np.float64) / s.cell_hint).astype(np.int64)
    i = int(np.argmax(w))
    same = int(np.all(cell == cell[i], axis=1).sum())
    assert w[i] == 26.0 + 9.0 * same and same > 1
    for world in (2, 4, 8):
        owner = slabs.partition_x_slabs(s.pos[:, 0], world, w)
        share = np.array([w[owner == r].sum() for r in range(world)]) / w.sum()
        assert np.all(np.abs(share - 1.0 / world) < 0.01)     # equal work ...
        xs = [s.pos[owner == r, 0] for r in range(world)]
        assert all(xs[r].max() <= xs[r + 1].min() for r in range(world - 1))   # ... in x-slabs
        count = slabs.partition_x_slabs(s.pos[:, 0], world)
        assert np.bincount(count, minlength=world).max() - np.bincount(count, minlength=world).min() <= 1


def _control_worker(rank, world, port, mode, out_dir):
    """rank 1 leaves the protocol (abort flag, or simply never arrives); rank 0 must not spin forever"""
    sys.path.insert(0, ROOT)
    import time
    import torch.distributed as dist
    from gunship_rs_b200 import slabs
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    if mode == "timeout":
        slabs.ShmControl.TIMEOUT_S = 1.0
    ctl = slabs.ShmControl(dist, None, rank, world)
    first = ctl.all_gather([float(rank + 1), 10.0 * (rank + 1)])       # a healthy round first
    assert first.shape == (world, 2) and first[:, 0].tolist() == [1.0, 2.0]
    ctl.barrier()
    t0 = time.monotonic()
    outcome = "returned"
    if rank == 0:
        try:
            ctl.all_gather([1.0])
        except RuntimeError as e:          # the peer aborted its step
            outcome = "aborted: " + str(e)
        except TimeoutError as e:          # the peer never arrived
            outcome = "timeout: " + str(e)
    elif mode == "abort":
        time.sleep(0.3)                    # rank 0 is spinning by now
        ctl.abort()
    else:
        time.sleep(2.5)                    # never joins the second round
    with open(os.path.join(out_dir, f"o_{rank}.txt"), "w") as f:
        f.write(f"{outcome}\n{time.monotonic() - t0:.3f}\n")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["abort", "timeout"])
def test_control_plane_fails_together_instead_of_spinning(tmp_path, mode):
    """a rank that leaves its step with an error (or dies) must take the others with it -- the spin on the
    shared-memory control block has an abort flag and a deadline."""
    import torch.multiprocessing as mp
    from gunship_rs_b200 import slabs
    if not slabs.ShmControl.usable():
        pytest.skip("shared-memory control plane is x86 only")
    mp.spawn(_control_worker, args=(2, _free_port(), mode, str(tmp_path)), nprocs=2, join=True)
    outcome, dt = (tmp_path / "o_0.txt").read_text().splitlines()
    if mode == "abort":
        assert outcome.startswith("aborted: ") and float(dt) < 2.0
    else:
        assert outcome.startswith("timeout: ") and 0.9 <= float(dt) < 2.4

"""Multi-GPU path on the box's GPU(s): the slab driver with real CUDA contexts.  Two ranks (two processes,
gloo rendezvous on 127.0.0.1, both on cuda:0 when the box has one GPU -- the halo then travels through host
staging) must together report exactly the single-context pair set, which is the oracle's."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cfg, n, frames, out_dir, ownership, push):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from gunship_rs_b200 import context as G
    from gunship_rs_b200 import scenes, slabs
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()
    s = scenes.by_config(cfg, n)
    if ownership == "slabs":
        owner = slabs.partition_x_slabs(s.pos[:, 0], world)
    else:  # scattered ownership: correct, merely a bigger halo
        owner = (np.arange(s.n) % world).astype(np.int32)
    mine = np.nonzero(owner == rank)[0]
    with G.CollisionContext(s.n, device=dev) as ctx:
        ctx.upload_colliders(s.entity[mine], s.kind[mine], s.offset[mine], s.size[mine], global_index=mine.astype(np.uint32))
        drv = slabs.SlabDriver(slabs.GpuSlabBackend(ctx, dev, host_staging=True))
        if push:
            drv.enable_peer_push()   # CUDA IPC: ghosts written straight into the other rank's arrays
        for f in frames:
            pos, quat = s.frame(f)
            ctx.update_transforms(pos[mine], quat[mine], s.scale[mine])
            out = drv.step()
            assert out.n_wide == 0
            np.save(os.path.join(out_dir, f"pairs_{ownership}_{f}_{rank}.npy"), ctx.pairs_packed(sorted=True))
            np.save(os.path.join(out_dir, f"halo_{ownership}_{f}_{rank}.npy"), np.array([drv.last_halo_in, drv.last_halo_out, len(mine)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cfg,n,ownership,push", [(4, 60000, "slabs", False), (3, 30000, "slabs", False), (2, 30000, "scattered", False),
                                                  (4, 60000, "slabs", True), (3, 30000, "scattered", True)])
def test_two_ranks_report_the_single_context_pair_set(tmp_path, cfg, n, ownership, push):
    import torch.multiprocessing as mp
    import oracle_lib as O
    from gunship_rs_b200 import scenes
    frames = (0, 2)
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), cfg, n, frames, str(tmp_path), ownership, push), nprocs=world, join=True)
    s = scenes.by_config(cfg, n)
    oracle = O.OracleSystem()
    for f in frames:
        pos, quat = s.frame(f)
        want = oracle.step(s, pos, quat)
        parts = [np.load(tmp_path / f"pairs_{ownership}_{f}_{r}.npy") for r in range(world)]
        got = np.sort(np.concatenate(parts))
        assert len(got) == len(want) and np.array_equal(got, want)         # union == oracle, no duplicates
        assert all(len(p) > 0 for p in parts)
        halo = [np.load(tmp_path / f"halo_{ownership}_{f}_{r}.npy") for r in range(world)]
        if ownership == "slabs" and not push:
            # the halo flows in +x: rank 1 receives the boundary layer (two cell columns), rank 0 only the
            # foreign volumes of the one cell column the cut runs through
            assert 0 < halo[1][0] < 0.35 * halo[1][2] and halo[0][0] < halo[1][0]
    oracle.close()


@pytest.mark.parametrize("world,cfg,n", [(2, 4, 1 << 21), (4, 4, 1 << 21), (8, 3, 400000)])
def test_one_rank_per_gpu_device_driven_exchange(tmp_path, world, cfg, n):
    """`world` processes on `world` DIFFERENT GPUs (skipped on smaller boxes): the exchange runs on the device (k_slab_*,
    CUDA IPC mailboxes, one host wait per step) and the union of the ranks' lists is the oracle's set -- at 2M colliders
    for 2 and 4 ranks (the N-GPU union had only been checked at 2 ranks and 60k)."""
    import torch
    import torch.multiprocessing as mp
    import oracle_lib as O
    from gunship_rs_b200 import scenes
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    frames = (0, 1)
    mp.spawn(_worker, args=(world, _free_port(), cfg, n, frames, str(tmp_path), "slabs", True), nprocs=world, join=True)
    s = scenes.by_config(cfg, n)
    oracle = O.OracleSystem()
    for f in frames:
        pos, quat = s.frame(f)
        want = oracle.step(s, pos, quat)
        parts = [np.load(tmp_path / f"pairs_slabs_{f}_{r}.npy") for r in range(world)]
        got = np.sort(np.concatenate(parts))
        assert len(got) == len(want) and np.array_equal(got, want)
        assert all(len(p) > 0 for p in parts)
    oracle.close()


def test_phase_calls_single_rank_equal_gsc_step():
    """gsc_phase_bounds / gsc_set_global_bounds / gsc_local_x_range / gsc_set_x_window / gsc_phase_finish on ONE context
    reproduce gsc_step; a window that misses a local collider is rejected."""
    import oracle_lib as O
    from gunship_rs_b200 import _native as N
    from gunship_rs_b200 import context as G
    from gunship_rs_b200 import scenes
    s = scenes.by_config(3, 20000)
    oracle = O.OracleSystem()
    want = oracle.step(s)
    vol = oracle.volumes()
    cell = np.float32(oracle.longest_axis)
    home_x = np.floor(vol["aabb_min"][:, 0] / cell).astype(np.int64)
    with G.CollisionContext(s.n) as ctx:
        ctx.upload_colliders(s.entity, s.kind, s.offset, s.size)
        ctx.update_transforms(s.pos, s.quat, s.scale)
        b = ctx.phase_bounds()
        assert np.float32(b.longest_axis) == cell and np.float32(b.home_x_max) == vol["aabb_min"][:, 0].max()
        assert np.array_equal(np.array(list(b.world_min), np.float32), vol["aabb_min"][:, :3].min(axis=0))
        ctx.set_global_bounds(b)
        assert ctx.local_x_range() == (int(home_x.min()), int(home_x.max()))
        ctx.set_x_window(int(home_x.min()) - 1, int(home_x.max()))
        out = ctx.phase_finish()
        assert out.grid_dims[0] == home_x.max() - home_x.min() + 3          # window clamped to the world, + apron
        assert np.array_equal(ctx.pairs_packed(sorted=True), want)
        # a window that does not cover every local collider is a caller bug
        ctx.update_transforms(s.pos, s.quat, s.scale)
        ctx.set_global_bounds(ctx.phase_bounds())
        ctx.set_x_window(int(home_x.min()) + 2, int(home_x.max()))
        with pytest.raises(G.GscError) as ei:
            ctx.phase_finish()
        assert ei.value.code == N.GSC_ERR_INVALID
        # and the plain step still works afterwards
        ctx.update_transforms(s.pos, s.quat, s.scale)
        ctx.step()
        assert np.array_equal(ctx.pairs_packed(sorted=True), want)
    oracle.close()

"""Multi-GPU host layer: x-slab partition of the collider field, one rank per GPU, halo exchange over
torch.distributed (NCCL on the GPU box, gloo in the CPU tests).

The reference already decomposes space: `GridCollisionSystem` cuts the world into 8 octant work units,
every unit scans all volumes, keeps those whose AABB touches its region and the duplicates across region
borders disappear when the per-unit hash sets are merged (grid_collision.rs:60-86, 161-193, 262-272).
Here the same idea is laid out for 8 GPUs without the merge:

* every rank OWNS a subset of the colliders -- `partition_x_slabs` cuts the field into count-balanced
  x-slabs, but any assignment is correct, a spatially coherent one is merely cheap;
* a contact pair is reported by exactly one rank, the owner of its *taker*: the volume that comes later in
  (home-cell key, global index) order.  The taker only ever looks at home x-cells {x-1, x}, so rank r needs
  as ghosts the foreign volumes whose home x-cell lies in [xmin_r - 1, xmax_r];
* per step ONE small all-gather of the local bounds gives every rank the global cell size, the world box and
  all ranks' home x-ranges; then the ghost counts are all-gathered and the ghost records go point to point.  No collective touches the pair lists: the union
  of the per-rank lists IS the single-GPU pair set.

The device work is behind a small backend interface so the same driver runs on the CUDA contexts
(`GpuSlabBackend`) and, in the CPU tests, on a numpy stand-in that checks the protocol with gloo.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

from . import _native as N

RECORD_BYTES = N.GSC_HALO_RECORD_BYTES


def partition_x_slabs(x: np.ndarray, world: int, weight: np.ndarray | None = None) -> np.ndarray:
    """x-slabs: owner rank of every collider from the quantiles of its x position -- count-balanced, or balanced
    by `weight` (an estimate of every collider's cost, see pair_work_estimate) when given.

    Returns int32[N].  Ties at a cut go to the lower rank.  (The cut need not fall on a cell boundary:
    ownership only decides who reports a pair, the ghost range is derived from the home cells.)"""
    n = x.shape[0]
    if world <= 1 or n == 0:
        return np.zeros(n, dtype=np.int32)
    order = np.argsort(x, kind="stable")
    owner = np.empty(n, dtype=np.int32)
    if weight is None:
        bounds = [(n * r) // world for r in range(world + 1)]
    else:
        cum = np.cumsum(weight[order], dtype=np.float64)
        bounds = [0] + [int(np.searchsorted(cum, cum[-1] * r / world, side="left")) for r in range(1, world)] + [n]
    for r in range(world):
        owner[order[bounds[r]:bounds[r + 1]]] = r
    return owner


def pair_work_estimate(pos: np.ndarray, cell: float) -> np.ndarray:
    """Relative cost of every collider in a step: the per-collider stages (bvh_update, keys, sort, table) cost about as
    much as 26 quantised AABB tests, and a collider in a cell with c occupants meets ~9 c candidates in its half shell
    (13 neighbour cells + its own, after the reach pruning).  Slabs cut by this weight even out the pairs stage on
    clustered scenes, where count-balanced slabs leave the ranks that hold a blob 30% longer in k_pairs."""
    c = np.floor(pos.astype(np.float64) / float(cell)).astype(np.int64)
    c -= c.min(axis=0)
    dims = c.max(axis=0) + 1
    key = (c[:, 0] * dims[1] + c[:, 1]) * dims[2] + c[:, 2]
    _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    return 26.0 + 9.0 * cnt[inv].astype(np.float64)


def needed_range(x_min: int, x_max: int):
    """Home x-cells whose volumes a rank with home range [x_min, x_max] must see: [x_min - 1, x_max]."""
    return x_min - 1, x_max


class ShmControl:
    """Host-side all-gather / barrier between the ranks of ONE box over a shared-memory file (/dev/shm).

    The per-step control data is ten numbers per rank and is already on the host when it is needed
    (gsc_phase_bounds returns it), so routing it through a device collective would cost two extra copies and a
    kernel launch per step; a cache-line spin on shared memory costs a few microseconds.  Two banks alternate
    so that a fast rank can never overwrite a row a slow rank has not read yet (to finish gather k+1 every rank
    must have published k+1, i.e. finished reading k)."""

    SLOTS = 16

    def __init__(self, dist, group, rank, world):
        import uuid
        self.rank, self.world, self.seq = rank, world, 0
        box = [f"/dev/shm/gsc_ctl_{uuid.uuid4().hex}" if rank == 0 else None]
        shape = (2, world, self.SLOTS + 1)
        if rank == 0:
            self.mem = np.memmap(box[0], dtype=np.float64, mode="w+", shape=shape)
            self.mem[:] = 0.0
            self.mem.flush()
        dist.broadcast_object_list(box, src=0 if group is None else dist.get_global_rank(group, 0), group=group)
        if rank != 0:
            self.mem = np.memmap(box[0], dtype=np.float64, mode="r+", shape=shape)
        dist.barrier(group=group)
        if rank == 0:
            os.unlink(box[0])  # stays alive while mapped; nothing is left behind

    TIMEOUT_S = float(os.environ.get("GSC_CONTROL_TIMEOUT_S", "60"))

    @staticmethod
    def usable():
        """The publish below is 'payload, then sequence number' through plain numpy stores: correct on hosts that keep
        store order (x86).  Anywhere else (aarch64: Grace) the driver falls back to torch.distributed collectives."""
        import platform
        return platform.machine().lower() in ("x86_64", "amd64", "i686", "i386")

    def all_gather(self, values):
        """(x86 only: the sequence number is published after the payload and relies on the store order of the
        platform; the default exchange -- GSC_EXCHANGE=device -- does not use this class at all)"""
        self.seq += 1
        bank = self.mem[self.seq & 1]
        n = len(values)
        if n:
            bank[self.rank, 1:1 + n] = values
        bank[self.rank, 0] = self.seq          # publish after the payload
        seqs = self.mem[:, :, 0]
        deadline = time.monotonic() + self.TIMEOUT_S
        spins = 0
        while (seqs[self.seq & 1] < self.seq).any():
            if (seqs < 0).any():               # a peer left its step with an error: fail together instead of spinning
                raise RuntimeError("slab control: a peer rank aborted its step")
            spins += 1
            if spins & 0xFFF == 0 and time.monotonic() > deadline:
                self.abort()
                raise TimeoutError(f"slab control: a peer rank did not arrive within {self.TIMEOUT_S:.0f} s")
        return np.array(bank[:, 1:1 + n])

    def barrier(self):
        self.all_gather(())

    def abort(self):
        """Poison both banks: every rank spinning in (or entering) a gather raises instead of waiting forever."""
        self.mem[0, self.rank, 0] = -1.0
        self.mem[1, self.rank, 0] = -1.0


class SlabDriver:
    """One collision step across `world` ranks.  `backend` does the device work of this rank:

    phase_bounds() -> (longest_axis, world_min[3], world_max[3], status_flags, home_x_max, n_local)
    set_global_bounds(longest_axis, world_min, world_max, status_flags)
    set_x_window(x_lo, x_hi)                                home x-cells this rank's grid must cover
    halo_select(x_lo, x_hi) -> (count, tensor)              uint8 tensor, count * RECORD_BYTES valid bytes
    new_buffer(count) -> tensor                             receive buffer on the backend's device
    halo_append(count, tensor)
    phase_finish() -> step result
    tensor_device                                           device of the small control tensors
    """

    def __init__(self, backend, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.backend = backend
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.last_halo_in = 0
        self.last_halo_out = 0
        self.push = False
        self.shm = None
        self.device_exchange = False
        self.trace = {} if os.environ.get("GSC_SLAB_TRACE") else None  # host seconds per phase, summed over steps

    def enable_peer_push(self):
        """GPUs of one box: open every other rank's record arrays over CUDA IPC so that the halo is written
        straight into the neighbour's memory by the selecting kernel (gsc_halo_push) -- no staging buffers, no
        count exchange, no send/recv.  Collective: every rank of the group must call it."""
        import torch
        dist, be = self.dist, self.backend
        dev = be.tensor_device
        mine = torch.frombuffer(bytearray(be.ipc_export()), dtype=torch.uint8).to(dev)
        blobs = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(blobs, mine, group=self.group)
        for s_, b in enumerate(blobs):
            if s_ != self.rank:
                be.ipc_open_peer(s_, bytes(b.cpu().numpy().tobytes()))
        self.push = True
        # default: the whole exchange runs on the device (bounds, halo and flags through peer memory, ONE host wait per
        # step, gsc_slab_step); GSC_EXCHANGE=host keeps the round-1 host-driven sequence (three syncs + two rendezvous)
        # Ranks that share a GPU must NOT wait for each other inside kernels (nothing guarantees that the kernels of two
        # processes on one GPU run at the same time): they keep the host-driven sequence.
        devs = [None] * self.world
        dist.all_gather_object(devs, getattr(be, "device_id", lambda: self.rank)(), group=self.group)
        distinct = len(set(devs)) == self.world
        if os.environ.get("GSC_EXCHANGE", "device") == "device" and hasattr(be, "slab_init") and distinct:
            be.slab_init(self.rank, self.world)
            self.device_exchange = True
        elif os.environ.get("GSC_CONTROL", "shm") == "shm" and ShmControl.usable():
            self.shm = ShmControl(dist, self.group, self.rank, self.world)  # same box: host control plane

    def _t(self, name, t0):
        if self.trace is not None:
            self.trace[name] = self.trace.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    def step(self, after_bounds=None):
        """One collision step over all ranks.  `after_bounds`, if given, is called as soon as stage 0 (bvh_update) of
        this step is done on this rank: the one reader of the frame's transform arrays has finished, so a host loop
        hands over the NEXT frame's transforms there and the PCIe copy runs under the exchange and stages 1-5."""
        import torch
        dist, be = self.dist, self.backend
        dev = be.tensor_device
        t0 = time.perf_counter()
        if self.device_exchange:
            be.slab_step_begin()                 # bvh_update, exchange and stages 1-5: all enqueued, no host wait
            t0 = self._t("enqueue", t0)
            if after_bounds is not None:
                after_bounds()                   # (the copy is ordered behind bvh_update of the step in flight)
                t0 = self._t("after_bounds", t0)
            out = be.step_end()
            self._t("wait", t0)
            self.last_halo_in = int(out.n_ghost)
            return out
        try:
            return self._step_host(after_bounds, t0)
        except Exception:
            if self.shm is not None:
                self.shm.abort()                 # the other ranks must not wait for this one
            raise

    def _step_host(self, after_bounds, t0):
        import torch
        dist, be = self.dist, self.backend
        dev = be.tensor_device
        longest, wmin, wmax, status, home_x_max, n_local = be.phase_bounds()
        t0 = self._t("phase_bounds", t0)
        if after_bounds is not None:
            after_bounds()
            t0 = self._t("after_bounds", t0)
        # ONE all-gather carries everything the ranks must agree on: the global cell size (max longest axis) and
        # world box (they place the grid), and every rank's home x-range in world units
        # (f64 carries the f32 values exactly, and the collider counts the peer push needs)
        mine = [longest, *wmin, *wmax, home_x_max, float(status), float(n_local)]
        if self.shm is not None:
            allb = self.shm.all_gather(mine)                             # [world, 10]
        else:
            t = torch.tensor(mine, dtype=torch.float64, device=dev)
            allb = [torch.zeros(10, dtype=torch.float64, device=dev) for _ in range(self.world)]
            dist.all_gather(allb, t, group=self.group)
            allb = torch.stack(allb).cpu().numpy()
        t0 = self._t("bounds_all_gather", t0)
        peer_n_local = allb[:, 9].astype(np.int64)
        allb = allb.astype(np.float32)
        cell = np.float32(allb[:, 0].max())
        g_min = allb[:, 1:4].min(axis=0)
        g_max = allb[:, 4:7].max(axis=0)
        g_status = int(np.bitwise_or.reduce(allb[:, 8].astype(np.int64)))
        be.set_global_bounds(float(cell), g_min.tolist(), g_max.tolist(), g_status)
        # home x-cell = floor(fl(aabb.min.x / cell)) (grid_collision.rs:329-335): IEEE f32 divide, same as the device
        # (one vectorised f32 divide for all ranks: this runs on the critical path of every step)
        with np.errstate(all="ignore"):
            lo_x = np.floor(allb[:, 1] / cell)
            hi_x = np.floor(allb[:, 7] / cell)
        has = (allb[:, 0] > 0) & np.isfinite(allb[:, 1]) & bool(cell > 0)
        ranges = [(int(lo_x[r]), int(hi_x[r])) if has[r] else (1, 0) for r in range(self.world)]  # (1, 0): no colliders
        x_min, x_max = ranges[self.rank]
        if x_min <= x_max:
            be.set_x_window(*needed_range(x_min, x_max))

        t0 = self._t("host_ranges_set_bounds", t0)
        if self.push:
            # fused select + transfer: my volumes in rank s's needed range go straight into its arrays
            pushed_to = 0
            if x_min <= x_max:
                for s, (smin, smax) in enumerate(ranges):
                    if s == self.rank or smin > smax:
                        continue
                    lo, hi = needed_range(smin, smax)
                    lo, hi = max(lo, x_min), min(hi, x_max)
                    if lo <= hi:
                        be.halo_push(s, lo, hi, int(peer_n_local[s]))
                        pushed_to += 1
            t0 = self._t("push_launch", t0)
            be.halo_sync()                       # my pushes have landed ...
            t0 = self._t("push_sync", t0)
            if self.shm is not None:             # ... and so have everyone else's
                self.shm.barrier()
            else:
                dist.barrier(group=self.group)
            t0 = self._t("barrier", t0)
            self.last_halo_in = be.halo_commit()
            self.last_halo_out = pushed_to
            t0 = self._t("commit", t0)
            out = be.phase_finish()
            self._t("phase_finish", t0)
            return out

        # what each other rank needs from me
        send = {}
        counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        if x_min <= x_max:
            for s, (smin, smax) in enumerate(ranges):
                if s == self.rank or smin > smax:
                    continue
                lo, hi = needed_range(smin, smax)
                lo, hi = max(lo, x_min), min(hi, x_max)
                if lo > hi:
                    continue
                cnt, buf = be.halo_select(lo, hi)
                if cnt:
                    send[s] = (cnt, buf)
                    counts[s] = cnt
        matrix = [torch.zeros(self.world, dtype=torch.int64, device=dev) for _ in range(self.world)]
        dist.all_gather(matrix, counts, group=self.group)
        recv_counts = [int(matrix[s][self.rank].item()) for s in range(self.world)]

        ops, recv = [], {}
        for s, (cnt, buf) in send.items():
            ops.append(dist.P2POp(dist.isend, buf[: cnt * RECORD_BYTES], self._peer(s), group=self.group))
        for s, cnt in enumerate(recv_counts):
            if s != self.rank and cnt:
                recv[s] = (cnt, be.new_buffer(cnt))
                ops.append(dist.P2POp(dist.irecv, recv[s][1][: cnt * RECORD_BYTES], self._peer(s), group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            be.sync()
        self.last_halo_out = sum(c for c, _ in send.values())
        self.last_halo_in = sum(c for c, _ in recv.values())
        for s in sorted(recv):
            be.halo_append(*recv[s])
        return be.phase_finish()

    def _peer(self, s: int) -> int:
        return s if self.group is None else self.dist.get_global_rank(self.group, s)


class GpuSlabBackend:
    """The CUDA context of this rank behind the SlabDriver interface (device buffers are torch tensors so
    that torch.distributed can move them; every compute call goes through the C ABI).

    host_staging=True routes the control tensors and the halo payload through host memory, for process
    groups that cannot move CUDA tensors (gloo) -- used by the 2-ranks-on-one-GPU parity test."""

    def __init__(self, ctx, device, host_staging=False):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.cuda = torch.device("cuda", device)
        self.host_staging = host_staging
        self.tensor_device = torch.device("cpu") if host_staging else self.cuda
        self._send = {}
        self._cap = 0

    def phase_bounds(self):
        b = self.ctx.phase_bounds()
        return (float(b.longest_axis), list(b.world_min), list(b.world_max), int(b.status_flags), float(b.home_x_max),
                self.ctx.n)

    def set_global_bounds(self, longest, wmin, wmax, status):
        b = N.GscBounds()
        b.longest_axis = longest
        for a in range(3):
            b.world_min[a] = wmin[a]
            b.world_max[a] = wmax[a]
        b.status_flags = status
        self.ctx.set_global_bounds(b)

    def set_x_window(self, x_lo, x_hi):
        self.ctx.set_x_window(x_lo, x_hi)

    def _device_buffer(self, count):
        return self.torch.empty(max(count, 1) * RECORD_BYTES, dtype=self.torch.uint8, device=self.cuda)

    def new_buffer(self, count):
        if self.host_staging:
            return self.torch.empty(max(count, 1) * RECORD_BYTES, dtype=self.torch.uint8)
        return self._device_buffer(count)

    def halo_select(self, x_lo, x_hi):
        key = (x_lo, x_hi)
        cap = max(self._cap, 1024)
        while True:
            buf = self._send.get(key)
            if buf is None or buf.numel() < cap * RECORD_BYTES:
                buf = self._device_buffer(cap)
                if len(self._send) > 8:
                    self._send = {}
                self._send[key] = buf
            try:
                cnt = self.ctx.halo_select(x_lo, x_hi, buf.data_ptr(), buf.numel() // RECORD_BYTES)
                self._cap = max(self._cap, cnt)
                return cnt, (buf[: max(cnt, 1) * RECORD_BYTES].cpu() if self.host_staging else buf)
            except N.GscError as e:
                if e.code != N.GSC_ERR_CAPACITY or not getattr(e, "needed", 0):
                    raise
                cap = int(e.needed * 1.25) + 1024

    def halo_append(self, count, buf):
        if self.host_staging:
            buf = buf.to(self.cuda)
        self.ctx.halo_append(count, buf.data_ptr())

    def sync(self):
        self.torch.cuda.synchronize(self.cuda)

    # peer push (CUDA IPC)
    def ipc_export(self):
        return self.ctx.ipc_export()

    def ipc_open_peer(self, peer, blob):
        self.ctx.ipc_open_peer(peer, blob)

    def halo_push(self, peer, x_lo, x_hi, peer_n_local):
        self.ctx.halo_push(peer, x_lo, x_hi, peer_n_local)

    def halo_sync(self):
        self.ctx.halo_sync()

    def halo_commit(self):
        return self.ctx.halo_commit()

    def phase_finish(self):
        return self.ctx.phase_finish()

    # device-driven exchange
    def device_id(self):
        """identifies the physical GPU of this rank (ranks that share one keep the host-driven exchange)"""
        try:
            return str(self.torch.cuda.get_device_properties(self.cuda).uuid)
        except Exception:
            return f"{os.uname().nodename}:{self.cuda.index}"

    def slab_init(self, rank, world):
        self.ctx.slab_init(rank, world)

    def slab_step_begin(self):
        self.ctx.slab_step_begin()

    def step_end(self):
        return self.ctx.step_end()


# ------------------------------------------------------------------------------------------------------
# bench.py --gpus N  (launched under torchrun, one rank per GPU)
# ------------------------------------------------------------------------------------------------------

def bench_multi_gpu(args, n, rank, world, local_rank, metric, unit, workload_name, measured_peak_gbs, ClockSampler,
                    stage_bytes, emit):
    import torch
    import torch.distributed as dist

    from . import context as G
    from . import scenes

    dev = torch.device("cuda", local_rank)
    scene = scenes.by_config(args.config, n)
    n = scene.n
    balance = os.environ.get("GSC_SLAB_BALANCE", "work")
    weight = pair_work_estimate(scene.pos, scene.cell_hint) if balance == "work" and world > 1 else None
    owner = partition_x_slabs(scene.pos[:, 0], world, weight)
    nframes = max(1, args.frames)
    frames_full = [scene.frame(f) for f in range(nframes)]
    f_box = float((scene.kind == 1).mean())
    cap = int(n / world * 1.6) + (1 << 16)  # room for ghosts and for the re-cut slabs of the calibration below
    ctx = G.CollisionContext(cap, device=local_rank, timing=True)
    driver = SlabDriver(GpuSlabBackend(ctx, local_rank))
    if os.environ.get("GSC_HALO", "push") == "push":
        driver.enable_peer_push()  # halo written straight into the neighbour's HBM over NVLink
    part = {}

    def load_partition(owner):
        """(re)build this rank's share: static collider data with global indices, pinned host frames, device frames"""
        mine = np.nonzero(owner == rank)[0]
        part["mine"] = mine
        part["h_pos"], part["h_quat"] = [], []
        for pos, quat in frames_full:
            hp = torch.empty((mine.shape[0], 3), dtype=torch.float32, pin_memory=True)
            hq = torch.empty((mine.shape[0], 4), dtype=torch.float32, pin_memory=True)
            hp.numpy()[:] = pos[mine]
            hq.numpy()[:] = quat[mine]
            part["h_pos"].append(hp)
            part["h_quat"].append(hq)
        part["h_scale"] = torch.from_numpy(np.ascontiguousarray(scene.scale[mine])).pin_memory()
        part["box"] = scene.kind[mine] == 1
        ctx.upload_colliders(scene.entity[mine], scene.kind[mine], scene.offset[mine], scene.size[mine],
                             global_index=mine.astype(np.uint32))
        finish_frames()
        # the ranks step together (the device-side waits carry a time limit): line them up after the host-side set-up,
        # whose duration differs from rank to rank by seconds
        dist.barrier()

    def finish_frames():
        # the host layer marshals rotation / scale for the boxes only (spheres ignore them, mod.rs:313-317)
        box = part["box"]
        part["n_box"] = int(box.sum())
        part["h_bquat"] = [torch.from_numpy(np.ascontiguousarray(t.numpy()[box])).pin_memory() for t in part["h_quat"]]
        part["h_bscale"] = torch.from_numpy(np.ascontiguousarray(part["h_scale"].numpy()[box])).pin_memory()
        part["d_pos"] = [t.to(dev) for t in part["h_pos"]]
        part["d_quat"] = [t.to(dev) for t in part["h_quat"]]
        part["d_scale"] = part["h_scale"].to(dev)

    def step_resident(i):
        f = i % nframes
        ctx.bind_device_transforms(part["d_pos"][f].data_ptr(), part["d_quat"][f].data_ptr(), part["d_scale"].data_ptr())
        return driver.step()

    load_partition(owner)

    # Slab calibration: the a-priori weights give the SHAPE of the cost density, a few measured steps its SCALE per
    # slab (device time of this rank's own stages, the exchange interval -- which holds the waits for the other ranks --
    # left out); the slabs are re-cut by the corrected weights.  What an engine would do with gsc_group_rebalance.
    calib = int(os.environ.get("GSC_SLAB_CALIBRATE", "2")) if (weight is not None and world > 1) else 0
    calib_log = []
    for _ in range(calib):
        ctx.set_flags(timing=True, graph=False)
        step_resident(0)
        busy = np.zeros(1)
        for i in range(3):
            o = step_resident(i)
            sm = list(o.stage_ms)
            busy += sm[0] + sm[2] + sm[3] + sm[4] + sm[5]
        mine = part["mine"]
        t = torch.tensor([float(busy[0]) / 3.0, float(weight[mine].sum())], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        allt = torch.stack(allt).cpu().numpy()
        alpha = allt[:, 0] / np.maximum(allt[:, 1], 1e-9)
        alpha /= alpha.mean()
        calib_log.append([round(float(x), 4) for x in allt[:, 0]])
        weight = weight * alpha[owner]
        owner = partition_x_slabs(scene.pos[:, 0], world, weight)
        load_partition(owner)
    n_local = int(part["mine"].shape[0])

    def timed(fn, steps):
        """barrier + synchronize on both sides, CUDA events on the current stream, MAX over ranks"""
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        outs = [fn(i) for i in range(steps)]
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dist.barrier()
        # the library runs on its own stream and every phase ends with a stream sync, so wall == device span here;
        # take the larger of the two clocks, then the max over ranks
        t = torch.tensor([max(wall, e0.elapsed_time(e1) * 1e-3)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), outs

    # timed region: no per-stage events, stages 1-5 replayed from one captured CUDA graph per step
    use_graph = os.environ.get("GSC_BENCH_GRAPH", "1") == "1"
    ctx.set_flags(timing=False, graph=use_graph)
    slot_order = not getattr(args, "assign_order", False)
    if slot_order:
        # temporal coherence: after one step every rank keeps its colliders in that frame's cell order and
        # marshals the later frames in slot order (gsc_reorder_colliders); the pair set does not change
        step_resident(0)
        order = torch.from_numpy(ctx.reorder_colliders().astype(np.int64))
        part["box"] = part["box"][order.numpy()]
        for f in range(nframes):
            part["h_pos"][f].copy_(part["h_pos"][f][order])
            part["h_quat"][f].copy_(part["h_quat"][f][order])
        part["h_scale"] = part["h_scale"][order].contiguous().pin_memory()
        finish_frames()
        dist.barrier()
    h_pos, h_quat, h_bquat, h_bscale, n_box = part["h_pos"], part["h_quat"], part["h_bquat"], part["h_bscale"], part["n_box"]
    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    span, _ = timed(lambda i: step_resident(args.warmup + i), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = n * args.steps / span
    # the same steps with a CUDA event between the stages (plain launches) for the per-stage table
    ctx.set_flags(timing=True, graph=False)
    step_resident(0)
    _, outs = timed(lambda i: step_resident(args.warmup + i), args.steps)
    ctx.set_flags(timing=False, graph=use_graph)

    # e2e: host transforms in, pairs out to pinned host memory, every step
    max_pairs = int(max(o.n_pairs for o in outs))
    h_pairs = torch.empty((max(max_pairs, 1) * 2 + 4096,), dtype=torch.int32, pin_memory=True)
    import ctypes as C
    npairs = C.c_uint64(0)

    def step_e2e(i):
        f = i % nframes
        ctx.update_transforms_packed_ptr(h_pos[f].data_ptr(), n_box, h_bquat[f].data_ptr(), 0)
        out = driver.step()
        rc = ctx._lib.gsc_download_pairs(ctx._h, h_pairs.data_ptr(), h_pairs.numel() // 2, 0, C.byref(npairs))
        assert rc == 0, ctx._lib.gsc_last_error(ctx._h)
        return out

    ctx.update_transforms_packed_ptr(h_pos[0].data_ptr(), n_box, h_bquat[0].data_ptr(), h_bscale.data_ptr())
    for i in range(args.warmup):
        step_e2e(i)
    span_sync, _ = timed(lambda i: step_e2e(args.warmup + i), args.steps)

    # the same per-step work as a host loop pipelines it: frame f+1 is handed over as soon as bvh_update of step f is
    # done (SlabDriver.step(after_bounds=...)), so its PCIe copy runs under the exchange and stages 1-5 of step f
    def step_e2e_pipelined(i):
        f = (i + 1) % nframes
        out = driver.step(after_bounds=lambda: ctx.update_transforms_packed_ptr(h_pos[f].data_ptr(), n_box, h_bquat[f].data_ptr(), 0))
        rc = ctx._lib.gsc_download_pairs(ctx._h, h_pairs.data_ptr(), h_pairs.numel() // 2, 0, C.byref(npairs))
        assert rc == 0, ctx._lib.gsc_last_error(ctx._h)
        return out

    ctx.update_transforms_packed_ptr(h_pos[0].data_ptr(), n_box, h_bquat[0].data_ptr(), 0)
    for i in range(args.warmup):
        step_e2e_pipelined(i)
    span_e2e, outs_e2e = timed(lambda i: step_e2e_pipelined(args.warmup + i), args.steps)

    # aggregate counters over ranks
    def total(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    stage_ms = np.mean(np.array([list(o.stage_ms) for o in outs], dtype=np.float64), axis=0)
    st = torch.tensor(stage_ms, dtype=torch.float64, device=dev)
    st_min = st.clone()
    dist.all_reduce(st, op=dist.ReduceOp.MAX)
    dist.all_reduce(st_min, op=dist.ReduceOp.MIN)
    stage_ms = st.cpu().numpy()
    stage_min = st_min.cpu().numpy()
    cnt = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(cnt, torch.tensor([n_local], dtype=torch.int64, device=dev))
    counts = [int(c.item()) for c in cnt]
    sums = total([np.mean([o.n_pairs for o in outs]), np.mean([o.n_sphere_obb for o in outs]),
                  np.mean([o.n_obb_obb for o in outs]), np.mean([o.n_sphere_sphere for o in outs]),
                  np.mean([o.n_cells for o in outs]), float(driver.last_halo_in), float(n_local),
                  np.mean([o.n_pairs for o in outs_e2e]) * 8.0, float(n_box)])
    passes = int(outs[-1].sort_passes)
    # a fingerprint of frame 0's pair set over ALL ranks (count, order-independent checksum mod 2^64): it must equal the
    # 1-GPU line's config.frame0_pair_set -- the union of the ranks' lists is the single-GPU set
    step_resident(0)
    pk = ctx.pairs_packed(sorted=False)
    fp = torch.tensor([int(pk.shape[0]), int(np.sum(pk, dtype=np.uint64).astype(np.int64))], dtype=torch.int64, device=dev)
    dist.all_reduce(fp, op=dist.ReduceOp.SUM)
    fp = fp.cpu().numpy()
    fingerprint = {"pairs": int(fp[0]), "checksum": f"{int(fp[1].astype(np.uint64)):016x}"}
    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        n_tot = sums[6] + sums[5]  # volumes processed including ghosts
        sb = stage_bytes(n_tot, f_box, sums[4], passes, sums[0], sums[1], sums[2], sums[3])
        dom = int(np.argmax(stage_ms))
        dom_name = N.STAGE_NAMES[dom]
        # per-rank kernel: bytes of ONE rank's launch (mean) over that kernel's time (max over ranks)
        achieved = sb[dom_name] / world / (stage_ms[dom] * 1e-3) / 1e9
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * span / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.config, n), "frames_cycled": nframes,
                       "partition": f"{world} {'work' if weight is not None else 'count'}-balanced x-slabs, one-cell halo (+x direction), " +
                                    (("pushed into peer HBM over NVLink (CUDA IPC); bounds, flags and ghost counts also travel through "
                                      "peer memory, one host wait per step" if driver.device_exchange else
                                      "pushed into peer HBM over NVLink (CUDA IPC); host-driven control (shared-memory file)")
                                     if driver.push else "NCCL send/recv"),
                       "l2": "inputs larger than L2 (no flush needed)" if n_local * 100 > 126e6 else "per-rank inputs near L2 size",
                       "box_fraction": round(f_box, 3), "contact_pairs_per_step": int(sums[0]),
                       "halo_colliders_per_step": int(sums[5]), "frame0_pair_set": fingerprint,
                       "slot_order": slot_order, "graph_replay": use_graph,
                       "slab_calibration_busy_ms": calib_log, "colliders_per_rank": [int(x) for x in counts]},
            "e2e": {"value": n * args.steps / span_e2e, "unit": unit, "ms_per_step": 1e3 * span_e2e / args.steps,
                    "h2d_bytes_per_step": int(n * 12 + sums[8] * 16), "d2h_bytes_per_step": int(sums[7]),
                    "mode": "pipelined host loop: every rank hands frame f+1 over right after bvh_update of step f",
                    "sync_value": n * args.steps / span_sync, "sync_ms_per_step": 1e3 * span_sync / args.steps},
            "gpu_launches": (23 if driver.device_exchange else 13 + 4 + 4) * args.steps * world,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_kind": peak_kind,
                         "note": "per-rank launch: mean bytes per rank over the slowest rank's stage time"},
            "stages": {N.STAGE_NAMES[i]: {"ms_max_over_ranks": round(float(stage_ms[i]), 4),
                                          "ms_min_over_ranks": round(float(stage_min[i]), 4)} for i in range(6)},
            "stages_note": "the cell_keys interval holds the device-side exchange and, with it, the wait for the slowest rank",
        }
        emit(line)
    if driver.trace is not None and rank == 0:
        total_steps = 2 * (args.warmup + args.steps)
        print("slab trace (us/step, rank 0):", {k: round(1e6 * v / total_steps, 1) for k, v in driver.trace.items()}, file=sys.stderr)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0

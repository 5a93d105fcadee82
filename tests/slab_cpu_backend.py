"""CPU stand-in for one rank's device work behind gunship_rs_b200.slabs.SlabDriver -- TEST ONLY.

It lets the multi-rank host logic (bounds agreement, home x-ranges, who needs which halo, the exchange itself,
taker ownership of pairs) run with gloo on machines without a GPU.  Shapes / AABBs come from the oracle's
bvh_update and the shape tests from the oracle's batch functions; the pair rule is the one the CUDA kernels
implement (gsc_kernels.cuh: taker = later in (home cell key, global index) order)."""
import numpy as np
import torch

import oracle_lib as O
from gunship_rs_b200.slabs import RECORD_BYTES

# 96-byte halo record of this backend (opaque to the driver, like the CUDA one)
REC = np.dtype([("entity", "<u4"), ("gidx", "<u4"), ("kind", "<u4"), ("aabb", "<f4", 6), ("shape", "<f4", 15)])
assert REC.itemsize == RECORD_BYTES


class StepResult:
    def __init__(self, pairs):
        self.pairs = pairs
        self.n_pairs = len(pairs)
        self.n_wide = 0


class CpuSlabBackend:
    tensor_device = torch.device("cpu")

    def __init__(self, entity, kind, offset, size, gidx):
        self.entity, self.kind, self.offset, self.size = entity, kind, offset, size
        self.gidx = gidx.astype(np.uint32)
        self.oracle = O.OracleSystem()
        self.window = None

    def set_frame(self, pos, quat, scale):
        self.pos, self.quat, self.scale = pos, quat, scale

    def _records(self):
        self.oracle.bvh_update(self.entity, self.kind, self.offset, self.size, self.pos, self.quat, self.scale)
        v = self.oracle.volumes()
        r = np.zeros(v.shape[0], dtype=REC)
        r["entity"], r["gidx"], r["kind"] = v["entity"], self.gidx, v["tag"]
        r["aabb"][:, :3], r["aabb"][:, 3:] = v["aabb_min"][:, :3], v["aabb_max"][:, :3]
        pay = v["payload"]
        box = v["tag"] == 1
        r["shape"][~box, 0:3] = pay[~box, 0:3]          # sphere: centre xyz(w), radius
        r["shape"][~box, 3] = pay[~box, 4]
        r["shape"][box, 0:3] = pay[box, 0:3]             # box: centre xyz(w), orientation 9, half widths 3
        r["shape"][box, 3:15] = pay[box, 4:16]
        return r

    def phase_bounds(self):
        self.local = self._records()
        self.ghosts = []
        self.window = None
        if len(self.local) == 0:
            inf = float("inf")
            return 0.0, [inf] * 3, [-inf] * 3, 0, -inf, 0
        a = self.local["aabb"]
        return (float(self.oracle.longest_axis), a[:, :3].min(axis=0).tolist(), a[:, 3:].max(axis=0).tolist(), 0,
                float(a[:, 0].max()), len(self.local))

    def set_global_bounds(self, longest, wmin, wmax, status):
        self.cell = np.float32(longest)

    def set_x_window(self, lo, hi):
        self.window = (lo, hi)

    def _home(self, rec):
        return np.floor(rec["aabb"][:, :3] / self.cell).astype(np.int64)

    def halo_select(self, lo, hi):
        hx = self._home(self.local)[:, 0]
        sel = self.local[(hx >= lo) & (hx <= hi)]
        return len(sel), torch.from_numpy(np.frombuffer(sel.tobytes() + b"\0", dtype=np.uint8).copy())

    def new_buffer(self, count):
        return torch.empty(max(count, 1) * RECORD_BYTES, dtype=torch.uint8)

    def halo_append(self, count, buf):
        self.ghosts.append(np.frombuffer(buf.numpy()[: count * RECORD_BYTES].tobytes(), dtype=REC))

    def sync(self):
        pass

    def phase_finish(self):
        A = self.local
        B = np.concatenate([self.local] + self.ghosts) if self.ghosts else self.local
        hA, hB = self._home(A), self._home(B)
        if self.window is not None and len(A):
            assert hA[:, 0].min() >= self.window[0] + 1 - 1 and hA[:, 0].max() <= self.window[1]
        out = []
        for i0 in range(0, len(A), 256):
            a, ha = A[i0:i0 + 256], hA[i0:i0 + 256]
            # taker rule: A later than B in (home cell, global index) lexicographic order
            later = np.zeros((len(a), len(B)), dtype=bool)
            tie = np.ones_like(later)
            for ax in range(3):
                d = ha[:, None, ax] - hB[None, :, ax]
                later |= tie & (d > 0)
                tie &= d == 0
            later |= tie & (a["gidx"][:, None] > B["gidx"][None, :])
            aa, bb = a["aabb"], B["aabb"]
            ov = np.ones_like(later)
            for ax in range(3):  # inclusive AABB overlap (bounding_volume.rs:411-419)
                ov &= ~((aa[:, None, ax] > bb[None, :, 3 + ax]) | (bb[None, :, ax] > aa[:, None, 3 + ax]))
            ia, ib = np.nonzero(later & ov)
            if len(ia) == 0:
                continue
            ra, rb = a[ia], B[ib]
            a_hi = ra["gidx"] > rb["gidx"]
            hi = np.where(a_hi[:, None], ra.view(np.uint8).reshape(-1, RECORD_BYTES), rb.view(np.uint8).reshape(-1, RECORD_BYTES))
            lo = np.where(a_hi[:, None], rb.view(np.uint8).reshape(-1, RECORD_BYTES), ra.view(np.uint8).reshape(-1, RECORD_BYTES))
            hi = np.ascontiguousarray(hi).view(REC).reshape(-1)
            lo = np.ascontiguousarray(lo).view(REC).reshape(-1)
            hit = np.zeros(len(hi), dtype=bool)
            hs, ls = hi["kind"] == 0, lo["kind"] == 0
            sph = lambda r: np.ascontiguousarray(r["shape"][:, :4])
            obb = lambda r: np.ascontiguousarray(np.concatenate([r["shape"], np.zeros((len(r), 1), np.float32)], axis=1))
            m = hs & ls
            if m.any():
                hit[m] = O.batch_sphere_sphere(sph(hi[m]), sph(lo[m])) != 0
            m = hs & ~ls
            if m.any():
                hit[m] = O.batch_sphere_obb(sph(hi[m]), obb(lo[m])) != 0
            m = ~hs & ls
            if m.any():
                hit[m] = O.batch_sphere_obb(sph(lo[m]), obb(hi[m])) != 0
            m = ~hs & ~ls
            if m.any():
                hit[m] = O.batch_obb_obb(obb(hi[m]), obb(lo[m])) != 0   # later.test_obb(earlier)
            out.append((hi["entity"][hit].astype(np.uint64) << np.uint64(32)) | lo["entity"][hit].astype(np.uint64))
        pairs = np.sort(np.concatenate(out)) if out else np.zeros(0, dtype=np.uint64)
        return StepResult(pairs)

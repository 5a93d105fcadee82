"""Makes tests/golden/transform_golden.npz: a small random transform forest (inputs stored, not regenerated) with the
oracle's derived values (TransformData::update, component/transform.rs:588-608) and a destroy_immediate sequence with
the resulting dense order (bounding_volume.rs:82-104).  Guards the oracle against drift on the CPU and pins the CUDA
path to the same bytes on the GPU.  Run from the repo root: python tests/golden/make_transform_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402

NOP = 0xFFFFFFFF


def main():
    rng = np.random.default_rng(20)
    rows = [64, 160, 400, 1000]
    begin = np.concatenate([[0], np.cumsum(rows)]).astype(np.uint32)
    parent = [np.full(rows[0], NOP, dtype=np.uint32)]
    for d in range(1, len(rows)):
        parent.append(rng.integers(begin[d - 1], begin[d], size=rows[d]).astype(np.uint32))
    parent = np.concatenate(parent)
    n = parent.shape[0]
    lpos = rng.uniform(-8, 8, size=(n, 3)).astype(np.float32)
    q = rng.normal(size=(n, 4))
    lquat = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    lquat[:8] = [0, 0, 0, 1]
    lscale = rng.uniform(0.5, 2.0, size=(n, 3)).astype(np.float32)
    lpos[:4] = 0.0
    lpos[4:8] = -0.0
    dpos, dquat, dscale, dmat = O.transform_update(begin, parent, lpos, lquat, lscale)
    order = np.arange(1, 501, dtype=np.uint32)
    seq, m = [], 500
    for _ in range(300):
        seq.append(int(rng.integers(0, m)))
        m -= 1
    seq = np.asarray(seq, dtype=np.uint32)
    after = O.swap_remove_sequence(order, seq)
    path = os.path.join(HERE, "transform_golden.npz")
    np.savez_compressed(path, row_begin=begin, parent=parent, lpos=lpos, lquat=lquat, lscale=lscale, dpos=dpos, dquat=dquat,
                        dscale=dscale, dmat=dmat, remove_seq=seq, order_after=after)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

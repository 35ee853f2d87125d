"""CPU tests of the N>1 path: feature partitioning and the all-gather plumbing with world_size 2 over gloo.
The compute backend here is built from the CPU oracle (test infrastructure); on the GPU box the same `sharded_update`
runs with the CUDA engine backend (see tests/test_gpu_parity.py::test_sharded_update_matches_single and bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from open_vins_b200 import capi, multigpu, sim


def test_partition_is_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 8):
        M = rng.integers(0, 43, size=400)
        off = np.concatenate([[0], np.cumsum(M)])
        parts = multigpu.partition_features(off, world)
        assert parts[0][0] == 0 and parts[-1][1] == 400
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        rows = np.maximum(2 * M - 3, 0)
        loads = [rows[a:b].sum() for a, b in parts]
        assert max(loads) - min(loads) <= 2 * rows.max() + 1
        assert capi.shard_partition(off, world) == parts  # the C implementation (ovb_shard_partition) agrees with the numpy twin
    # degenerate: fewer features than ranks
    parts = multigpu.partition_features([0, 5, 9], 4)
    assert parts[0][0] == 0 and parts[-1][1] == 2 and all(a <= b for a, b in parts)
    assert capi.shard_partition([0, 5, 9], 4) == parts


class OracleBackend:
    """Stands in for the CUDA engine: per-shard Jacobians + gate + Givens compression from the oracle, on CPU tensors."""

    def __init__(self, oracle, P, layout):
        self.o, self.P, self.layout = oracle, np.array(P), layout

    def partition(self, meas_off, world):
        return capi.shard_partition(meas_off, world)  # the C partition (host code of libovb200.so)

    def shard_compress(self, frame, feats, f0, f1, opts, world):
        feats = feats.subset(np.arange(f0, f1))
        cols = []
        for off, sz in sorted([(o, 6) for o in self.layout.clone_off] + [(o, 6) for o in self.layout.cam_ext_off if o >= 0] +
                              [(o, 8) for o in self.layout.cam_intr_off if o >= 0]):
            cols += list(range(off, off + sz))
        self.cols = cols
        n = len(cols)
        tri, _ = self.o.triangulate(frame, feats, opts)
        self.out = tri
        _, Hx, res, _ = self.o.feature_jacobians(frame, feats, opts, tri, 1, cols, P=self.P)
        R = np.zeros((n, n))
        z = np.zeros(n)
        if Hx.shape[0] > 0:
            Rc, zc = self.o.compress(Hx, res)
            R[:Rc.shape[0]] = Rc
            z[:len(zc)] = zc
        self.n = n
        return torch.from_numpy(np.concatenate([R, z[:, None]], axis=1).copy())

    def gather_target(self, world):
        return torch.empty((world * self.n, self.n + 1), dtype=torch.float64)

    def finish(self, stacked, world, n_feats):
        S = stacked.numpy()
        R, z = self.o.compress(S[:, :-1], S[:, -1])
        st, Pn, dx = self.o.ekf_update(self.P, self.cols, [1] * len(self.cols), R, z, sigma2=1.0)
        self.P_new = Pn
        return st, self.out, dx, None


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ovo_py
    case = sim.make_update_case(n_feats=40, n_clones=8, n_cams=2, seed=5, calib_ext=True, calib_intr=True)
    opts = capi.default_opts(do_calib_camera_pose=1, do_calib_camera_intrinsics=1, col_order=capi.COLS_CANONICAL)
    be = OracleBackend(ovo_py, case.P, case.layout)
    st, out, dx, _, (f0, f1) = multigpu.sharded_update(be, dist, case.frame, case.feats, opts, rank, world, replicate_below_rows=0)
    q.put((rank, st, dx, be.P_new, out.status.copy(), f0, f1))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_update_matches_monolithic_gloo(oracle):
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = sim.make_update_case(n_feats=40, n_clones=8, n_cams=2, seed=5, calib_ext=True, calib_intr=True)
    opts = capi.default_opts(do_calib_camera_pose=1, do_calib_camera_intrinsics=1, col_order=capi.COLS_CANONICAL)
    ref = oracle.msckf_update(case.frame, case.feats, opts, case.P, dumps=False)
    # both ranks end with the same state, equal to the single-process update
    assert np.array_equal(res[0][3], res[1][3]) and np.array_equal(res[0][2], res[1][2])
    assert np.linalg.norm(res[0][3] - ref["P"]) <= 1e-9 * np.linalg.norm(ref["P"])
    assert np.linalg.norm(res[0][2] - ref["dx"]) <= 1e-9 * np.linalg.norm(ref["dx"])
    status = np.concatenate([res[0][4], res[1][4]])
    assert res[0][5] == 0 and res[0][6] == res[1][5] and res[1][6] == case.feats.n_feats
    assert np.array_equal(status, ref["out"].status)

"""Feature-sharded MSCKF update across ranks (one process per GPU, torch.distributed for the plumbing).

SURVEY.md §8e: triangulation, Jacobians, nullspace projection, the chi² gate and the first compression level are
independent per feature, so features are partitioned across ranks (balanced by stacked rows); the ranks exchange their
compressed blocks [R_g | z_g] (n x (n+1) doubles each) with ONE all-gather per update; every rank then compresses the
G stacked triangles and applies the identical EKF update to its replica of P. No other collective is on the data path.

Small updates are NOT sharded: one config-2 update (400 features) is ~0.6 ms of latency-bound kernels, and splitting it
only adds an all-gather and a second compression. Below REPLICATE_BELOW_ROWS stacked rows every rank simply runs the
whole update on its replica (deterministic kernels keep the replicas bitwise equal; no collective at all).

The compute backend is abstract so the plumbing can be exercised on CPU (gloo) in tests; the product backend is
`EngineBackend` (the CUDA engine through the C ABI).
"""
from __future__ import annotations

import contextlib
import json
import time

import numpy as np

REPLICATE_BELOW_ROWS = 60_000  # ~1000 features of the rpng_sim stereo track-length mix


def stacked_rows(meas_off) -> int:
    meas_off = np.asarray(meas_off, dtype=np.int64)
    M = meas_off[1:] - meas_off[:-1]
    return int(np.maximum(2 * M - 3, 0).sum())


def partition_features(meas_off, world: int):
    """Contiguous feature ranges with (nearly) equal stacked-row counts sum(max(2M-3,0)). Returns [(f0, f1)] * world.
    numpy twin of ovb_shard_partition (csrc/ovb_api.cu), used by the CPU test backend."""
    meas_off = np.asarray(meas_off, dtype=np.int64)
    M = meas_off[1:] - meas_off[:-1]
    rows = np.maximum(2 * M - 3, 0)
    F = len(M)
    cum = np.concatenate([[0], np.cumsum(rows)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        f = int(np.searchsorted(cum, target, side="left"))
        f = min(max(f, bounds[-1]), F)
        bounds.append(f)
    bounds.append(F)
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


class EngineBackend:
    """CUDA engine + torch device tensors for the exchanged blocks. Engine kernels and the NCCL all-gather are issued on
    ONE dedicated, explicit CUDA stream (never the legacy default stream: the engine's own streams are non-blocking and
    have no implicit ordering with it), so shard_compress -> all_gather -> finish is ordered without host syncs."""

    def __init__(self, engine, device):
        import torch
        self.torch = torch
        self.eng = engine
        self.device = device
        self.R_local = None
        self.R_all = None
        self.stream = torch.cuda.Stream(device)
        self.eng.set_stream(self.stream.cuda_stream)

    def stream_ctx(self):
        return self.torch.cuda.stream(self.stream)

    def partition(self, meas_off, world):
        from . import capi
        return capi.shard_partition(meas_off, world)

    def full_update(self, frame, feats, opts):
        return self.eng.msckf_update(frame, feats, opts)

    def shard_compress(self, frame, feats, f0, f1, opts, world):
        torch = self.torch
        cap = 512 * 520
        if self.R_local is None:
            self.R_local = torch.empty(cap, dtype=torch.float64, device=self.device)
            self.R_all = torch.empty(cap * world, dtype=torch.float64, device=self.device)
        n, ld = self.eng.shard_compress_range(frame, feats, f0, f1, opts, self.R_local.data_ptr(), cap)
        self.n, self.ld = n, ld
        return self.R_local[: n * ld]

    def gather_target(self, world):
        return self.R_all[: world * self.n * self.ld]

    def finish(self, stacked, world, n_feats):
        return self.eng.shard_finish(stacked.data_ptr(), world, n_feats)


def sharded_update(backend, dist, frame, feats, opts, rank: int, world: int, replicate_below_rows: int = REPLICATE_BELOW_ROWS):
    """One MSCKF update on `world` ranks. Returns (status, out, dx, stats, (f0, f1)); (f0, f1) is the feature range whose
    per-feature results `out` holds (the whole batch when the update was replicated)."""
    F = len(feats.meas_off) - 1
    if world == 1 or stacked_rows(feats.meas_off) < replicate_below_rows:
        st, out, dx, stats = backend.full_update(frame, feats, opts)
        return st, out, dx, stats, (0, F)
    f0, f1 = backend.partition(feats.meas_off, world)[rank]
    with getattr(backend, "stream_ctx", contextlib.nullcontext)():
        block = backend.shard_compress(frame, feats, f0, f1, opts, world)
        stacked = backend.gather_target(world)
        dist.all_gather_into_tensor(stacked, block)
        st, out, dx, stats = backend.finish(stacked, world, f1 - f0)
    return st, out, dx, stats, (f0, f1)


def _relerr(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def time_sharded(eng, backend, dist, torch, case, opts, rank, world, K, W, threshold):
    """K timed sharded updates of `case` (after W warm-ups). Device time per step = CUDA events on the shared stream from
    'inputs resident' to 'EKF update done' (all-gather inside), host time = wall clock around the three calls.
    Returns dict(t_dev, t_host, used, sharded_vs_single) with times as max over ranks."""
    dev = backend.device
    # correctness first (untimed): the sharded result equals the single-GPU result of the same engine on the same inputs
    eng.cov_set(case.P)
    st1, out1, dx1, _ = eng.msckf_update(case.frame, case.feats, opts)
    P1 = eng.cov_get()
    eng.cov_set(case.P)
    st, out, dx, stats, (f0, f1) = sharded_update(backend, dist, case.frame, case.feats, opts, rank, world, threshold)
    torch.cuda.synchronize()
    P2 = eng.cov_get()
    err = max(_relerr(P2, P1), _relerr(dx, dx1))
    same_gate = bool(np.array_equal(out.status, out1.status[f0:f1]))
    for _ in range(W):
        eng.cov_set(case.P)
        sharded_update(backend, dist, case.frame, case.feats, opts, rank, world, threshold)
    torch.cuda.synchronize()
    dist.barrier()
    t_host = t_dev = 0.0
    used = 0
    for _ in range(K):
        eng.cov_set(case.P)
        torch.cuda.synchronize()
        dist.barrier()
        t = time.perf_counter()
        st, out, dx, stats, _ = sharded_update(backend, dist, case.frame, case.feats, opts, rank, world, threshold)
        t_host += time.perf_counter() - t
        t_dev += eng.last_stage_ms()[0] * 1e-3
        used = stats.n_feats_used
    tt = torch.tensor([t_host, t_dev, err], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    used_t = torch.tensor([used, int(same_gate)], dtype=torch.int64, device=dev)
    dist.all_reduce(used_t, op=dist.ReduceOp.SUM)
    Pchk = torch.from_numpy(eng.cov_get()).to(dev)
    Pmax = Pchk.clone()
    dist.all_reduce(Pmax, op=dist.ReduceOp.MAX)
    return {"t_host": float(tt[0]), "t_dev": float(tt[1]), "sharded_vs_single_relerr": float(tt[2]), "features_used": int(used_t[0]),
            "gate_decisions_equal_on_all_ranks": int(used_t[1]) == world, "replicas_bitwise_equal": bool(torch.equal(Pchk, Pmax))}

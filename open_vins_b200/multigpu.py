"""Feature-sharded MSCKF update across ranks (one process per GPU, torch.distributed for the plumbing).

SURVEY.md §8e: triangulation, Jacobians, nullspace projection, the chi² gate and the first compression level are
independent per feature, so features are partitioned across ranks (balanced by stacked rows); the ranks exchange their
compressed blocks [R_g | z_g] (n x (n+1) doubles each) with ONE all-gather per update; every rank then compresses the
G stacked triangles and applies the identical EKF update to its replica of P. No other collective is on the data path.

The compute backend is abstract so the plumbing can be exercised on CPU (gloo) in tests; the product backend is
`EngineBackend` (the CUDA engine through the C ABI).
"""
from __future__ import annotations

import json
import time

import numpy as np


def partition_features(meas_off, world: int):
    """Contiguous feature ranges with (nearly) equal stacked-row counts sum(max(2M-3,0)). Returns [(f0, f1)] * world."""
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
    """CUDA engine + torch device tensors for the exchanged blocks."""

    def __init__(self, engine, device):
        import torch
        self.torch = torch
        self.eng = engine
        self.device = device
        self.R_local = None
        self.R_all = None
        eng_stream = torch.cuda.current_stream(device).cuda_stream
        self.eng.set_stream(eng_stream)  # kernels and NCCL calls are ordered on one stream: no host sync in between

    def shard_compress(self, frame, feats, opts, world):
        torch = self.torch
        cap = 512 * 520
        if self.R_local is None:
            self.R_local = torch.empty(cap, dtype=torch.float64, device=self.device)
            self.R_all = torch.empty(cap * world, dtype=torch.float64, device=self.device)
        n, ld = self.eng.shard_compress(frame, feats, opts, self.R_local.data_ptr(), cap)
        self.n, self.ld = n, ld
        return self.R_local[: n * ld]

    def gather_target(self, world):
        return self.R_all[: world * self.n * self.ld]

    def finish(self, stacked, world, n_feats):
        return self.eng.shard_finish(stacked.data_ptr(), world, n_feats)


def sharded_update(backend, dist, frame, feats, opts, rank: int, world: int):
    """One MSCKF update with features sharded over `world` ranks. Returns (status, out_shard, dx, stats, (f0, f1))."""
    parts = partition_features(feats.meas_off, world)
    f0, f1 = parts[rank]
    shard = feats.subset(np.arange(f0, f1))
    block = backend.shard_compress(frame, shard, opts, world)
    if world > 1:
        stacked = backend.gather_target(world)
        dist.all_gather_into_tensor(stacked, block)
    else:
        stacked = block
    st, out, dx, stats = backend.finish(stacked, world, shard.n_feats)
    return st, out, dx, stats, (f0, f1)


def bench_sharded(args, rank, local_rank, world, case, opts, workload_name, ClockSampler, peaks, emit=None):
    """bench.py's N>1 leg: the config-2 update with its 400 features sharded over N GPUs (strong scaling)."""
    import torch
    import torch.distributed as dist
    from . import capi

    dev = torch.device("cuda", local_rank)
    cap = max(1024, case.feats.n_feats)
    eng = capi.Engine(max_state=256, max_feats=cap, max_meas=cap * 48, device=local_rank)
    backend = EngineBackend(eng, dev)
    K, W = args.steps, args.warmup
    F = case.feats.n_feats
    sampler = ClockSampler(local_rank)
    for _ in range(W):
        eng.cov_set(case.P)
        sharded_update(backend, dist, case.frame, case.feats, opts, rank, world)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        sampler.start()
    t_host = 0.0
    t_dev = 0.0
    used = 0
    for _ in range(K):
        eng.cov_set(case.P)
        torch.cuda.synchronize()
        dist.barrier()
        t = time.perf_counter()
        st, out, dx, stats, _ = sharded_update(backend, dist, case.frame, case.feats, opts, rank, world)
        t_host += time.perf_counter() - t
        t_dev += eng.last_stage_ms()[5] * 1e-3  # CUDA events: first shard kernel .. results on the host
        used = stats.n_feats_used
    clocks = sampler.stop() if rank == 0 else None
    tt = torch.tensor([t_host, t_dev], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    used_t = torch.tensor([used], dtype=torch.int64, device=dev)
    dist.all_reduce(used_t, op=dist.ReduceOp.SUM)
    # replicas must agree bit for bit
    Pchk = torch.from_numpy(eng.cov_get()).to(dev)
    Pmax = Pchk.clone()
    dist.all_reduce(Pmax, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(Pchk, Pmax))
    cnt = eng.last_counters()
    if rank == 0:
        t_host, t_dev = float(tt[0]), float(tt[1])
        line = {
            "metric": "msckf_updates_per_sec", "value": K / t_dev, "unit": "updates/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "feats_per_sec": F * K / t_dev,
            "config": {"workload": workload_name, "features_in": F, "features_used": int(used_t[0]), "sharding": f"features over {world} ranks, "
                       "one all-gather of the compressed (R,z) block per update, EKF update replicated", "l2": "inputs re-uploaded every step",
                       "replicas_bitwise_equal": same},
            "e2e": {"value": K / t_host, "unit": "updates/s", "ms_per_step": 1e3 * t_host / K, "h2d_bytes_per_step": cnt["h2d_bytes"],
                    "d2h_bytes_per_step": cnt["d2h_bytes"], "timing": "host clock around shard_compress + all_gather + finish, max over ranks"},
            "gpu_launches": cnt["launches"] * K, "gpu_launches_per_step": cnt["launches"],
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": None, "peak": peaks()[0], "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "see the N=1 line: the per-rank kernels are the same; at N>1 the step is latency-bound (NCCL + second-level QR)"},
        }
        (emit or (lambda l: print(json.dumps(l), flush=True)))(line)
    eng.close()
    dist.destroy_process_group()

"""Multi-GPU plumbing: one process per GPU, poses sharded contiguously, no collective inside the
loop; one gather of the results (and of the per-pose errors) at the end.

Sharding rule = lib/dataset/EvaSampler.py:79-112 of the reference (contiguous chunks, the first
``N % W`` ranks get one extra item).  The gather works with the NCCL backend (CUDA tensors, NVLink)
and with gloo (CPU tensors; used by the world_size-2 tests).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .engine import shard_range


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; initialises the process group when
    WORLD_SIZE > 1 (MASTER_ADDR should be 127.0.0.1 on a single node)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All ranks contribute their contiguous row shard ``local`` [n_r, ...]; every rank receives the
    full [n_total, ...] tensor in pose order.  Uneven shards are padded to the largest one."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert local.shape[0] == n_total
        return local
    world = dist.get_world_size()
    sizes = shard_sizes(n_total, world)
    assert local.shape[0] == sizes[dist.get_rank()], "local shard does not follow shard_range()"
    width = max(sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous())
    out = out.reshape((world, width) + tuple(local.shape[1:]))
    return torch.cat([out[r, : sizes[r]] for r in range(world)], dim=0)


def global_batch_mean(per_pose: torch.Tensor) -> torch.Tensor:
    """Mean over the GLOBAL batch of a per-pose quantity held as local shards (one all_reduce of a (sum, count)
    pair).  The Langevin corrector's step size uses batch means of the gradient / noise norms (reference
    sampling.py:281-283), so a sharded run has to take them over all ranks to match the single-process value."""
    acc = torch.stack([per_pose.double().sum(), torch.tensor(float(per_pose.numel()), dtype=torch.float64,
                                                             device=per_pose.device)])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return (acc[0] / acc[1]).to(per_pose.dtype)


def global_sum_(stats: torch.Tensor) -> torch.Tensor:
    """In-place SUM all_reduce of a small device vector over the ranks (no-op in a single process): the
    (sum |score_row|, sum |z_row|, rows) statistics ``ScorePlan.score_stats`` returns become global, so the fused
    Langevin update uses the batch means of the WHOLE batch (reference sampling.py:281-283)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def run_sharded(plan_or_factory, db_2d, K, clusters, cfg, hypo=1, mode=None, gt=None, protocol2=False,
                actions=None, local_shard=False, n_total=None, gather_results=True, **run_kw):
    """The whole job on this rank's shard: slice the (host) arrays by ``shard_range``, copy the shard to the device,
    run IPO + OIL (IPO gradients scaled by the GLOBAL batch: the reference's loss is a mean over the whole batch,
    run/opt_main.py:191), optionally evaluate, and gather with ONE ``all_gather_into_tensor`` per result tensor
    (results [N,S,J,3] f32, err_min [N] f64, argmin [N] i32).  Returns
    (results [N,S,J,3] on every rank -- the local shard when ``gather_results`` is false --,
    (err_min [N], argmin [N]) or None).

    ``plan_or_factory``: a ``ScorePlan`` that is large enough, or a callable ``n_local -> ScorePlan``.
    ``local_shard=True``: the arrays passed ARE this rank's shard (weak scaling: every rank generates its own
    poses) and ``n_total`` is the global number of poses.  ``run_kw`` goes to ``run_pose_optimisation`` (infant
    driver switches)."""
    from . import engine
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if local_shard:
        n = int(n_total if n_total is not None else db_2d.shape[0] * world)
        lo, hi = shard_range(n, rank, world)
        assert hi - lo == db_2d.shape[0], "local shard does not follow shard_range()"
        sl = slice(None)
    else:
        n = db_2d.shape[0]
        lo, hi = shard_range(n, rank, world)
        sl = slice(lo, hi)
    dev = torch.device("cuda", torch.cuda.current_device())
    plan = plan_or_factory(hi - lo) if callable(plan_or_factory) else plan_or_factory

    def to_dev(a, dtype=torch.float32):
        t = a if isinstance(a, torch.Tensor) else torch.as_tensor(a)
        return t.to(device=dev, dtype=dtype, non_blocking=True)

    res = engine.run_pose_optimisation(plan, to_dev(db_2d[sl]), to_dev(K[sl]), to_dev(clusters), cfg, hypo=hypo,
                                       mode=mode, b_global=n, **run_kw)
    ev = None
    if gt is not None:
        err, idx = engine.eval_multi(res, to_dev(gt[sl], torch.float64), protocol2=protocol2)
        ev = (gather_rows(err, n), gather_rows(idx, n))
    return (gather_rows(res, n) if gather_results else res), ev

"""Multi-GPU path: pose-sharded run over NCCL == the single-GPU run, bit-exact (needs >= 2 GPUs; skipped on a
one-GPU box -- the sharding/gather logic itself is covered on CPU by tests/test_host_logic.py over gloo)."""
import os
import socket
import sys

import numpy as np
import pytest

import zedo_oracle as zo
from conftest import ROOT

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    import zedo_release_b200 as zr
    from zedo_release_b200 import parallel
    r, w, local = parallel.init_from_env("nccl")
    N, S = 1001, 2  # uneven shards
    ds = zo.make_synthetic_dataset(N, seed=5, n_clusters=S)
    W = zo.make_weights(seed=0)
    cfg = dict(zo.H36M_ZEDO_CFG)
    cfg["OIL_iterations"] = 8
    res, ev = parallel.run_sharded(lambda n: zr.ScorePlan(W, n_joints=17, max_batch=max(n, 1) * S, device=local),
                                   ds["db_2d"], ds["camera_param"], ds["clusters"], cfg, hypo=S,
                                   gt=ds["db_3d"].astype(np.float64), protocol2=True)
    if r == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), res=res.cpu().numpy(), err=ev[0].cpu().numpy(),
                 idx=ev[1].cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_run_equals_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import zedo_release_b200 as zr
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "sharded.npz")
    N, S = 1001, 2
    ds = zo.make_synthetic_dataset(N, seed=5, n_clusters=S)
    cfg = dict(zo.H36M_ZEDO_CFG)
    cfg["OIL_iterations"] = 8
    dev = torch.device("cuda:0")
    plan = zr.ScorePlan(zo.make_weights(seed=0), n_joints=17, max_batch=N * S, device=0)
    res = zr.run_pose_optimisation(plan, torch.tensor(ds["db_2d"], device=dev), torch.tensor(ds["camera_param"], device=dev),
                                   torch.tensor(ds["clusters"], device=dev), cfg, hypo=S)
    err, idx = zr.eval_multi(res, torch.tensor(ds["db_3d"], dtype=torch.float64, device=dev), protocol2=True)
    plan.close()
    assert np.array_equal(got["res"], res.cpu().numpy())  # rows are independent: sharding changes nothing
    assert np.array_equal(got["idx"], idx.cpu().numpy()) and np.array_equal(got["err"], err.cpu().numpy())


def test_plan_on_a_non_current_device():
    """One process, two devices: a plan built for cuda:1 while cuda:0 is current leaves the current device alone,
    computes on cuda:1 through the engine, and the raw C ABI refuses a call made with the wrong device current."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import ctypes as C
    import zedo_release_b200 as zr
    from zedo_release_b200 import _native as nat
    torch.cuda.set_device(0)
    W = zo.make_weights(seed=0)
    x = np.random.default_rng(0).normal(0, 0.3, (300, 17, 3)).astype(np.float32)
    p0 = zr.ScorePlan(W, n_joints=17, max_batch=300, device=0)
    p1 = zr.ScorePlan(W, n_joints=17, max_batch=300, device=1)
    assert torch.cuda.current_device() == 0
    y0 = p0.forward(torch.tensor(x, device="cuda:0"), 50.0)
    x1 = torch.tensor(x, device="cuda:1")
    y1 = p1.forward(x1, 50.0)
    assert y1.device.index == 1 and torch.cuda.current_device() == 0
    assert torch.equal(y0.cpu(), y1.cpu())
    ds = zo.make_synthetic_dataset(300, seed=1)
    g1, T1 = zr.grad_field(torch.tensor(ds["db_2d"][:, :, :2], device="cuda:1"), x1,
                           torch.tensor(ds["camera_param"], device="cuda:1"))
    g0, T0 = zr.grad_field(torch.tensor(ds["db_2d"][:, :, :2], device="cuda:0"), torch.tensor(x, device="cuda:0"),
                           torch.tensor(ds["camera_param"], device="cuda:0"))
    assert g1.device.index == 1 and torch.equal(g0.cpu(), g1.cpu()) and torch.equal(T0.cpu(), T1.cpu())
    out = torch.empty_like(x1)
    rc = nat.lib.zedo_score_forward(p1._h, C.c_void_p(x1.data_ptr()), 50.0, C.c_void_p(out.data_ptr()), 300, 0,
                                    C.c_void_p(0))  # cuda:0 is current
    assert rc == -5  # ZEDO_E_STATE (include/zedo_b200.h)
    p0.close()
    p1.close()
    assert torch.cuda.current_device() == 0


def _langevin_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    import zedo_release_b200 as zr
    from zedo_release_b200 import parallel
    r, w, local = parallel.init_from_env("nccl")
    dev = torch.device("cuda", local)
    N = 1001
    rng = np.random.default_rng(9)
    x = rng.normal(0, 0.4, (N, 17, 3)).astype(np.float32)
    z = rng.normal(0, 1.0, (N, 17, 3)).astype(np.float32)
    lo, hi = zr.shard_range(N, r, w)
    plan = zr.ScorePlan(zo.make_weights(seed=0), n_joints=17, max_batch=hi - lo, device=local)
    xs, zs = torch.tensor(x[lo:hi], device=dev), torch.tensor(z[lo:hi], device=dev)
    stats = plan.score_stats(xs, 43.7, z=zs, std_div=0.31, want_stats=True)
    parallel.global_sum_(stats)  # NCCL all_reduce of (sum |score_row|, sum |z_row|, rows)
    xn, xm = plan.noise_update("langevin", xs, zs, 0.31, 0.16, 0.998, stats=stats)
    full = parallel.gather_rows(xn, N)
    if r == 0:
        np.savez(os.path.join(out_dir, "langevin.npz"), x=full.cpu().numpy(), stats=stats.cpu().numpy())
    plan.close()
    dist.barrier()
    dist.destroy_process_group()


def test_langevin_step_size_uses_the_global_batch_mean(tmp_path):
    """sampling.py:281-283: the Langevin step size takes BATCH means of the gradient / noise norms; sharded over two
    ranks the per-rank sums are all-reduced over NCCL between the two fused calls and the result equals the
    single-GPU run on the whole batch."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import zedo_release_b200 as zr
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_langevin_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "langevin.npz")
    N = 1001
    rng = np.random.default_rng(9)
    x = torch.tensor(rng.normal(0, 0.4, (N, 17, 3)).astype(np.float32), device="cuda:0")
    z = torch.tensor(rng.normal(0, 1.0, (N, 17, 3)).astype(np.float32), device="cuda:0")
    plan = zr.ScorePlan(zo.make_weights(seed=0), n_joints=17, max_batch=N, device=0)
    stats = plan.score_stats(x, 43.7, z=z, std_div=0.31, want_stats=True)
    xn, _ = plan.noise_update("langevin", x, z, 0.31, 0.16, 0.998, stats=stats)
    plan.close()
    assert got["stats"][2] == N and np.allclose(got["stats"], stats.cpu().numpy(), rtol=1e-12)
    assert np.allclose(got["x"], xn.cpu().numpy(), rtol=0, atol=1e-6)

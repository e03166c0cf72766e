"""Cluster-pose generation: the producer of ``clusters/{h36m,3dhp}_cluster{S}.npy``.

The reference drivers load S cluster poses as hypothesis initialisations (run/opt_main.py:58-65:
``np.load(f'clusters/h36m_cluster{args.hypo}.npy')``, used root-relative at :167-168); the files are shipped, their
generator is not (run/opt_main_infant.py:25,34 only imports ``scipy.cluster.vq`` / ``sklearn.cluster.KMeans``).  This
module is that generator on the device: Lloyd's k-means over training poses (``zedo_kmeans_fit``), initial centres
drawn from the data with a seeded host generator, output float32 ``[S, J, 3]`` as ``np.load`` expects it.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as nat
from .engine import _device_scoped, _f32, _ptr, _stream


def initial_centres(n_poses: int, n_clusters: int, seed: int = 0) -> np.ndarray:
    """Indices of ``n_clusters`` distinct training poses (seeded ``numpy`` generator) used as the initial centres."""
    if n_clusters > n_poses:
        raise ValueError(f"cannot draw {n_clusters} distinct centres from {n_poses} poses")
    return np.sort(np.random.default_rng(seed).choice(n_poses, size=n_clusters, replace=False))


@_device_scoped
def kmeans_clusters(poses: torch.Tensor, n_clusters: int, iters: int = 50, seed: int = 0,
                    init: Optional[torch.Tensor] = None, root_relative: bool = True
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """poses [N, J, 3] float32 CUDA -> (centres [S, J, 3] float32, labels [N] int32, sq. distances [N] float64).
    ``root_relative`` subtracts joint 0 first (the drivers use ``sample_poses - sample_poses[:, 0:1]``)."""
    poses = _f32(poses, "poses")
    N, J = poses.shape[0], poses.shape[1]
    x = (poses - poses[:, 0:1]) if root_relative else poses
    x = x.reshape(N, J * 3).contiguous()
    if init is None:
        idx = torch.as_tensor(initial_centres(N, n_clusters, seed), device=x.device)
        centres = x[idx].clone()
    else:
        centres = _f32(init, "init").reshape(n_clusters, J * 3).clone()
    labels = torch.empty((N,), dtype=torch.int32, device=x.device)
    dist = torch.empty((N,), dtype=torch.float64, device=x.device)
    nat.check(nat.lib.zedo_kmeans_fit(_ptr(x), N, J * 3, int(n_clusters), int(iters), _ptr(centres), _ptr(labels),
                                      _ptr(dist), _stream()), "zedo_kmeans_fit")
    return centres.reshape(n_clusters, J, 3), labels, dist


def save_cluster_file(path: str, centres: torch.Tensor) -> None:
    """Write the centres the way run/opt_main.py:59 reads them (``np.load`` -> float32 [S, J, 3])."""
    np.save(path, centres.detach().cpu().numpy().astype(np.float32))

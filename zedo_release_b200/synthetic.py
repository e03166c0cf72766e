"""Synthetic H36M-format inputs and random-init weights for benchmarks and examples.

Host-side numpy only (input generation is not part of the hot path).  The same generators live
in ``oracle/zedo_oracle.py`` for the tests; ``tests/test_host_logic.py`` checks that both produce
identical arrays, so golden vectors and benchmark inputs are interchangeable.
Layouts follow the reference's dataset objects: ``db_3d`` [N,J,3] root-relative metres,
``db_2d`` [N,J,3] = (u, v, conf), ``camera_param`` [N,3,3] (run/opt_main.py:115-118).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

f32 = np.float32
Weights = Dict[str, np.ndarray]

H36M_SKELETON = [[0, 1], [1, 2], [2, 3], [0, 4], [4, 5], [5, 6], [0, 7], [7, 8], [8, 9], [9, 10],
                 [8, 11], [11, 12], [12, 13], [8, 14], [14, 15], [15, 16]]  # h36m.py:445-448

_H36M_TEMPLATE = np.array([
    [0.00, 0.00, 0.00], [-0.13, 0.00, 0.00], [-0.13, 0.44, 0.00], [-0.13, 0.88, 0.00],
    [0.13, 0.00, 0.00], [0.13, 0.44, 0.00], [0.13, 0.88, 0.00], [0.00, -0.24, 0.00],
    [0.00, -0.48, 0.00], [0.00, -0.58, 0.00], [0.00, -0.70, 0.00], [0.17, -0.44, 0.00],
    [0.30, -0.20, 0.00], [0.32, 0.04, 0.00], [-0.17, -0.44, 0.00], [-0.30, -0.20, 0.00],
    [-0.32, 0.04, 0.00]], dtype=f32)  # metres, y down (camera frame), root = pelvis


def skeleton_template(n_joints=17):
    if n_joints == 17:
        return _H36M_TEMPLATE.copy()
    rng = np.random.default_rng(99)
    t = rng.normal(0, 0.3, (n_joints, 3)).astype(f32)
    t[0] = 0
    return t


def make_weights(seed=0, n_joints=17, hidden=1024, embed=512, n_blocks=2, control=False) -> Weights:
    """Random-init weights with the reference's state_dict names and the default
    ``nn.Linear`` init range U(-1/sqrt(fan_in), 1/sqrt(fan_in)); GroupNorm affine is
    randomised (U(0.5,1.5), U(-0.2,0.2)) so the affine path is exercised.  Generated with
    numpy's PCG64 so the same seed gives the same weights on every machine."""
    rng = np.random.default_rng(seed)
    D = n_joints * 3
    W: Weights = {}

    def lin(name, fin, fout):
        bound = 1.0 / math.sqrt(fin)
        W[name + ".weight"] = rng.uniform(-bound, bound, (fout, fin)).astype(f32)
        W[name + ".bias"] = rng.uniform(-bound, bound, (fout,)).astype(f32)

    def gn(name):
        W[name + ".weight"] = rng.uniform(0.5, 1.5, (hidden,)).astype(f32)
        W[name + ".bias"] = rng.uniform(-0.2, 0.2, (hidden,)).astype(f32)

    lin("pre_dense", D, hidden)
    lin("pre_dense_t", embed, hidden)
    gn("pre_gnorm")
    lin("shared_time_embed.0", embed, embed)
    for k in range(1, n_blocks + 1):
        lin(f"b{k}_dense1", hidden, hidden)
        lin(f"b{k}_dense1_t", embed, hidden)
        gn(f"b{k}_gnorm1")
        lin(f"b{k}_dense2", hidden, hidden)
        lin(f"b{k}_dense2_t", embed, hidden)
        gn(f"b{k}_gnorm2")
    lin("post_dense", hidden, D)
    if control:
        W["infant_cond"] = rng.normal(0, 1, (D,)).astype(f32)
        lin("zc_layer_1", D, D)
        lin("zc_layer_2", hidden, hidden)
        lin("pre_dense_copy", D, hidden)
        lin("pre_dense_t_copy", embed, hidden)
        gn("pre_gnorm_copy")
        for k in range(1, n_blocks + 1):
            lin(f"zc_b{k}_1", hidden, hidden)
            lin(f"zc_b{k}_2", hidden, hidden)
            lin(f"b{k}_dense1_copy", hidden, hidden)
            lin(f"b{k}_dense1_t_copy", embed, hidden)
            gn(f"b{k}_gnorm1_copy")
            lin(f"b{k}_dense2_copy", hidden, hidden)
            lin(f"b{k}_dense2_t_copy", embed, hidden)
            gn(f"b{k}_gnorm2_copy")
    return W


def make_synthetic_dataset(n_poses, n_joints=17, seed=1234, detected_2d=True, n_clusters=1,
                           dtype_gt=np.float32):
    """Synthetic H36M-format inputs: returns a dict with
    db_3d [N,J,3] (root-relative metres), db_2d [N,J,3]=(u,v,conf), camera_param [N,3,3],
    actions [N] in 2..16, clusters [S,J,3], root [N,3]."""
    rng = np.random.default_rng(seed)
    tmpl = skeleton_template(n_joints)
    ang = rng.uniform(-np.pi, np.pi, n_poses)
    c, s = np.cos(ang), np.sin(ang)
    Ry = np.zeros((n_poses, 3, 3))
    Ry[:, 0, 0], Ry[:, 0, 2], Ry[:, 1, 1], Ry[:, 2, 0], Ry[:, 2, 2] = c, s, 1, -s, c
    gt = np.einsum("bij,nj->bni", Ry, tmpl) + rng.normal(0, 0.05, (n_poses, n_joints, 3))
    gt = gt - gt[:, 0:1]
    root = np.stack([rng.uniform(-1, 1, n_poses), rng.uniform(-1, 1, n_poses),
                     rng.uniform(3, 7, n_poses)], axis=-1)
    cam = gt + root[:, None, :]
    K = np.zeros((n_poses, 3, 3))
    K[:, 0, 0] = 1145.0 + rng.uniform(-5, 5, n_poses)
    K[:, 1, 1] = 1145.0 + rng.uniform(-5, 5, n_poses)
    K[:, 0, 2] = 512.0 + rng.uniform(-4, 4, n_poses)
    K[:, 1, 2] = 515.0 + rng.uniform(-4, 4, n_poses)
    K[:, 2, 2] = 1.0
    proj = np.einsum("bij,bnj->bni", K, cam)
    uv = proj[:, :, :2] / proj[:, :, 2:]
    if detected_2d:
        uv = uv + rng.normal(0, 5.0, uv.shape)
        conf = rng.uniform(0.3, 1.0, (n_poses, n_joints))
    else:
        conf = np.ones((n_poses, n_joints))
    clusters = tmpl[None] + rng.normal(0, 0.1, (n_clusters, n_joints, 3))
    clusters = clusters - clusters[:, 0:1]
    return dict(
        db_3d=gt.astype(dtype_gt),
        db_2d=np.concatenate([uv, conf[..., None]], axis=-1).astype(f32),
        camera_param=K.astype(f32),
        actions=(2 + np.arange(n_poses) % 15).astype(np.int64),
        clusters=clusters.astype(f32),
        root=root.astype(f32),
    )


PW3D_ORDER = [5, 2, 6, 3, 11, 14, 12, 15, 13, 16, 1, 4, 8, 10, 0, 7, 9]  # lib/dataset/pw3d.py:76


def h36m_items_from_arrays(ds, dtype=np.float64):
    """The synthetic dataset as the list of ground-truth items ``h36m_<subset>.pkl`` holds
    (lib/dataset/h36m.py:205-234): ``joint_3d_camera`` [17,3] millimetres (absolute), ``joint_3d_image``
    [17,3] (u, v, depth), ``camera_param`` {fx, fy, cx, cy} as 0-d arrays, ``image_path``, ``action``."""
    items = []
    cam_mm = (ds["db_3d"].astype(np.float64) + ds["root"].astype(np.float64)[:, None, :]) * 1000.0
    for n in range(len(cam_mm)):
        K = ds["camera_param"][n]
        img = np.concatenate([ds["db_2d"][n, :, :2], cam_mm[n, :, 2:3]], axis=-1)
        items.append({"joint_3d_camera": cam_mm[n].astype(dtype), "joint_3d_image": img.astype(dtype),
                      "camera_param": {"fx": np.array(K[0, 0]), "fy": np.array(K[1, 1]), "cx": np.array(K[0, 2]),
                                       "cy": np.array(K[1, 2])},
                      "image_path": f"synthetic/{n:08d}.jpg", "action": int(ds["actions"][n])})
    return items


def pw3d_npz_from_arrays(ds):
    """The synthetic dataset under the keys of ``pw3d_<subset>.npz`` (lib/dataset/pw3d.py:184-199): joints in the
    file's own order (``order_change`` maps file joint i to H36M joint PW3D_ORDER[i]), root-relative + ``root_cam``."""
    N = len(ds["db_3d"])
    rel = np.empty((N, 17, 3), np.float64)
    for i in range(17):
        rel[:, i] = ds["db_3d"][:, PW3D_ORDER[i]]
    K = ds["camera_param"].astype(np.float64)
    return {"keypoints3d17_relative": rel, "root_cam": ds["root"].astype(np.float64),
            "cam_param": np.array({"f": np.stack([K[:, 0, 0], K[:, 1, 1]], -1), "c": np.stack([K[:, 0, 2], K[:, 1, 2]], -1)},
                                  dtype=object),
            "image_width": np.full(N, 1920), "image_height": np.full(N, 1080),
            "image_path": np.array([f"synthetic/{n:08d}.jpg" for n in range(N)])}


H36M_ZEDO_CFG = dict(IPO_iterations=500, IPO_keylist=[0, 1, 4], RotAxes="z", IPO_T=3,
                     IPO_minScaleT=0.5, IPO_maxScaleT=2, OIL_iterations=1000,
                     sampling_eps=0.01)  # configs/optim/concat_pose_optimization_h36m.py:70-81
PW3D_ZEDO_CFG = dict(IPO_iterations=500, IPO_keylist=list(range(17)), RotAxes="z", IPO_T=8,
                     IPO_minScaleT=0.2, IPO_maxScaleT=2, OIL_iterations=1000,
                     sampling_eps=0.01)  # configs/optim/concat_pose_optimization_pw3d.py:72-81
SKI_ZEDO_CFG = dict(IPO_iterations=500, IPO_keylist=list(range(17)), RotAxes="y", IPO_T=20,
                    IPO_minScaleT=0.5, IPO_maxScaleT=2, OIL_iterations=1000,
                    sampling_eps=0.01)  # configs/optim/concat_pose_optimization_ski.py:72-81
MINI_ZEDO_CFG = dict(IPO_iterations=500, IPO_keylist=list(range(17)), RotAxes="xyz", IPO_T=1,
                     IPO_minScaleT=0, IPO_maxScaleT=4, OIL_iterations=1000,
                     sampling_eps=0.01)  # configs/optim/concat_pose_optimization_mini.py:73-85
SYRIP_ZEDO_CFG = dict(IPO_iterations=500, IPO_keylist=list(range(12)), RotAxes="xyz", IPO_T=1,
                      IPO_minScaleT=0.5, IPO_maxScaleT=8, OIL_iterations=1000,
                      sampling_eps=0.01)  # configs/optim/concat_pose_optimization_syrip.py:73-86

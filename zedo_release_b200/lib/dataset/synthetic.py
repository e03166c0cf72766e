"""A dataset object with the attribute contract the drivers rely on (run/opt_main.py:115-118,227-228):
``db_3d`` [N,J,3], ``db_2d`` [N,J,3] = (u, v, conf), ``camera_param`` [N,3,3] and
``eval_multi(preds, protocol2=..., print_verbose=...)``.  File loaders of the reference
(lib/dataset/*.py) are I/O and out of scope; this one is filled from arrays (e.g. the synthetic
H36M-format generator).  ``eval_multi`` follows lib/dataset/h36m.py:365-442 (action-wise mean over
actions 2..16) when ``actions`` is given, else lib/dataset/pw3d.py:286-345 (plain mean), and runs
on the GPU (csrc/eval.cu)."""
import numpy as np
import torch

from zedo_release_b200 import engine
from zedo_release_b200 import synthetic as _syn


class ArrayPoseDataset:
    def __init__(self, db_3d, db_2d, camera_param, actions=None, joint_subset=None, name="synthetic"):
        self.db_3d = np.asarray(db_3d)
        self.db_2d = np.asarray(db_2d, dtype=np.float32)
        self.camera_param = np.asarray(camera_param, dtype=np.float32)
        self.actions = None if actions is None else np.asarray(actions)
        self.joint_subset = joint_subset
        self.name = name
        self.last_index = None  # argmin over hypotheses of the last eval_multi call
        self.gt_eval = None     # root-relative metres in the source's own dtype, when built from dataset items
        self.image_name = None

    @classmethod
    def synthetic_h36m(cls, n_poses, seed=1234, detected_2d=True):
        ds = _syn.make_synthetic_dataset(n_poses, n_joints=17, seed=seed, detected_2d=detected_2d)
        return cls(ds["db_3d"], ds["db_2d"], ds["camera_param"], actions=ds["actions"], name="h36m-synthetic")

    @classmethod
    def from_h36m_items(cls, items, gt2d=True, detections=None, abs_coord=True, name="h36m"):
        """From the ground-truth items of ``h36m_<subset>.pkl`` (the in-memory format of
        lib/dataset/h36m.py:205-263; reading the pickle is the caller's business): ``db_3d`` =
        joint_3d_camera / 1000 (float32, root-relative unless ``abs_coord``), ``camera_param`` from
        fx, fy, cx, cy, ``db_2d`` = joint_3d_image[..., :2] + confidence 1 (``gt2d``) or
        ``detections`` = (joint3d_image [N,17,>=2], confidence [N,17,1]) of ``h36m_sh_dt_ft.pkl``.
        ``eval_multi`` scores against the items' own joint_3d_camera in its stored dtype
        (h36m.py:402-403), action-wise."""
        labels = np.array([it["joint_3d_camera"] for it in items], dtype=np.float32)
        image = np.array([it["joint_3d_image"] for it in items], dtype=np.float32)
        K = np.zeros((len(items), 3, 3), np.float32)
        for n, it in enumerate(items):
            cp = it["camera_param"]
            K[n, 0, 0], K[n, 1, 1] = np.asarray(cp["fx"]).item(), np.asarray(cp["fy"]).item()
            K[n, 0, 2], K[n, 1, 2], K[n, 2, 2] = np.asarray(cp["cx"]).item(), np.asarray(cp["cy"]).item(), 1
        if not abs_coord:
            labels = labels - labels[:, 0:1]
        labels = labels / 1000.0
        if gt2d:
            d2 = np.concatenate((image[..., :2], np.ones((len(items), image.shape[1], 1))), axis=-1)
        else:
            if detections is None:
                raise ValueError("gt2d=False needs detections=(joint3d_image, confidence)")
            d2 = np.concatenate((np.asarray(detections[0])[:, :, :2], np.asarray(detections[1])), axis=-1)
        ds = cls(labels, d2, K, actions=[it["action"] for it in items], name=name)
        gt = np.array([it["joint_3d_camera"] for it in items])
        ds.gt_eval = (gt - gt[:, 0:1]) / 1000.0
        ds.image_name = [it.get("image_path") for it in items]
        return ds

    @classmethod
    def from_pw3d_npz(cls, data, abs_coord=True, name="3dpw"):
        """From the arrays of ``pw3d_<subset>.npz`` (lib/dataset/pw3d.py:177-227): joints re-ordered to the
        H36M topology (``order_change``), absolute = relative + root_cam, K from cam_param f / c, ``db_2d`` =
        the projection (K X) / z whose third column, 1, serves as the confidence."""
        cam = data["cam_param"].item() if hasattr(data["cam_param"], "item") else data["cam_param"]
        rel, root = np.asarray(data["keypoints3d17_relative"]), np.asarray(data["root_cam"])
        N = len(rel)
        X = np.empty((N, 17, 3), np.float64)
        absx = rel[:, :, :3] + root[:, None, :]
        for i in range(17):
            X[:, _syn.PW3D_ORDER[i]] = absx[:, i]
        K = np.zeros((N, 3, 3), np.float64)
        K[:, 0, 0], K[:, 1, 1] = cam["f"][:, 0], cam["f"][:, 1]
        K[:, 0, 2], K[:, 1, 2], K[:, 2, 2] = cam["c"][:, 0], cam["c"][:, 1], 1
        proj = np.einsum("bij,bnj->bni", K, X)
        d2 = proj / proj[:, :, 2:]
        labels = X.astype(np.float32)
        if not abs_coord:
            labels = labels - labels[:, 0:1]
        ds = cls(labels, d2.astype(np.float32), K.astype(np.float32), name=name)
        ds.image_name = list(data["image_path"]) if "image_path" in data else None
        return ds

    def __len__(self):
        return len(self.db_3d)

    def eval_multi(self, preds, protocol2=False, print_verbose=False, sample_interval=None, valid_ind=None):
        """preds [N, m, j, 3] -> scalar error in metres; per pose the minimum over the m hypotheses of
        the mean per-joint error (after Procrustes alignment when protocol2).

        ``valid_ind`` (h36m.py:399-401): per-pose collections of admissible hypothesis indices; the others
        are skipped and ``last_index`` counts inside the kept list, like the reference's ``np.argmin`` over
        its filtered list.  ``sample_interval`` keeps the reference's semantics literally (h36m.py:385-386,
        pw3d.py:296-297): ``preds[::k]`` is compared with the FIRST ``len(preds[::k])`` ground truths, and
        the action-wise H36M aggregate then indexes the shortened result array with full-length indices,
        which raises IndexError in the reference as it does here.
        With ``name='3dhp'`` the 3DHP extras (mpii3dHP.py:480-490) are computed too: ``last_pck``,
        ``last_auc`` on the selected hypotheses and ``last_std`` (diversity)."""
        assert len(preds) == len(self.db_3d)
        gt = self.gt_eval if self.gt_eval is not None else self.db_3d - self.db_3d[:, 0:1]
        actions = self.actions
        if sample_interval is not None:
            preds = preds[::sample_interval]
            gt = gt[:len(preds)]
            if actions is not None and len(preds) != len(actions):
                raise IndexError(f"index {len(actions) - 1} is out of bounds for axis 0 with size {len(preds)} "
                                 "(action-wise aggregate over a sub-sampled result array, h36m.py:430-431)")
        dev = torch.device("cuda", torch.cuda.current_device())
        p = torch.as_tensor(np.ascontiguousarray(preds, dtype=np.float32), device=dev)
        g = torch.as_tensor(np.ascontiguousarray(gt, dtype=np.float64), device=dev)
        if valid_ind is None:
            err, idx = engine.eval_multi(p, g, protocol2=protocol2, joint_subset=self.joint_subset)
            sel = idx
            self.last_index = idx.cpu().numpy()
        else:
            _, _, err_all = engine.eval_multi(p, g, protocol2=protocol2, joint_subset=self.joint_subset,
                                              return_all=True)
            keep = np.zeros(err_all.shape, dtype=bool)
            for n, allowed in enumerate(valid_ind):
                keep[n, [s for s in range(keep.shape[1]) if s in allowed]] = True
            if not keep.any(axis=1).all():
                raise ValueError("attempt to get argmin of an empty sequence")  # np.argmin([]) in the reference
            keep_t = torch.as_tensor(keep, device=dev)
            masked = torch.where(keep_t, err_all, torch.full_like(err_all, float("inf")))
            err, sel = masked.min(dim=1)
            # position of the winner inside the kept list
            self.last_index = (torch.cumsum(keep_t, dim=1).gather(1, sel[:, None])[:, 0] - 1).cpu().numpy()
        error = engine.aggregate_errors(err, actions)
        if self.name == "3dhp":
            self.last_pck, self.last_auc = engine.pck_auc(p, g, select=sel, joint_subset=None)
            self.last_std = engine.hypothesis_std(p)
        if print_verbose:
            print(f"{self.name} {'p2' if protocol2 else 'p1'}: {error:.5f}")
        return error

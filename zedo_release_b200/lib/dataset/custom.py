"""``CustomDataset`` for in-the-wild inputs (reference lib/dataset/custom.py:9-101, used by run/inference.py:118-121).

The reference leaves ``read_data`` as a TODO (custom.py:53-60: "read 2d keypoints [N,17,3] with confidence score, 3d
keypoints [N,17,3] for evaluation only -- can be a zero array for inference --, camera parameters [N,3,3], image
name [N]") and the constructor call in the driver as "your dataset setup".  This class keeps the reference's
constructor (``root_path``, ``sample_interval``), attributes and ``eval_multi`` contract and fills the stub in: the
four arrays are read from ``<root_path>/custom.npz`` (or ``root_path`` itself when it names an ``.npz`` file) with the
keys ``keypoints_2d``, ``keypoints_3d`` (optional), ``camera_params``, ``image_name`` (optional), or handed over in
memory through ``from_arrays``.  Evaluation runs on the device (csrc/eval.cu); the reported value is the plain mean of
the per-pose minimum over hypotheses, printed as the reference prints it (custom.py:95-99).
"""
import os

import numpy as np

from .synthetic import ArrayPoseDataset


class CustomDataset(ArrayPoseDataset):
    def __init__(self, root_path, sample_interval=None):
        self.w = None
        self.h = None
        self.root_path = root_path
        self.sample_interval = sample_interval
        labels_2d, labels_3d, camera_params, image_name = self.read_data()
        super().__init__(labels_3d, labels_2d, camera_params, name="wild")
        self.image_name = image_name
        if self.sample_interval:
            self._sample(sample_interval)
        self.real_data_len = len(self.db_2d)
        self.left_joints = [4, 5, 6, 11, 12, 13]
        self.right_joints = [1, 2, 3, 14, 15, 16]

    @classmethod
    def from_arrays(cls, keypoints_2d, camera_params, keypoints_3d=None, image_name=None, sample_interval=None):
        self = cls.__new__(cls)
        self.w = self.h = None
        self.root_path, self.sample_interval = None, sample_interval
        k2, k3, K, names = cls._validate(keypoints_2d, keypoints_3d, camera_params, image_name)
        ArrayPoseDataset.__init__(self, k3, k2, K, name="wild")
        self.image_name = names
        if sample_interval:
            self._sample(sample_interval)
        self.real_data_len = len(self.db_2d)
        self.left_joints, self.right_joints = [4, 5, 6, 11, 12, 13], [1, 2, 3, 14, 15, 16]
        return self

    @staticmethod
    def _validate(k2, k3, K, names):
        k2 = np.asarray(k2, dtype=np.float32)
        if k2.ndim != 3 or k2.shape[2] not in (2, 3):
            raise ValueError(f"keypoints_2d must be [N, J, 3] = (u, v, confidence) or [N, J, 2]; got {k2.shape}")
        if k2.shape[2] == 2:  # no detector confidence: weight every joint equally
            k2 = np.concatenate((k2, np.ones(k2.shape[:2] + (1,), np.float32)), axis=-1)
        N, J = k2.shape[:2]
        k3 = np.zeros((N, J, 3), np.float32) if k3 is None else np.asarray(k3, dtype=np.float32)
        K = np.asarray(K, dtype=np.float32)
        if K.shape == (3, 3):
            K = np.broadcast_to(K, (N, 3, 3)).copy()
        if k3.shape != (N, J, 3) or K.shape != (N, 3, 3):
            raise ValueError(f"shapes do not agree: 2D {k2.shape}, 3D {k3.shape}, K {K.shape}")
        names = [f"{i:08d}" for i in range(N)] if names is None else [str(s) for s in names]
        return k2, k3, K, names

    def read_data(self):
        path = self.root_path
        if os.path.isdir(str(path)):
            path = os.path.join(str(path), "custom.npz")
        with np.load(str(path), allow_pickle=False) as f:
            return self._validate(f["keypoints_2d"], f["keypoints_3d"] if "keypoints_3d" in f else None,
                                  f["camera_params"], f["image_name"] if "image_name" in f else None)

    def __getitem__(self, idx):
        return self.db_2d[idx % self.real_data_len], self.db_3d[idx % self.real_data_len]

    def __len__(self):
        return len(self.db_2d)

    def _sample(self, sample_interval):
        print(f'Class CustomDataset: sample dataset every {sample_interval} frame')
        self.db_2d = self.db_2d[::sample_interval]
        self.db_3d = self.db_3d[::sample_interval]
        self.camera_param = self.camera_param[::sample_interval]
        self.image_name = self.image_name[::sample_interval]

    def eval_multi(self, preds, protocol2=False, print_verbose=False, sample_interval=None, valid_ind=None):
        print('eval multi-hypothesis...')
        error = super().eval_multi(preds, protocol2=protocol2, print_verbose=False, sample_interval=sample_interval,
                                   valid_ind=valid_ind)
        print(f'mean PA-MPJPE : {error}' if protocol2 else f'mean MPJPE : {error}')
        return error

    @staticmethod
    def get_skeleton():
        return [[0, 1], [1, 2], [2, 3], [0, 4], [4, 5], [5, 6], [0, 7], [7, 8], [8, 9], [9, 10], [8, 11], [11, 12],
                [12, 13], [8, 14], [14, 15], [15, 16]]

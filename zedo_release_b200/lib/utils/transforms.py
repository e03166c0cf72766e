"""``procrustes`` / ``align_to_gt`` with the reference's interface (lib/utils/transforms.py:42-148),
served by the batched Procrustes of csrc/eval.cu (one-sided Jacobi 3x3 SVD in float64)."""
import numpy as np
import torch

from zedo_release_b200 import engine


def align_to_gt(pose, pose_gt):
    """Similarity-align ``pose`` [J,3] to ``pose_gt`` [J,3] (scaling on, reflections allowed because
    the reference applies no determinant fix) and return the aligned pose as float64 numpy."""
    dev = torch.device("cuda", torch.cuda.current_device())
    p = torch.as_tensor(np.asarray(pose, dtype=np.float32), device=dev)[None, None]
    g = torch.as_tensor(np.asarray(pose_gt, dtype=np.float64), device=dev)[None]
    _, _, aligned = engine.eval_multi(p, g, protocol2=True, return_aligned=True)
    return aligned[0, 0].cpu().numpy()


def procrustes(A, B, scaling=True, reflection='best'):
    """MATLAB-style procrustes(A = target, B = input): returns (d, Z, tform) with Z the transformed
    B.  Only the configuration ``align_to_gt`` uses (scaling=True, reflection='best') is implemented;
    ``tform`` is reduced to the entries that follow from Z."""
    if not scaling or reflection != 'best':
        raise NotImplementedError("only scaling=True, reflection='best' (the align_to_gt call) is implemented")
    A = np.asarray(A, dtype=np.float64)
    Z = align_to_gt(B, A)
    A0 = A - A.mean(0)
    d = float(((A - Z) ** 2).sum() / (A0 ** 2).sum())
    return d, Z, {'translation': None, 'rotation': None, 'scale': None}

"""ctypes binding of ``libzedo_b200.so`` (the C ABI declared in ``include/zedo_b200.h``).

There is no CPU fallback: importing this module fails loudly when the shared library has not
been built (``python -c "import __graft_entry__ as g; g.build()"``), and every compute entry
point returns a CUDA error when no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Iterable, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZEDO_B200_LIB: load another build of the same sources (tools/: the -DZEDO_EXPERIMENTS=1 timing build); never a non-CUDA path
LIB_PATH = os.environ.get("ZEDO_B200_LIB") or os.path.join(_HERE, "libzedo_b200.so")

GEMM_SPLIT3, GEMM_FP16, GEMM_FP32, GEMM_SPLIT2, GEMM_FP8LO = 0, 1, 2, 3, 4
NET_SCORE_FC_ADV, NET_CONTROL = 0, 1
PRED_EULER_MARUYAMA, PRED_REVERSE_DIFFUSION = 0, 1
UPD_ANCESTRAL_VP, UPD_ANCESTRAL_VE, UPD_LANGEVIN, UPD_ALD = 0, 1, 2, 3
OPT_GEOM_KERNEL, OPT_PDL, OPT_SMALL_TILES, OPT_CTA_PAIRS, OPT_FP8LO_FORCE, OPT_EXPERIMENT, OPT_LEAN_EW, OPT_GRAPH, OPT_TMA_2SM = range(9)

#: every symbol ``include/zedo_b200.h`` declares (checked by tests/test_abi.py)
EXPORTS = (
    "zedo_plan_create", "zedo_plan_destroy", "zedo_plan_capacity", "zedo_score_forward", "zedo_grad_field",
    "zedo_sde_step", "zedo_oil_loop", "zedo_ipo_fit", "zedo_rotopt_forward", "zedo_rotopt_backward",
    "zedo_eval_multi", "zedo_strerror", "zedo_abi_version", "zedo_launch_count", "zedo_subvp_scalars",
    "zedo_blocked_offset", "zedo_plan_profile", "zedo_plan_profile_read", "zedo_ipo_fit_ex", "zedo_pck_counts",
    "zedo_hypothesis_std", "zedo_plan_reserve", "zedo_set_option", "zedo_get_option", "zedo_score_stats",
    "zedo_noise_update", "zedo_kmeans_fit",
)


class ZedoError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        super().__init__(f"{where} failed with code {code}: {strerror(code)}")


class NetDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_joints", C.c_int32), ("hidden", C.c_int32), ("embed", C.c_int32),
                ("n_blocks", C.c_int32), ("gn_eps", C.c_float)]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
            "zedo_release_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    p, i32, i64, f32, vp = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_void_p
    sig = {
        "zedo_plan_create": (C.c_int, [C.POINTER(vp), C.POINTER(NetDesc), i32, C.POINTER(C.c_char_p),
                                       C.POINTER(vp), C.POINTER(i64), i64, i32]),
        "zedo_plan_destroy": (C.c_int, [vp]),
        "zedo_plan_capacity": (i64, [vp]),
        "zedo_plan_reserve": (C.c_int, [vp, i32, i32, vp]),
        "zedo_set_option": (C.c_int, [i32, i32]),
        "zedo_get_option": (C.c_int, [i32, C.POINTER(i32)]),
        "zedo_score_forward": (C.c_int, [vp, p, f32, p, i64, i32, vp]),
        "zedo_score_stats": (C.c_int, [vp, p, f32, p, f32, p, i64, i32, vp]),
        "zedo_noise_update": (C.c_int, [vp, i32, p, p, f32, f32, f32, f32, p, p, p, i64, vp]),
        "zedo_grad_field": (C.c_int, [p, p, p, p, p, i32, i32, p, p, i64, i32, vp]),
        "zedo_sde_step": (C.c_int, [vp, p, f32, p, i32, i32, f32, f32, i32, p, p, i64, i32, vp]),
        "zedo_oil_loop": (C.c_int, [vp, p, p, p, p, p, C.POINTER(f32), i32, i32, f32, f32, i32, p,
                                    C.POINTER(i32), i32, i64, i32, vp]),
        "zedo_ipo_fit": (C.c_int, [p, p, p, C.POINTER(i32), i32, i32, f32, f32, f32, i32, i64, f32, p, p, p, p,
                                   i64, i32, vp]),
        "zedo_ipo_fit_ex": (C.c_int, [p, p, p, C.POINTER(i32), i32, i32, i32, i32, i32, f32, f32, f32, i32, i64, f32, p, p,
                                      p, p, i64, i32, vp]),
        "zedo_rotopt_forward": (C.c_int, [p, p, p, p, p, f32, f32, p, i64, i32, vp]),
        "zedo_rotopt_backward": (C.c_int, [p, p, p, p, p, f32, f32, p, p, p, i64, i32, vp]),
        "zedo_eval_multi": (C.c_int, [p, p, i32, i64, i32, i32, C.POINTER(i32), i32, p, p, p, p, vp]),
        "zedo_pck_counts": (C.c_int, [p, p, p, i64, i32, i32, C.POINTER(i32), i32, p, vp]),
        "zedo_hypothesis_std": (C.c_int, [p, i64, i32, i32, p, vp]),
        "zedo_kmeans_fit": (C.c_int, [p, i64, i32, i32, i32, p, p, p, vp]),
        "zedo_strerror": (C.c_char_p, [C.c_int]),
        "zedo_abi_version": (C.c_int, []),
        "zedo_launch_count": (i64, []),
        "zedo_subvp_scalars": (C.c_int, [f32, f32, f32, C.POINTER(f32), C.POINTER(f32), C.POINTER(f32)]),
        "zedo_blocked_offset": (i64, [i64, i64, i64, i32, i32]),
        "zedo_plan_profile": (C.c_int, [vp, i32, i32]),
        "zedo_plan_profile_read": (C.c_int, [vp, i32, C.POINTER(f32), C.POINTER(i32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def strerror(code: int) -> str:
    return lib.zedo_strerror(int(code)).decode()


def check(code: int, where: str) -> None:
    if code != 0:
        raise ZedoError(code, where)


def set_option(option: int, value: int) -> None:
    check(lib.zedo_set_option(int(option), int(value)), "zedo_set_option")


def get_option(option: int) -> int:
    v = C.c_int32()
    check(lib.zedo_get_option(int(option), C.byref(v)), "zedo_get_option")
    return int(v.value)


def launch_count() -> int:
    return int(lib.zedo_launch_count())


def subvp_scalars(t: float, beta_min: float = 0.1, beta_max: float = 20.0):
    b, g, s = C.c_float(), C.c_float(), C.c_float()
    check(lib.zedo_subvp_scalars(t, beta_min, beta_max, C.byref(b), C.byref(g), C.byref(s)), "zedo_subvp_scalars")
    return b.value, g.value, s.value


def blocked_offset(row: int, col: int, cols: int, tile_rows: int, hl: int) -> int:
    return int(lib.zedo_blocked_offset(row, col, cols, tile_rows, hl))


def f32_array(values: Sequence[float]):
    return (C.c_float * len(values))(*[float(v) for v in values])


def i32_array(values: Iterable[int]):
    values = [int(v) for v in values]
    return (C.c_int32 * len(values))(*values)


def plan_create(desc: NetDesc, tensors: Dict[str, "object"], max_batch: int, device: int) -> C.c_void_p:
    """tensors: name -> contiguous float32 torch tensor (CPU or CUDA) or numpy array."""
    names, ptrs, numels, keep = [], [], [], []
    for k, v in tensors.items():
        if hasattr(v, "data_ptr"):
            t = v.detach().contiguous().float()
            keep.append(t)
            ptrs.append(t.data_ptr())
            numels.append(t.numel())
        else:
            import numpy as np
            a = np.ascontiguousarray(v, dtype=np.float32)
            keep.append(a)
            ptrs.append(a.ctypes.data)
            numels.append(a.size)
        names.append(k.encode())
    n = len(names)
    handle = C.c_void_p()
    rc = lib.zedo_plan_create(C.byref(handle), C.byref(desc), n, (C.c_char_p * n)(*names), (C.c_void_p * n)(*ptrs),
                              (C.c_int64 * n)(*numels), int(max_batch), int(device))
    check(rc, "zedo_plan_create")
    return handle

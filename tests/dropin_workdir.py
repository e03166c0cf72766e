"""A working directory the reference's own driver (run/opt_main.py) can run in: synthetic H36M-format
``data/h36m/h36m_test.pkl`` (+ detections), ``clusters/h36m_cluster{S}.npy``, a checkpoint in the reference's
format and a config FILE in the reference's format (its get_config() starts from the shipped
``configs/optim/concat_pose_optimization_h36m.py`` and only changes sizes / iteration counts)."""
import os
import pickle

import numpy as np

CONFIG_TEMPLATE = '''\
from configs.optim.concat_pose_optimization_h36m import get_config as _shipped


def get_config():
    config = _shipped()
    config.ZeDO.sample = 1
    config.ZeDO.batch = {batch}
    config.ZeDO.IPO_iterations = {ipo}
    config.ZeDO.OIL_iterations = {oil}
    return config
'''


def make_workdir(path, n_poses=64, hypo=2, ipo=10, oil=100, seed=1234):
    import torch
    from zedo_release_b200 import synthetic as sy
    ds = sy.make_synthetic_dataset(n_poses, seed=seed, n_clusters=hypo)
    os.makedirs(os.path.join(path, "data", "h36m"), exist_ok=True)
    os.makedirs(os.path.join(path, "clusters"), exist_ok=True)
    os.makedirs(os.path.join(path, "ckpt"), exist_ok=True)
    with open(os.path.join(path, "data", "h36m", "h36m_test.pkl"), "wb") as f:
        pickle.dump(sy.h36m_items_from_arrays(ds), f)
    det = {"test": {"joint3d_image": np.concatenate([ds["db_2d"][:, :, :2], np.zeros((n_poses, 17, 1), np.float32)], -1),
                    "confidence": ds["db_2d"][:, :, 2:3].copy()}}
    with open(os.path.join(path, "data", "h36m", "h36m_sh_dt_ft.pkl"), "wb") as f:
        pickle.dump(det, f)
    np.save(os.path.join(path, "clusters", f"h36m_cluster{hypo}.npy"), ds["clusters"].astype(np.float32))
    W = sy.make_weights(seed=0)
    sd = {"module." + k: torch.tensor(v) for k, v in W.items()}
    sd["module.sigmas"] = torch.tensor(np.exp(np.linspace(np.log(50), np.log(0.01), 1000)))
    shadow = [torch.tensor(v) for k, v in W.items()]
    torch.save({"model_state_dict": sd, "ema": {"decay": 0.9999, "num_updates": 7, "shadow_params": shadow},
                "step": 1500}, os.path.join(path, "ckpt", "checkpoint_1500.pth"))
    cfg = os.path.join(path, "zedo_test_config.py")
    with open(cfg, "w") as f:
        f.write(CONFIG_TEMPLATE.format(batch=n_poses, ipo=ipo, oil=oil))
    return dict(config=cfg, ckpt_dir=os.path.join(path, "ckpt"), ckpt_name="checkpoint_1500.pth", ds=ds)


PW3D_CONFIG_TEMPLATE = CONFIG_TEMPLATE.replace("concat_pose_optimization_h36m", "concat_pose_optimization_pw3d")


def make_workdir_pw3d(path, n_poses=48, hypo=2, ipo=10, oil=100, seed=4321):
    """The same for ``run/inference.py`` with the shipped 3DPW config: ``data/3dpw/pw3d_test.npz`` in the file format
    lib/dataset/pw3d.py:184-199 reads, ``clusters/h36m_cluster{S}.npy`` (run/inference.py:64-65), checkpoint, config."""
    import torch
    from zedo_release_b200 import synthetic as sy
    ds = sy.make_synthetic_dataset(n_poses, seed=seed, n_clusters=hypo)
    os.makedirs(os.path.join(path, "data", "3dpw"), exist_ok=True)
    os.makedirs(os.path.join(path, "clusters"), exist_ok=True)
    os.makedirs(os.path.join(path, "ckpt"), exist_ok=True)
    np.savez(os.path.join(path, "data", "3dpw", "pw3d_test.npz"), **sy.pw3d_npz_from_arrays(ds))
    np.save(os.path.join(path, "clusters", f"h36m_cluster{hypo}.npy"), ds["clusters"].astype(np.float32))
    W = sy.make_weights(seed=0)
    sd = {"module." + k: torch.tensor(v) for k, v in W.items()}
    sd["module.sigmas"] = torch.tensor(np.exp(np.linspace(np.log(50), np.log(0.01), 1000)))
    shadow = [torch.tensor(v) for k, v in W.items()]
    torch.save({"model_state_dict": sd, "ema": {"decay": 0.9999, "num_updates": 7, "shadow_params": shadow},
                "step": 1500}, os.path.join(path, "ckpt", "checkpoint_1500.pth"))
    cfg = os.path.join(path, "zedo_test_config_pw3d.py")
    with open(cfg, "w") as f:
        f.write(PW3D_CONFIG_TEMPLATE.format(batch=n_poses, ipo=ipo, oil=oil))
    return dict(config=cfg, ckpt_dir=os.path.join(path, "ckpt"), ckpt_name="checkpoint_1500.pth", ds=ds)


HP3D_CONFIG_TEMPLATE = CONFIG_TEMPLATE.replace("concat_pose_optimization_h36m", "concat_pose_optimization_3dhp")


def make_workdir_3dhp(path, n_poses=56, hypo=2, ipo=10, oil=100, seed=777):
    """The same for the shipped MPI-INF-3DHP config of ``run/opt_main.py``: ``data/3dhp/mpii3d_test.pkl`` as
    lib/dataset/mpii3dHP.py:255-300 reads it (a list of items: joint_3d_camera in mm, joint_2d [17,3], w, h, camera_param,
    imageid, valid_i, action 1..7), ``clusters/3dhp_cluster{S}.npy`` (run/opt_main.py:60-61), checkpoint, config."""
    import torch
    from zedo_release_b200 import synthetic as sy
    ds = sy.make_synthetic_dataset(n_poses, seed=seed, n_clusters=hypo)
    os.makedirs(os.path.join(path, "data", "3dhp"), exist_ok=True)
    os.makedirs(os.path.join(path, "clusters"), exist_ok=True)
    os.makedirs(os.path.join(path, "ckpt"), exist_ok=True)
    cam_mm = (ds["db_3d"].astype(np.float64) + ds["root"][:, None, :].astype(np.float64)) * 1000.0
    items = []
    for n in range(n_poses):
        K = ds["camera_param"][n]
        items.append({"joint_3d_camera": cam_mm[n].astype(np.float32),
                      "joint_2d": np.concatenate([ds["db_2d"][n, :, :2], np.ones((17, 1), np.float32)], -1),
                      "w": 2048, "h": 2048,
                      "camera_param": {"fx": float(K[0, 0]), "fy": float(K[1, 1]), "cx": float(K[0, 2]), "cy": float(K[1, 2])},
                      "imageid": f"synthetic/{n:08d}.jpg", "valid_i": 1, "action": 1 + n % 7})
    with open(os.path.join(path, "data", "3dhp", "mpii3d_test.pkl"), "wb") as f:
        pickle.dump(items, f)
    np.save(os.path.join(path, "clusters", f"3dhp_cluster{hypo}.npy"), ds["clusters"].astype(np.float32))
    W = sy.make_weights(seed=0)
    sd = {"module." + k: torch.tensor(v) for k, v in W.items()}
    sd["module.sigmas"] = torch.tensor(np.exp(np.linspace(np.log(50), np.log(0.01), 1000)))
    shadow = [torch.tensor(v) for k, v in W.items()]
    torch.save({"model_state_dict": sd, "ema": {"decay": 0.9999, "num_updates": 7, "shadow_params": shadow},
                "step": 1500}, os.path.join(path, "ckpt", "checkpoint_1500.pth"))
    cfg = os.path.join(path, "zedo_test_config_3dhp.py")
    with open(cfg, "w") as f:
        f.write(HP3D_CONFIG_TEMPLATE.format(batch=n_poses, ipo=ipo, oil=oil))
    return dict(config=cfg, ckpt_dir=os.path.join(path, "ckpt"), ckpt_name="checkpoint_1500.pth", ds=ds)


def parse_3dhp(stdout):
    """What MPII3DHP.eval_multi prints (lib/dataset/mpii3dHP.py:480-511): 'PCK :', 'AUC :', 'std: x.., y.., z..' and the
    '3DHP' table rows 'p1' / 'p2' (7 actions + average) -> {'p1': {...}, 'p2': {...}} in order of appearance."""
    runs, cur = [], {}
    for line in stdout.splitlines():
        t = line.strip()
        if t.startswith("PCK :"):
            cur = {"pck": float(t.split(":")[1])}
        elif t.startswith("AUC :"):
            cur["auc"] = float(t.split(":")[1])
        elif t.startswith("std: x"):
            xs = t[len("std: x"):].replace("y", " ").replace("z", " ").replace(",", " ").split()
            cur["std"] = [float(v) for v in xs]
        else:
            cells = [c.strip() for c in t.strip("|").split("|")]
            if cells and cells[0] in ("p1", "p2") and len(cells) == 9:
                cur["row"] = [float(c) for c in cells[1:]]
                cur["proto"] = cells[0]
                runs.append(cur)
                cur = {}
    return {r["proto"]: r for r in runs}


def parse_means(stdout):
    """The 'mean MPJPE : x' / 'mean PA-MPJPE : x' lines PW3D.eval_multi prints (lib/dataset/pw3d.py:338-341)."""
    out = {}
    for line in stdout.splitlines():
        line = line.strip()
        if line.startswith("mean MPJPE :"):
            out["p1"] = float(line.split(":")[1])
        elif line.startswith("mean PA-MPJPE :"):
            out["p2"] = float(line.split(":")[1])
    return out


def parse_table(stdout):
    """The 'p1' / 'p2' rows run/opt_main.py prints through eval_multi(print_verbose=True): -> {'p1': [...], 'p2': [...]}."""
    out = {}
    for line in stdout.splitlines():
        cells = [c.strip() for c in line.strip().strip("|").split("|")]
        if cells and cells[0] in ("p1", "p2") and len(cells) == 17:
            out[cells[0]] = [float(c) for c in cells[1:]]
    return out

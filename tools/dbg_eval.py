import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import zedo_oracle as zo
import zedo_release_b200 as zr
g = dict(np.load(os.path.join(ROOT, "tests/golden/eval.npz")))
pred = torch.tensor(g["preds"], device="cuda"); gt = torch.tensor(g["gts"], dtype=torch.float64, device="cuda")
for p2 in (0, 1):
    e, idx, ea = zr.eval_multi(pred, gt, protocol2=bool(p2), return_all=True)
    _, res, idx_o = zo.eval_multi(g["preds"], g["gts"], protocol2=bool(p2))
    ea = ea.cpu().numpy()
    full = np.zeros_like(ea)
    for n in range(ea.shape[0]):
        for s in range(ea.shape[1]):
            p = g["preds"][n, s]
            if p2: p = zo.procrustes_align(p, g["gts"][n])
            full[n, s] = zo.mpjpe(p, g["gts"][n])
    d = np.abs(ea - full)
    print("p2", p2, "max abs diff", d.max(), "at", np.unravel_index(d.argmax(), d.shape), "median", np.median(d), "idx equal", np.array_equal(idx.cpu().numpy(), idx_o))
    print(" worst rows", np.argsort(d.max(1))[-5:], np.sort(d.max(1))[-5:])

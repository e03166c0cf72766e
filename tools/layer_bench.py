"""GPU experiment: per-kernel time of one network forward at a given batch (CUDA events around every launch,
zedo_plan_profile), per GEMM mode and experiment switch, plus the forward error against the CUDA-core float32 mode.

    python tools/layer_bench.py [B] [reps] [modes,comma] [experiment bits,comma]
    ZEDO_B200_LIB=zedo_release_b200/libzedo_b200_exp.so python tools/layer_bench.py 262144 20 fp8lo 0,1,4,8
"""
import json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import zedo_release_b200 as zr
from zedo_release_b200 import synthetic as sy, _native as nat

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
modes = (sys.argv[3] if len(sys.argv) > 3 else "split3,fp8lo").split(",")
bits = [int(b) for b in (sys.argv[4] if len(sys.argv) > 4 else "0").split(",")]


class Clocks:
    """SM clock while the timed launches run (NVML, sampled every 5 ms on a thread); None without pynvml."""
    def __init__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        except Exception:
            self.nv = None
    def __enter__(self):
        self.samples, self.stop = [], False
        if self.nv:
            def run():
                while not self.stop:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    time.sleep(0.005)
            self.t = threading.Thread(target=run, daemon=True)
            self.t.start()
        return self
    def __exit__(self, *a):
        self.stop = True
        if self.nv:
            self.t.join()
    def median(self):
        return float(np.median(self.samples)) if self.samples else None


W = sy.make_weights(seed=0)
plan = zr.ScorePlan(W, n_joints=17, max_batch=B, device=0)
x = torch.tensor(np.random.default_rng(0).normal(0, 0.4, (B, 17, 3)).astype(np.float32), device="cuda")
ref = plan.forward(x[:8192].contiguous(), 49.95, mode="fp32")
out = {"B": B, "reps": reps, "lib": os.path.basename(nat.LIB_PATH)}
for mode in modes:
    for bit in bits:
        if bit:
            nat.set_option(nat.OPT_EXPERIMENT, bit)
        for _ in range(3):
            y = plan.forward(x, 49.95, mode=mode)
        torch.cuda.synchronize()
        plan.profile(True, 1)
        with Clocks() as clk:
            for _ in range(reps):
                y = plan.forward(x, 49.95, mode=mode)
            torch.cuda.synchronize()
        prof = plan.profile_read()
        plan.profile(False)
        err = float((y[:8192] - ref).abs().max() / ref.abs().max())
        out[f"{mode}/exp{bit}"] = {k: round(v[0], 4) for k, v in prof.items() if v[1]} | {"rel_err_vs_fp32": err, "sm_mhz": clk.median()}
        if bit:
            nat.set_option(nat.OPT_EXPERIMENT, 0)
print(json.dumps(out))

"""GPU experiment (experiments build only): clock64 trace of the operand pipeline's hand-offs in CTA pair 0 of the
hidden-layer kernel (mlp_tc2.cu, ZEDO_TRACE), 512 consecutive stage iterations in steady state.

    ZEDO_B200_LIB=zedo_release_b200/libzedo_b200_exp.so python tools/stage_trace.py [B] [extra experiment bits ...]

Prints, per variant, medians (and 10 / 90 % quantiles) in SM clocks of: the stage period, copy issue -> stage seen full,
leader's extra wait for the peer, MMA issue time, commit issued -> stage seen free again, and the producer's idle time.
"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import zedo_release_b200 as zr
from zedo_release_b200 import synthetic as sy, _native as nat

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
extra = [int(b) for b in sys.argv[2:]] or [0]
S, NEV, LEN = 3, 6, 512
W = sy.make_weights(seed=0)
plan = zr.ScorePlan(W, n_joints=17, max_batch=B, device=0)
x = torch.tensor(np.random.default_rng(0).normal(0, 0.4, (B, 17, 3)).astype(np.float32), device="cuda")
fn = nat.lib.zedo_debug_stage_trace
fn.argtypes, fn.restype = [ctypes.c_void_p, ctypes.c_int], ctypes.c_int


def q(a):
    a = np.asarray(a, dtype=np.float64)
    return [float(np.percentile(a, p)) for p in (10, 50, 90)]


out = {"B": B, "lib": os.path.basename(nat.LIB_PATH), "clocks": "10 / 50 / 90 % quantiles, SM clocks; last hidden layer of a forward"}
for bits in extra:
    nat.set_option(nat.OPT_EXPERIMENT, 16 | bits)
    for _ in range(3):
        plan.forward(x, 49.95, mode="fp8lo")
    torch.cuda.synchronize()
    buf = np.zeros((2, NEV, LEN), dtype=np.uint64)
    rc = fn(buf.ctypes.data, buf.size)
    assert rc == 0, rc
    t = buf.astype(np.int64)
    res = {}
    for cta, name in ((0, "leader"), (1, "peer")):
        e = t[cta]
        r = {"period": q(np.diff(e[5])),
             "copy_issue_to_full_seen": q(e[3] - e[2]),
             "producer_idle_waiting_for_free_stage": q(e[1] - e[0]),
             "producer_issue": q(e[2] - e[1]),
             "commit_or_relay_to_stage_free_again": q(e[1][S:] - e[5][:-S])}
        if cta == 0:
            r["leader_wait_for_peer_after_own_full"] = q(e[4] - e[3])
            r["mma_issue"] = q(e[5] - e[4])
        else:
            r["relay_arrive"] = q(e[5] - e[3])
        res[name] = r
    out[f"experiment_bits_{bits}"] = res
nat.set_option(nat.OPT_EXPERIMENT, 0)
print(json.dumps(out))

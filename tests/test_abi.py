"""The C-ABI library loads and exports every symbol include/zedo_b200.h declares; host-side
helpers agree with the oracle.  No compute call is made (no GPU needed)."""
import ctypes
import os
import re

import numpy as np

import zedo_oracle as zo
from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "zedo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zedo_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    syms = _declared_symbols()
    assert len(syms) >= 16
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/zedo_b200.h but not exported"
    assert set(syms) == set(built_lib.EXPORTS)
    assert built_lib.lib.zedo_abi_version() == 2


def test_strerror_and_argument_errors(built_lib):
    assert built_lib.strerror(0) == "ok"
    assert "invalid" in built_lib.strerror(-1)
    assert "shape" in built_lib.strerror(-2)
    # NULL arguments are rejected before any CUDA call
    assert built_lib.lib.zedo_grad_field(None, None, None, None, None, 0, 0, None, None, 1, 17, None) == -1
    assert built_lib.lib.zedo_eval_multi(None, None, 0, 1, 1, 17, None, 0, None, None, None, None, None) == -1
    assert built_lib.lib.zedo_grad_field(None, None, None, None, None, 0, 0, None, None, 0, 17, None) == 0  # empty batch
    assert built_lib.lib.zedo_score_forward(None, None, 0.0, None, 1, 0, None) == -1


def test_subvp_scalars_match_oracle(built_lib):
    for t in zo.oil_time_grid():
        b, g, s = built_lib.subvp_scalars(float(t))
        bo, go = zo.subvp_sde_scalars(t)
        so = zo.subvp_marginal_std(t)
        assert b == float(bo) and g == float(go) and s == float(so)  # same float32 op order, bit-exact


def test_blocked_layout_is_an_interleaved_bijection(built_lib):
    """Every (row, col, hl) of a [256, 128] operand maps to a distinct 2-byte slot; inside a tile
    image the eight 16-byte column chunks are the outer index and the 128 rows the inner one (the
    K-major SWIZZLE_NONE core-matrix layout of tcgen05: SBO = 128 B, LBO = tile_rows * 16 B)."""
    rows, cols, tile = 256, 128, 128
    seen = set()
    for r in range(rows):
        for c in range(cols):
            for hl in (0, 1):
                seen.add(built_lib.blocked_offset(r, c, cols, tile, hl))
    assert len(seen) == rows * cols * 2 and min(seen) == 0 and max(seen) == rows * cols * 2 * 2 - 2
    for r in (0, 1, 5, 7, 8, 13, 127):
        for chunk in range(8):
            assert built_lib.blocked_offset(r, chunk * 8, cols, tile, 0) == chunk * tile * 16 + r * 16
            assert built_lib.blocked_offset(r, chunk * 8 + 3, cols, tile, 0) == chunk * tile * 16 + r * 16 + 6
    # 8 consecutive rows of one chunk form a contiguous 128-byte core matrix
    assert [built_lib.blocked_offset(r, 16, cols, tile, 0) for r in range(8, 16)] == \
        [2 * tile * 16 + r * 16 for r in range(8, 16)]
    # tile (rt, kb) images are contiguous: hi image then lo image, 16 KiB each
    assert built_lib.blocked_offset(0, 0, cols, tile, 1) - built_lib.blocked_offset(0, 0, cols, tile, 0) == 16384
    assert built_lib.blocked_offset(0, 64, cols, tile, 0) == 32768
    assert built_lib.blocked_offset(128, 0, cols, tile, 0) == 65536


def test_mode_constants_match_the_header(built_lib):
    """The Python binding's GEMM-mode / predictor / network-kind constants are the header's."""
    text = open(os.path.join(ROOT, "include", "zedo_b200.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(ZEDO_[A-Z0-9_]+)\s+(-?\d+)\b", text)}
    from zedo_release_b200 import engine
    for name, value in engine.GEMM_MODES.items():
        assert defs[f"ZEDO_GEMM_{name.upper()}"] == value
    assert sorted(engine.GEMM_MODES.values()) == sorted(v for k, v in defs.items() if k.startswith("ZEDO_GEMM_"))
    assert defs["ZEDO_NET_SCORE_FC_ADV"] == built_lib.NET_SCORE_FC_ADV and defs["ZEDO_NET_CONTROL"] == built_lib.NET_CONTROL
    opts = {k[len("ZEDO_"):]: v for k, v in defs.items() if k.startswith("ZEDO_OPT_") and k != "ZEDO_OPT_COUNT"}
    assert sorted(opts.values()) == list(range(defs["ZEDO_OPT_COUNT"]))
    for name, value in opts.items():
        assert getattr(built_lib, name) == value, name


def test_options_and_argument_validation(built_lib):
    """zedo_set_option / zedo_get_option round trip; entry points validate sizes instead of throwing across the ABI."""
    nat = built_lib
    assert nat.get_option(nat.OPT_GEOM_KERNEL) == 0 and nat.get_option(nat.OPT_PDL) == 1
    nat.set_option(nat.OPT_GEOM_KERNEL, 2)
    assert nat.get_option(nat.OPT_GEOM_KERNEL) == 2
    nat.set_option(nat.OPT_GEOM_KERNEL, 0)
    assert nat.get_option(nat.OPT_LEAN_EW) == 16 and nat.get_option(nat.OPT_GRAPH) == 0  # defaults of the r02b options
    nat.set_option(nat.OPT_GRAPH, 1)
    assert nat.get_option(nat.OPT_GRAPH) == 1
    nat.set_option(nat.OPT_GRAPH, 0)
    assert nat.get_option(nat.OPT_TMA_2SM) == 1  # r02c: tensor-map stage copies of the CTA-pair kernel are the default
    assert nat.lib.zedo_set_option(99, 1) == -1
    assert nat.lib.zedo_set_option(nat.OPT_EXPERIMENT, 1) == -5  # timing experiments are not in the shipped build
    # negative / absurd element counts are argument errors, not std::length_error through ctypes
    desc = nat.NetDesc(nat.NET_SCORE_FC_ADV, 17, 1024, 512, 2, 1e-5)
    h = ctypes.c_void_p()
    names = (ctypes.c_char_p * 1)(b"pre_dense.weight")
    buf = (ctypes.c_float * 4)()
    ptrs = (ctypes.c_void_p * 1)(ctypes.addressof(buf))
    for bad in (-1, 1 << 40):
        rc = nat.lib.zedo_plan_create(ctypes.byref(h), ctypes.byref(desc), 1, names, ptrs, (ctypes.c_int64 * 1)(bad), 8, 0)
        assert rc == -1 and not h.value
    assert nat.lib.zedo_plan_create(ctypes.byref(h), ctypes.byref(desc), -3, names, ptrs, (ctypes.c_int64 * 1)(4), 8, 0) == -1
    assert nat.lib.zedo_plan_reserve(None, 1000, 0, None) == -1

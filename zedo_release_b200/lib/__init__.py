"""Mirror of the reference's ``lib`` package for the functions on the hot path.

Put ``<repo>/zedo_release_b200`` (and the repo root) on ``sys.path`` ahead of the reference's own
tree, or call ``zedo_release_b200.lib.install()``, and ``from lib.algorithms.advanced import
sde_lib, sampling`` in run/opt_main.py / run/inference.py resolves to these modules: same names,
argument meaning and error behaviour, arithmetic in the sm_100a kernels of ``libzedo_b200.so``.
"""
import importlib
import sys

_SUBMODULES = ("algorithms", "algorithms.advanced", "algorithms.advanced.sde_lib", "algorithms.advanced.utils",
               "algorithms.advanced.model", "algorithms.advanced.control_model", "algorithms.advanced.sampling",
               "algorithms.advanced.simple_zeroshot_opt", "algorithms.ema", "utils", "utils.transforms",
               "dataset", "dataset.synthetic")


def install(alias: str = "lib") -> None:
    """Register this package as top-level ``lib`` so the reference drivers import it unchanged."""
    sys.modules[alias] = sys.modules[__name__]
    for sub in _SUBMODULES:
        sys.modules[f"{alias}.{sub}"] = importlib.import_module(f"{__name__}.{sub}")

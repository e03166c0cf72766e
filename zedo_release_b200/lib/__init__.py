"""Mirror of the reference's ``lib`` package for the functions on the hot path.

``from lib.algorithms.advanced import sde_lib, sampling`` etc. in run/opt_main.py / run/inference.py
resolve to these modules: same names, argument meaning and error behaviour, arithmetic in the sm_100a
kernels of ``libzedo_b200.so``.  Everything the mirror does NOT carry -- the dataset loaders
(``lib.dataset.h36m`` ...), ``lib.utils.generic``, ``lib.algorithms.advanced.losses``, and single names
such as ``lib.utils.transforms.image_to_camera_frame`` -- falls through to the reference's own files
when the reference checkout is known (``install(reference_root=...)``, or the environment variable
``ZEDO_REFERENCE_ROOT``): the mirror packages get the reference's directories appended to their
``__path__`` and every mirror module gets a PEP 562 ``__getattr__`` that looks a missing name up in the
same-named reference file.

Two ways to switch a reference checkout over:

    python -m zedo_release_b200.dropin /path/to/ZeDO-Release run/opt_main.py --config ... --hypo 1
    PYTHONPATH=<repo>:<repo>/zedo_release_b200 ZEDO_REFERENCE_ROOT=/path/to/ZeDO-Release python run/opt_main.py ...
"""
import importlib
import importlib.util
import os
import sys

_SUBMODULES = ("algorithms", "algorithms.advanced", "algorithms.advanced.sde_lib", "algorithms.advanced.utils",
               "algorithms.advanced.model", "algorithms.advanced.control_model", "algorithms.advanced.sampling",
               "algorithms.advanced.simple_zeroshot_opt", "algorithms.ema", "utils", "utils.transforms",
               "dataset", "dataset.synthetic", "dataset.custom")
_PACKAGES = ("", "algorithms", "algorithms.advanced", "utils", "dataset")
_HERE = os.path.dirname(os.path.abspath(__file__))
_reference_root = None


def _ref_fallback(mod, ref_file):
    """Module-level __getattr__ (PEP 562): a name the mirror module does not define is taken from the reference's
    file of the same name, executed once under a private module name inside the mirror package (so its relative and
    ``lib.*`` imports resolve to the mirror first)."""
    state = {}

    def __getattr__(name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        if "mod" not in state:
            pkg, _, base = mod.__name__.rpartition(".")
            spec = importlib.util.spec_from_file_location(f"{pkg}._reference_{base}", ref_file)
            ref_mod = importlib.util.module_from_spec(spec)
            sys.modules[spec.name] = ref_mod
            spec.loader.exec_module(ref_mod)
            state["mod"] = ref_mod
        try:
            return getattr(state["mod"], name)
        except AttributeError:
            raise AttributeError(f"module {mod.__name__!r} (mirror and reference) has no attribute {name!r}") from None

    return __getattr__


def use_reference(reference_root, alias=None):
    """Let everything outside the hot path resolve to the reference checkout at ``reference_root``."""
    global _reference_root
    ref_lib = os.path.join(os.path.abspath(reference_root), "lib")
    if not os.path.isdir(ref_lib):
        raise FileNotFoundError(f"{reference_root!r} is not a ZeDO-Release checkout (no lib/ directory)")
    _reference_root = os.path.abspath(reference_root)
    if _reference_root not in sys.path:
        sys.path.append(_reference_root)  # `configs`, `run`
    names = [n for n in {__name__, alias} if n]
    for top in names:
        for pkg in _PACKAGES:
            m = importlib.import_module(f"{top}.{pkg}" if pkg else top)
            d = os.path.join(ref_lib, *pkg.split(".")) if pkg else ref_lib
            if os.path.isdir(d) and d not in m.__path__:
                m.__path__.append(d)
        for sub in _SUBMODULES:
            if sub in _PACKAGES:
                continue
            ref_file = os.path.join(ref_lib, *sub.split(".")) + ".py"
            if os.path.isfile(ref_file):
                m = importlib.import_module(f"{top}.{sub}")
                if "__getattr__" not in vars(m):
                    m.__getattr__ = _ref_fallback(m, ref_file)


def install(alias: str = "lib", reference_root=None) -> None:
    """Register this package as top-level ``lib`` so the reference drivers import it unchanged.  With
    ``reference_root`` (default: $ZEDO_REFERENCE_ROOT) the modules the mirror does not carry resolve to the
    reference's files (``lib.dataset.h36m``, ``lib.utils.generic``, ...)."""
    sys.modules[alias] = sys.modules[__name__]
    for sub in _SUBMODULES:
        sys.modules[f"{alias}.{sub}"] = importlib.import_module(f"{__name__}.{sub}")
    reference_root = reference_root or os.environ.get("ZEDO_REFERENCE_ROOT")
    if reference_root:
        use_reference(reference_root)


# imported as the top-level `lib` through PYTHONPATH: pick the reference checkout up from the environment
if __name__ == "lib" and os.environ.get("ZEDO_REFERENCE_ROOT"):
    use_reference(os.environ["ZEDO_REFERENCE_ROOT"])

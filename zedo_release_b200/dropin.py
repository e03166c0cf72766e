"""Run one of the reference's own drivers, unmodified, against the B200 mirror.

    python -m zedo_release_b200.dropin /path/to/ZeDO-Release run/opt_main.py --config configs/optim/... \
        --ckpt_dir ... --ckpt_name ... --hypo 1 [--gt]

registers ``zedo_release_b200.lib`` as the top-level ``lib`` package (hot-path modules -> sm_100a kernels,
everything else -> the reference checkout, see ``zedo_release_b200/lib/__init__.py``), puts the checkout on
``sys.path`` (for ``configs.*``) and executes the driver file as ``__main__`` from the current working
directory -- which, as with the reference itself, must hold ``data/`` and ``clusters/``
(run/opt_main.py:58-65,82-113).
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2:
        raise SystemExit(__doc__)
    root, driver = os.path.abspath(argv[0]), argv[1]
    path = driver if os.path.isabs(driver) else os.path.join(root, driver)
    if not os.path.isfile(path):
        raise SystemExit(f"driver {path!r} not found")
    from . import lib as mirror
    mirror.install(reference_root=root)
    sys.argv = [path] + argv[2:]
    runpy.run_path(path, run_name="__main__")


if __name__ == "__main__":
    main()

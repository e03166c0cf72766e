"""Stand-in for h5py: lib/dataset/skiPose.py and custom.py import it at module level; only File() would be used."""


def File(*args, **kwargs):
    raise ImportError("h5py is not installed in this image (oracle/shims/h5py.py is an import stand-in)")

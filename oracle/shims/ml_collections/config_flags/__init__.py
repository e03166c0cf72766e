"""``from ml_collections.config_flags import config_flags`` (run/opt_main.py:27): the flag
``--config=<file.py>`` loads the file and holds the result of its ``get_config()``."""
from . import config_flags  # noqa: F401
from .config_flags import DEFINE_config_file  # noqa: F401

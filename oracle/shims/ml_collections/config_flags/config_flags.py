import importlib.util
import os
import sys

from absl import flags


class _ConfigFileParser(flags.ArgumentParser):
    syntactic_help = "path to a Python file defining get_config()"

    def parse(self, argument):
        if not isinstance(argument, str):
            return argument
        path, _, arg = argument.partition(":")
        spec = importlib.util.spec_from_file_location("_zedo_config_" + os.path.basename(path).replace(".", "_"), path)
        if spec is None:
            raise ValueError(f"cannot load config file {path!r}")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        return mod.get_config(arg) if arg else mod.get_config()

    def flag_type(self):
        return "config file"


def DEFINE_config_file(name, default=None, help_string="path to config file.", flag_values=flags.FLAGS,
                       lock_config=True, **kwargs):
    return flags.DEFINE(_ConfigFileParser(), name, default, help_string, flag_values, **kwargs)

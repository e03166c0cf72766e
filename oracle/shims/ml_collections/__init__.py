"""Minimal stand-in for ``ml_collections``: an attribute dictionary with the ConfigDict surface the
reference's config files use (``configs/default_pose_gen_configs.py``, ``configs/optim/*.py``)."""


class ConfigDict(dict):
    def __init__(self, initial=None, **kwargs):
        super().__init__()
        for k, v in dict(initial or {}, **kwargs).items():
            self[k] = v

    def __setitem__(self, key, value):
        if isinstance(value, dict) and not isinstance(value, ConfigDict):
            value = ConfigDict(value)
        super().__setitem__(key, value)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, ConfigDict) else v) for k, v in self.items()}

    def lock(self):
        return self

    def unlock(self):
        return self

    def get_ref(self, key):
        return self[key]


FrozenConfigDict = ConfigDict

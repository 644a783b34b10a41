"""Minimal stand-in for mmcv.Config (mmcv is not a dependency): executes a python config
file and exposes its top-level names as a nested attribute dict (tools/test.py:134)."""
import os
import types


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError("'ConfigDict' object has no attribute '%s'" % k)

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(o):
    if isinstance(o, dict):
        return ConfigDict({k: _wrap(v) for k, v in o.items()})
    if isinstance(o, (list, tuple)):
        return type(o)(_wrap(v) for v in o)
    return o


class Config(object):
    def __init__(self, cfg_dict=None, filename=None):
        object.__setattr__(self, "_cfg_dict", _wrap(cfg_dict or {}))
        object.__setattr__(self, "filename", filename)

    @staticmethod
    def fromfile(filename):
        filename = os.path.abspath(os.path.expanduser(filename))
        if not os.path.isfile(filename):
            raise IOError('file "{}" does not exist'.format(filename))
        scope = {"__file__": filename}
        with open(filename) as f:
            exec(compile(f.read(), filename, "exec"), scope)
        d = {k: v for k, v in scope.items()
             if not k.startswith("__") and not isinstance(v, (types.ModuleType, types.FunctionType))}
        return Config(d, filename)

    def __getattr__(self, k):
        return getattr(self._cfg_dict, k)

    def __setattr__(self, k, v):
        self._cfg_dict[k] = _wrap(v)

    def __getitem__(self, k):
        return self._cfg_dict[k]

    def __contains__(self, k):
        return k in self._cfg_dict

    def get(self, k, default=None):
        return self._cfg_dict.get(k, default)

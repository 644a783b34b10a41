"""Name -> class registries with the reference's names (mmdet/models/registry.py:4-45)."""
import torch.nn as nn


class Registry(object):
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def register_module(self, cls=None, name=None):
        """Usable as ``@REG.register_module`` (reference style) or with an alias name."""
        def _do(c):
            if not issubclass(c, nn.Module):
                raise TypeError('module must be a child of nn.Module, but got {}'.format(c))
            key = name or c.__name__
            if key in self._module_dict:
                raise KeyError('{} is already registered in {}'.format(key, self.name))
            self._module_dict[key] = c
            return c
        return _do(cls) if cls is not None else _do


BACKBONES = Registry('backbone')
NECKS = Registry('neck')
ROI_EXTRACTORS = Registry('roi_extractor')
SHARED_HEADS = Registry('shared_head')
HEADS = Registry('head')
LOSSES = Registry('loss')
DETECTORS = Registry('detector')

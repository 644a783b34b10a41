"""ConvModule parameter container (mmdet/models/utils/conv_module.py:44-163, norm.py:12-55).

Holds nn.Conv2d / BatchNorm2d so that state_dict keys equal the reference's
(``<name>.conv.weight`` ...).  The arithmetic is executed by the conv engine
(engine.Engine.conv) with norm / bias / ReLU folded into the GEMM epilogue.
"""
import warnings

import torch.nn as nn

from .weight_init import constant_init, kaiming_init

# GN is a parameter container only (IoUawareFCOSHead): the conv engine does not plan GroupNorm layers
_NORMS = {'BN': ('bn', nn.BatchNorm2d), 'SyncBN': ('bn', nn.BatchNorm2d), 'GN': ('gn', nn.GroupNorm)}


def build_norm_layer(cfg, num_features, postfix=''):
    assert isinstance(cfg, dict) and 'type' in cfg
    cfg_ = dict(cfg)
    layer_type = cfg_.pop('type')
    if layer_type not in _NORMS:
        raise KeyError('Unrecognized norm type {}'.format(layer_type))
    abbr, cls = _NORMS[layer_type]
    requires_grad = cfg_.pop('requires_grad', True)
    cfg_.setdefault('eps', 1e-5)
    if layer_type == 'GN':                      # norm.py:48-50
        assert 'num_groups' in cfg_
        layer = cls(num_channels=num_features, **cfg_)
    else:
        layer = cls(num_features, **cfg_)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return abbr + str(postfix), layer


def build_conv_layer(cfg, *args, **kwargs):
    if cfg is not None and cfg.get('type', 'Conv') != 'Conv':
        raise NotImplementedError("conv type %r is outside the accelerated path" % cfg.get('type'))
    return nn.Conv2d(*args, **kwargs)


class ConvModule(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias='auto', conv_cfg=None, norm_cfg=None, activation='relu', inplace=True,
                 activate_last=True):
        super(ConvModule, self).__init__()
        assert activate_last, "activate_last=False is not used by the accelerated path"
        self.conv_cfg, self.norm_cfg, self.activation = conv_cfg, norm_cfg, activation
        self.with_norm = norm_cfg is not None
        self.with_activatation = activation is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.with_bias = bias
        if self.with_norm and self.with_bias:
            warnings.warn('ConvModule has norm and bias at the same time')
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, stride=stride,
                                     padding=padding, dilation=dilation, groups=groups, bias=bias)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = self.conv.kernel_size, self.conv.stride, self.conv.padding
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activatation and activation != 'relu':
            raise ValueError('{} is currently not supported.'.format(activation))
        self.init_weights()

    def init_weights(self):
        kaiming_init(self.conv, nonlinearity='relu' if self.activation is None else self.activation)
        if self.with_norm:
            constant_init(getattr(self, self.norm_name), 1, bias=0)

    def forward(self, x, activate=True, norm=True):
        raise RuntimeError("ConvModule is a parameter container here; it is executed by the owning "
                           "module's conv-engine plan")

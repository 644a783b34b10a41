"""ResNet / ResNeXt backbones with the reference's constructor arguments and state_dict keys
(mmdet/models/backbones/resnet.py:333-527, resnext.py:157-226).

The modules own the fp32 parameters; ``forward`` runs the conv-engine plan
(engine.Engine.add_backbone): every conv+BN(+ReLU)(+residual) is one tcgen05 launch.
"""
import math

import torch
import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm

from .. import engine as E
from .conv_module import build_conv_layer, build_norm_layer
from .engine_cache import PlanCache, cuda_state_dict, param_stamp, require_cuda
from .registry import BACKBONES
from .weight_init import constant_init, kaiming_init


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, style='pytorch', groups=1,
                 base_width=4, conv_cfg=None, norm_cfg=dict(type='BN')):
        super(Bottleneck, self).__init__()
        assert style in ['pytorch', 'caffe']
        width = planes if groups == 1 else math.floor(planes * (base_width / 64)) * groups
        self.inplanes, self.planes, self.stride, self.groups, self.width = inplanes, planes, stride, groups, width
        self.style = style
        # resnet.py:129-134: 'pytorch' strides the 3x3 conv2, 'caffe' the 1x1 conv1
        s1, s2 = (1, stride) if style == 'pytorch' else (stride, 1)
        self.conv1 = build_conv_layer(conv_cfg, inplanes, width, kernel_size=1, stride=s1, bias=False)
        self.add_module('bn1', build_norm_layer(norm_cfg, width, postfix=1)[1])
        self.conv2 = build_conv_layer(conv_cfg, width, width, kernel_size=3, stride=s2, padding=1,
                                      groups=groups, bias=False)
        self.add_module('bn2', build_norm_layer(norm_cfg, width, postfix=2)[1])
        self.conv3 = build_conv_layer(conv_cfg, width, planes * self.expansion, kernel_size=1, bias=False)
        self.add_module('bn3', build_norm_layer(norm_cfg, planes * self.expansion, postfix=3)[1])
        self.downsample = downsample

    @property
    def norm3(self):
        return self.bn3


def make_res_layer(inplanes, planes, blocks, stride=1, style='pytorch', groups=1, base_width=4,
                   conv_cfg=None, norm_cfg=dict(type='BN')):
    downsample = None
    if stride != 1 or inplanes != planes * Bottleneck.expansion:
        downsample = nn.Sequential(
            build_conv_layer(conv_cfg, inplanes, planes * Bottleneck.expansion, kernel_size=1,
                             stride=stride, bias=False),
            build_norm_layer(norm_cfg, planes * Bottleneck.expansion)[1])
    layers = [Bottleneck(inplanes, planes, stride, downsample, style, groups, base_width, conv_cfg, norm_cfg)]
    for _ in range(1, blocks):
        layers.append(Bottleneck(planes * Bottleneck.expansion, planes, 1, None, style, groups, base_width,
                                 conv_cfg, norm_cfg))
    return nn.Sequential(*layers)


@BACKBONES.register_module
class ResNet(nn.Module):
    arch_settings = {50: (Bottleneck, (3, 4, 6, 3)), 101: (Bottleneck, (3, 4, 23, 3)),
                     152: (Bottleneck, (3, 8, 36, 3))}

    def __init__(self, depth, num_stages=4, strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1),
                 out_indices=(0, 1, 2, 3), style='pytorch', frozen_stages=-1, conv_cfg=None,
                 norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, dcn=None,
                 stage_with_dcn=(False, False, False, False), gcb=None,
                 stage_with_gcb=(False, False, False, False), gen_attention=None,
                 stage_with_gen_attention=((), (), (), ()), with_cp=False, zero_init_residual=True,
                 groups=1, base_width=4):
        super(ResNet, self).__init__()
        if depth not in self.arch_settings:
            raise KeyError('invalid depth {} for resnet'.format(depth))
        if dcn is not None or gcb is not None or gen_attention is not None:
            raise NotImplementedError("dcn / gcb / attention plugins are outside the accelerated path")
        if tuple(strides) != (1, 2, 2, 2)[:num_stages] or tuple(dilations) != (1, 1, 1, 1)[:num_stages]:
            raise NotImplementedError("only the standard stride/dilation schedule is built")
        assert 1 <= num_stages <= 4 and max(out_indices) < num_stages
        self.depth, self.num_stages, self.out_indices, self.style = depth, num_stages, out_indices, style
        self.frozen_stages, self.norm_eval, self.zero_init_residual = frozen_stages, norm_eval, zero_init_residual
        self.groups, self.base_width = groups, base_width
        self.stage_blocks = self.arch_settings[depth][1][:num_stages]
        self.conv1 = build_conv_layer(conv_cfg, 3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.add_module('bn1', build_norm_layer(norm_cfg, 64, postfix=1)[1])
        inplanes = 64
        self.res_layers = []
        for i, nb in enumerate(self.stage_blocks):
            planes = 64 * 2 ** i
            layer = make_res_layer(inplanes, planes, nb, stride=strides[i], style=style, groups=groups,
                                   base_width=base_width, conv_cfg=conv_cfg, norm_cfg=norm_cfg)
            inplanes = planes * Bottleneck.expansion
            name = 'layer{}'.format(i + 1)
            self.add_module(name, layer)
            self.res_layers.append(name)
        self.feat_dim = Bottleneck.expansion * 64 * 2 ** (len(self.stage_blocks) - 1)
        self._plans = PlanCache()
        self._freeze_stages()

    @property
    def norm1(self):
        return self.bn1

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            for m in [self.conv1, self.bn1]:
                for p in m.parameters():
                    p.requires_grad = False
        for i in range(1, self.frozen_stages + 1):
            for p in getattr(self, 'layer{}'.format(i)).parameters():
                p.requires_grad = False

    def init_weights(self, pretrained=None):
        if isinstance(pretrained, str):
            if "://" in pretrained:
                raise RuntimeError("cannot fetch %r: model-zoo URLs need network access; pass a local "
                                   "checkpoint path or pretrained=None (tools/test.py:138 does the latter)"
                                   % pretrained)
            sd = torch.load(pretrained, map_location='cpu')
            self.load_state_dict(sd.get('state_dict', sd), strict=False)
        elif pretrained is None:
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    kaiming_init(m)
                elif isinstance(m, _BatchNorm):
                    constant_init(m, 1)
            if self.zero_init_residual:
                for m in self.modules():
                    if isinstance(m, Bottleneck):
                        constant_init(m.norm3, 0)
        else:
            raise TypeError('pretrained must be a str or None')

    # ---- engine-backed forward -------------------------------------------------------------
    def plan_into(self, eng, sd, img, prefix=""):
        if self.num_stages != 4:
            raise NotImplementedError("only 4-stage backbones are planned")
        feats = eng.add_backbone(sd, img, depth=self.depth, groups=self.groups, prefix=prefix, style=self.style)
        return feats

    def forward(self, x):
        require_cuda(x, "ResNet.forward")
        x = x.float().contiguous()
        key = (tuple(x.shape), x.device, param_stamp(self))

        def build():
            eng = E.Engine(x.device)
            inp = torch.empty_like(x)
            feats = self.plan_into(eng, cuda_state_dict(self, x.device), inp)
            outs = [eng.unpack_output(f) for f in feats]
            return eng, inp, outs
        eng, inp, outs = self._plans.get(key, build)
        inp.copy_(x)
        with torch.cuda.device(x.device):
            eng.run()
        # fresh tensors, like the reference's modules: the plan's own output buffers are overwritten by the next call
        return tuple(outs[i].clone() for i in self.out_indices)

    def train(self, mode=True):
        super(ResNet, self).train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, _BatchNorm):
                    m.eval()


@BACKBONES.register_module
class ResNeXt(ResNet):
    """resnext.py:157-226: grouped 3x3 with width = floor(planes*base_width/64)*groups."""

    def __init__(self, groups=1, base_width=4, **kwargs):
        super(ResNeXt, self).__init__(groups=groups, base_width=base_width, **kwargs)

"""FPN neck with the reference's constructor and state_dict keys (mmdet/models/necks/fpn.py:10-136).
``forward`` runs engine.Engine.add_fpn: laterals with the nearest-2x top-down add fused into the
1x1 conv epilogue, 3x3 output convs, stride-2 extra levels via phase maps."""
import torch
import torch.nn as nn

from .. import engine as E
from .conv_module import ConvModule
from .engine_cache import PlanCache, cuda_state_dict, param_stamp, require_cuda
from .registry import NECKS
from .weight_init import xavier_init


@NECKS.register_module
class FPN(nn.Module):
    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1,
                 add_extra_convs=False, extra_convs_on_inputs=True, relu_before_extra_convs=False,
                 conv_cfg=None, norm_cfg=None, activation=None):
        super(FPN, self).__init__()
        assert isinstance(in_channels, list)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_ins, self.num_outs = len(in_channels), num_outs
        self.activation, self.relu_before_extra_convs = activation, relu_before_extra_convs
        if end_level == -1:
            self.backbone_end_level = self.num_ins
            assert num_outs >= self.num_ins - start_level
        else:
            self.backbone_end_level = end_level
            assert end_level <= len(in_channels) and num_outs == end_level - start_level
        self.start_level, self.end_level = start_level, end_level
        self.add_extra_convs, self.extra_convs_on_inputs = add_extra_convs, extra_convs_on_inputs
        self.lateral_convs, self.fpn_convs = nn.ModuleList(), nn.ModuleList()
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(ConvModule(in_channels[i], out_channels, 1, conv_cfg=conv_cfg,
                                                 norm_cfg=norm_cfg, activation=activation, inplace=False))
            self.fpn_convs.append(ConvModule(out_channels, out_channels, 3, padding=1, conv_cfg=conv_cfg,
                                             norm_cfg=norm_cfg, activation=activation, inplace=False))
        extra = num_outs - self.backbone_end_level + self.start_level
        if add_extra_convs and extra >= 1:
            for i in range(extra):
                cin = in_channels[self.backbone_end_level - 1] if (i == 0 and extra_convs_on_inputs) \
                    else out_channels
                self.fpn_convs.append(ConvModule(cin, out_channels, 3, stride=2, padding=1, conv_cfg=conv_cfg,
                                                 norm_cfg=norm_cfg, activation=activation, inplace=False))
        self._plans = PlanCache()

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                xavier_init(m, distribution='uniform')

    def _check_supported(self):
        if not self.add_extra_convs or self.activation is not None or self.end_level != -1:
            raise NotImplementedError("only FPNs whose extra levels are stride-2 convs (RetinaNet / FCOS variants, "
                                      "no activation) are planned")
        if any(m.with_norm for m in self.lateral_convs):
            raise NotImplementedError("FPN with norm layers is not planned")

    def plan_into(self, eng, sd, feats, prefix=""):
        self._check_supported()
        return eng.add_fpn(sd, feats, prefix=prefix, start_level=self.start_level, num_outs=self.num_outs,
                           out_channels=self.out_channels, extra_convs_on_inputs=self.extra_convs_on_inputs,
                           relu_before_extra_convs=self.relu_before_extra_convs)

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        for t in inputs:
            require_cuda(t, "FPN.forward")
        inputs = [t.float().contiguous() for t in inputs]
        dev = inputs[0].device
        key = (tuple(tuple(t.shape) for t in inputs), dev, param_stamp(self))

        def build():
            eng = E.Engine(dev)
            ins = [torch.empty_like(t) for t in inputs]
            maps = [eng.pack_input(t) for t in ins]
            F = self.plan_into(eng, cuda_state_dict(self, dev), maps)
            outs = [eng.unpack_output(F, s) for s in range(len(F.segs))]
            return eng, ins, outs
        eng, ins, outs = self._plans.get(key, build)
        for a, b in zip(ins, inputs):
            a.copy_(b)
        with torch.cuda.device(dev):
            eng.run()
        return tuple(o.clone() for o in outs)      # fresh tensors (the plan's buffers are reused by the next call)

"""Loss modules named by the configs (LOSSES registry).  The head constructor builds them
(mmdet/models/anchor_heads/anchor_head.py:71-72) so the config files load unchanged; training
is outside the accelerated path, so calling one raises."""
import torch.nn as nn

from .registry import LOSSES


class _TrainingOnlyLoss(nn.Module):
    def forward(self, *args, **kwargs):
        raise NotImplementedError("%s: training losses are outside the accelerated inference path"
                                  % self.__class__.__name__)


@LOSSES.register_module
class FocalLoss(_TrainingOnlyLoss):
    def __init__(self, use_sigmoid=False, loss_weight=1.0, gamma=2.0, alpha=0.25):
        super(FocalLoss, self).__init__()
        assert use_sigmoid is True, 'Only sigmoid focaloss supported now.'
        self.use_sigmoid, self.loss_weight, self.gamma, self.alpha = use_sigmoid, loss_weight, gamma, alpha


@LOSSES.register_module
class SmoothL1Loss(_TrainingOnlyLoss):
    def __init__(self, beta=1.0, loss_weight=1.0):
        super(SmoothL1Loss, self).__init__()
        self.beta, self.loss_weight = beta, loss_weight


@LOSSES.register_module
class CrossEntropyLoss(_TrainingOnlyLoss):
    def __init__(self, use_sigmoid=False, use_mask=False, loss_weight=1.0):
        super(CrossEntropyLoss, self).__init__()
        self.use_sigmoid, self.use_mask, self.loss_weight = use_sigmoid, use_mask, loss_weight

"""``mmdet.ops`` surface backed by libiou_b200 (mmdet/ops/__init__.py:1-19).

nms                -> iou_nms                (ops/nms/nms_wrapper.py:8-49, nms_kernel.cu:70-131)
sigmoid_focal_loss -> iou_sigmoid_focal_loss_* (ops/sigmoid_focal_loss/functions/sigmoid_focal_loss.py:8-42)
soft_nms           -> iou_soft_nms           (ops/nms/nms_wrapper.py:52-78, ops/nms/src/soft_nms_cpu.pyx:22-127)
There is no CPU implementation here: CPU tensors raise (the CPU path is the oracle's job).
"""
import types

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import lib as L
from .. import postproc as PP


def _nms_cpu_unavailable(dets, iou_thr):
    raise RuntimeError("libiou_b200 implements nms for CUDA tensors only (no CPU fallback); "
                       "pass a CUDA tensor or a device_id")


nms_cuda = types.SimpleNamespace(nms=PP.nms_cuda)
nms_cpu = types.SimpleNamespace(nms=_nms_cpu_unavailable)


def nms(dets, iou_thr, device_id=None):
    """Same contract as nms_wrapper.nms: returns (dets[inds, :], inds), inds ascending."""
    if isinstance(dets, torch.Tensor):
        is_numpy, dets_th = False, dets
    elif isinstance(dets, np.ndarray):
        is_numpy = True
        device = 'cpu' if device_id is None else 'cuda:{}'.format(device_id)
        dets_th = torch.from_numpy(dets).to(device)
    else:
        raise TypeError('dets must be either a Tensor or numpy array, but got {}'.format(type(dets)))
    if dets_th.shape[0] == 0:
        inds = dets_th.new_zeros(0, dtype=torch.long)
    elif dets_th.is_cuda:
        inds = nms_cuda.nms(dets_th, iou_thr)
    else:
        inds = nms_cpu.nms(dets_th, iou_thr)
    if is_numpy:
        inds = inds.cpu().numpy()
    return dets[inds, :], inds


def _soft_nms_device(dets_np_or_tensor, iou_thr, method=1, sigma=0.5, min_score=0.001):
    """Stands where the reference's Cython module function soft_nms_cpu.soft_nms_cpu stood (same arguments and
    return order); the rows are processed on the current CUDA device."""
    if isinstance(dets_np_or_tensor, np.ndarray):
        if not torch.cuda.is_available():
            raise RuntimeError("libiou_b200 implements soft_nms on a CUDA device only (no CPU fallback)")
        d = torch.from_numpy(np.ascontiguousarray(dets_np_or_tensor, dtype=np.float32)).cuda()
        new_dets, inds = PP.soft_nms_cuda(d, iou_thr, method, sigma, min_score)
        return new_dets.cpu().numpy(), inds.cpu().numpy()
    return PP.soft_nms_cuda(dets_np_or_tensor, iou_thr, method, sigma, min_score)


soft_nms_cpu = types.SimpleNamespace(soft_nms_cpu=_soft_nms_device)


def soft_nms(dets, iou_thr, method='linear', sigma=0.5, min_score=1e-3):
    """Same contract as nms_wrapper.soft_nms (ops/nms/nms_wrapper.py:52-78): Tensor or ndarray in,
    (new_dets with decayed scores, inds) out, in selection order.  CUDA tensors stay on their device; CPU
    tensors raise (there is no CPU implementation in this library); ndarrays go through the current device."""
    if isinstance(dets, torch.Tensor):
        is_tensor = True
        if not dets.is_cuda:
            raise RuntimeError("libiou_b200 implements soft_nms for CUDA tensors only (no CPU fallback)")
    elif isinstance(dets, np.ndarray):
        is_tensor = False
    else:
        raise TypeError('dets must be either a Tensor or numpy array, but got {}'.format(type(dets)))
    method_codes = PP.SOFT_NMS_METHODS
    if method not in method_codes:
        raise ValueError('Invalid method for SoftNMS: {}'.format(method))
    new_dets, inds = soft_nms_cpu.soft_nms_cpu(dets, iou_thr, method=method_codes[method], sigma=sigma,
                                               min_score=min_score)
    if is_tensor:
        return new_dets.to(dets.dtype), inds
    return new_dets.astype(np.float32), inds.astype(np.int64)


_FOCAL_DTYPES = {torch.float32: L.DTYPE_F32, torch.float16: L.DTYPE_F16, torch.float64: L.DTYPE_F64}


def _focal_dtype(logits):
    if not logits.is_cuda:
        raise RuntimeError("sigmoid_focal_loss_cuda: logits must be a CUDA tensor")      # sigmoid_focal_loss.cpp:21-25
    if logits.dim() != 2:
        raise RuntimeError("sigmoid_focal_loss_cuda: logits should be NxClass")          # .cu:113
    if logits.dtype not in _FOCAL_DTYPES:                 # AT_DISPATCH_FLOATING_TYPES_AND_HALF (.cu:128)
        raise RuntimeError("sigmoid_focal_loss_cuda: not implemented for %s" % logits.dtype)
    return _FOCAL_DTYPES[logits.dtype]


class _FocalLossCuda(object):
    """Stands where the reference's pybind module sigmoid_focal_loss_cuda stood: fp16 / fp32 / fp64 logits in,
    losses / gradients of the same dtype out."""

    @staticmethod
    def forward(logits, targets, num_classes, gamma, alpha):
        dt = _focal_dtype(logits)
        x = logits.detach().contiguous()
        t = targets.detach().long().contiguous()
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            L.check(L.load().iou_sigmoid_focal_loss_forward_dtype(x.data_ptr(), dt, t.data_ptr(), x.shape[0],
                                                                  x.shape[1], gamma, alpha, out.data_ptr(),
                                                                  L.stream_ptr()))
        L.launch_count += 1
        return out

    @staticmethod
    def backward(logits, targets, d_losses, num_classes, gamma, alpha):
        dt = _focal_dtype(logits)
        if logits.shape[1] != num_classes:
            raise RuntimeError("logits.size(1) should be num_classes")                   # .cu:151-152
        x = logits.detach().contiguous()
        t = targets.detach().long().contiguous()
        g = d_losses.detach().to(x.dtype).contiguous()
        if g.numel() != x.numel():
            raise RuntimeError("sigmoid_focal_loss_cuda.backward: d_losses must have one element per logit")
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            L.check(L.load().iou_sigmoid_focal_loss_backward_dtype(x.data_ptr(), dt, t.data_ptr(), g.data_ptr(),
                                                                   x.shape[0], num_classes, gamma, alpha,
                                                                   out.data_ptr(), L.stream_ptr()))
        L.launch_count += 1
        return out


sigmoid_focal_loss_cuda = _FocalLossCuda


class SigmoidFocalLossFunction(Function):
    """ops/sigmoid_focal_loss/functions/sigmoid_focal_loss.py:8-42.  One deviation, on purpose: for the 'mean' and
    'sum' reductions the reference hands its kernel the 0-dim upstream gradient, which the kernel then indexes per
    element (sigmoid_focal_loss_cuda.cu:103, an out-of-bounds read); here that scalar is broadcast -- and divided by
    numel for 'mean' -- so the gradient is the derivative of what forward returned."""

    @staticmethod
    def forward(ctx, input, target, gamma=2.0, alpha=0.25, reduction='mean'):
        ctx.save_for_backward(input, target)
        ctx.num_classes, ctx.gamma, ctx.alpha, ctx.reduction = input.shape[1], gamma, alpha, reduction
        loss = sigmoid_focal_loss_cuda.forward(input, target, input.shape[1], gamma, alpha)
        if reduction == 'none':
            return loss
        if reduction == 'mean':
            return loss.mean()
        if reduction == 'sum':
            return loss.sum()
        raise ValueError("{} is not a valid value for reduction".format(reduction))     # F._Reduction.get_enum

    @staticmethod
    @once_differentiable
    def backward(ctx, d_loss):
        input, target = ctx.saved_tensors
        if ctx.reduction == 'none':
            d = d_loss.contiguous()
        else:
            scale = 1.0 / input.numel() if ctx.reduction == 'mean' else 1.0
            d = (d_loss * scale).expand_as(input).contiguous()
        d_input = sigmoid_focal_loss_cuda.backward(input, target, d, ctx.num_classes, ctx.gamma, ctx.alpha)
        return d_input, None, None, None, None


sigmoid_focal_loss = SigmoidFocalLossFunction.apply


class SigmoidFocalLoss(nn.Module):
    def __init__(self, gamma, alpha):
        super(SigmoidFocalLoss, self).__init__()
        self.gamma, self.alpha = gamma, alpha

    def forward(self, logits, targets):
        assert logits.is_cuda
        # modules/sigmoid_focal_loss.py:13-16: the function's default reduction ('mean') and then .sum() of that
        # 0-dim result, i.e. the MEAN over the N*C elements
        loss = sigmoid_focal_loss(logits, targets, self.gamma, self.alpha)
        return loss.sum()

    def __repr__(self):
        return "{}(gamma={}, alpha={})".format(self.__class__.__name__, self.gamma, self.alpha)

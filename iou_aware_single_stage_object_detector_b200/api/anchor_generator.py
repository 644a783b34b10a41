"""AnchorGenerator with the reference's interface (mmdet/core/anchor/anchor_generator.py:4-84).

On the inference path only ``base_anchors`` is consumed: the decode kernel rebuilds
``base[a] + (x*s, y*s, x*s, y*s)`` per candidate, so no anchor grid is ever materialised.
``grid_anchors`` stays available for API parity (tiny torch code, default device follows the
base anchors instead of the reference's hard-coded 'cuda', anchor_generator.py:53).
"""
import torch


class AnchorGenerator(object):
    def __init__(self, base_size, scales, ratios, scale_major=True, ctr=None):
        self.base_size = base_size
        self.scales = torch.Tensor(scales)
        self.ratios = torch.Tensor(ratios)
        self.scale_major = scale_major
        self.ctr = ctr
        self.base_anchors = self.gen_base_anchors()

    @property
    def num_base_anchors(self):
        return self.base_anchors.size(0)

    def gen_base_anchors(self):
        size = self.base_size
        cx, cy = (0.5 * (size - 1), 0.5 * (size - 1)) if self.ctr is None else self.ctr
        hr = torch.sqrt(self.ratios)
        wr = 1 / hr
        if self.scale_major:      # rows ordered ratio-major: a = ratio_idx * n_scales + scale_idx
            ws = (size * wr[:, None] * self.scales[None, :]).reshape(-1)
            hs = (size * hr[:, None] * self.scales[None, :]).reshape(-1)
        else:
            ws = (size * self.scales[:, None] * wr[None, :]).reshape(-1)
            hs = (size * self.scales[:, None] * hr[None, :]).reshape(-1)
        half_w, half_h = 0.5 * (ws - 1), 0.5 * (hs - 1)
        return torch.stack([cx - half_w, cy - half_h, cx + half_w, cy + half_h], dim=-1).round()

    def grid_anchors(self, featmap_size, stride=16, device=None):
        base = self.base_anchors if device is None else self.base_anchors.to(device)
        feat_h, feat_w = featmap_size
        sx = torch.arange(0, feat_w, device=base.device) * stride
        sy = torch.arange(0, feat_h, device=base.device) * stride
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        shifts = torch.stack([xx, yy, xx, yy], dim=-1).reshape(-1, 1, 4).type_as(base)
        return (base[None, :, :] + shifts).reshape(-1, 4)

    def valid_flags(self, featmap_size, valid_size, device=None):
        feat_h, feat_w = featmap_size
        valid_h, valid_w = valid_size
        assert valid_h <= feat_h and valid_w <= feat_w
        dev = self.base_anchors.device if device is None else device
        vy = (torch.arange(feat_h, device=dev) < valid_h)
        vx = (torch.arange(feat_w, device=dev) < valid_w)
        valid = (vy[:, None] & vx[None, :]).reshape(-1)
        return valid[:, None].expand(valid.size(0), self.num_base_anchors).reshape(-1)

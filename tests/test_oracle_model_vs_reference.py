"""Pins oracle/model.py (the torch-CPU restatement of backbone+FPN+head) to the REAL reference's
modules on identical weights.  Needs /root/reference, so it runs in the build container only; it
executes in a subprocess because the reference's package is also called `mmdet`."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r'''
import sys, torch
sys.path.insert(0, %r)
from oracle import ref_shim, model as om
torch.manual_seed(0)
m, cfg = ref_shim.build_reference_detector()
sd = {k: v.clone() for k, v in m.state_dict().items()}
om.spread_weights_(sd, seed=1)
m.load_state_dict(sd)
img = torch.randn(1, 3, 128, 160)
with torch.no_grad():
    feats = m.extract_feat(img)
    outs = m.bbox_head(feats)
mine = om.detector_forward(sd, img)
worst = 0.0
for a_list, b_list in zip(outs, mine):
    for a, b in zip(a_list, b_list):
        assert a.shape == b.shape
        worst = max(worst, (a - b).abs().max().item())
print("WORST", worst)
assert worst == 0.0, worst
# end to end: reference get_bboxes == oracle get_bboxes on the reference's own maps
from oracle import postproc as op
meta = dict(ori_shape=(128,157,3), img_shape=(128,157,3), pad_shape=(128,160,3), scale_factor=1.0, flip=False)
with torch.no_grad():
    res = m.bbox_head.get_bboxes(*outs, [torch.zeros(0,4)], [torch.zeros(0,dtype=torch.long)], [meta], cfg.test_cfg, rescale=True)
sc = op.retina_anchor_scales(4, 3)
bases = [op.base_anchors(s, sc, [0.5, 1.0, 2.0]) for s in (8, 16, 32, 64, 128)]
d, l = op.get_bboxes_single([c[0] for c in outs[0]], [r[0] for r in outs[1]], [q[0] for q in outs[2]],
                            [8, 16, 32, 64, 128], bases, meta["img_shape"], 1.0, dict(cfg.test_cfg), rescale=True, nms_mode="cpu")
assert torch.equal(d, res[0][0]) and torch.equal(l, res[0][1]), (d.shape, res[0][0].shape)
print("DETS", d.shape[0])
'''


@pytest.mark.reference
def test_oracle_model_equals_reference_modules():
    out = subprocess.run([sys.executable, "-c", CODE % ROOT], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    assert "WORST 0.0" in out.stdout


FCOS_CODE = r'''
import sys, torch
sys.path.insert(0, %r)
from oracle import ref_shim, model as om
ref_shim.load_reference()
from mmdet.models.anchor_heads import IoUawareFCOSHead
torch.manual_seed(0)
head = IoUawareFCOSHead(num_classes=81, in_channels=256, stacked_convs=4, feat_channels=256, strides=[8, 16, 32, 64, 128])
head.init_weights()
g = torch.Generator().manual_seed(5)
with torch.no_grad():
    for k, v in head.state_dict().items():          # spread the affine GN / Scale parameters and the convs
        if ".gn.weight" in k: v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif ".gn.bias" in k: v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith(".scale"): v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif k.endswith("conv.weight"): v.copy_(torch.randn(v.shape, generator=g) * 0.03)
head.eval()
feats = [torch.randn(2, 256, h, w, generator=g) for (h, w) in [(16, 20), (8, 10), (4, 5), (2, 3), (1, 2)]]
with torch.no_grad():
    outs = head(feats)
sd = {"bbox_head." + k: v for k, v in head.state_dict().items()}
mine = om.fcos_head_forward(sd, feats)
worst = 0.0
for a_list, b_list in zip(outs, mine):
    for a, b in zip(a_list, b_list):
        assert a.shape == b.shape
        worst = max(worst, (a - b).abs().max().item())
print("WORST", worst)
assert worst == 0.0, worst
'''


@pytest.mark.reference
def test_oracle_fcos_head_equals_reference_module():
    """oracle.model.fcos_head_forward == the reference IoUawareFCOSHead.forward, bit for bit."""
    out = subprocess.run([sys.executable, "-c", FCOS_CODE % ROOT], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    assert "WORST 0.0" in out.stdout


FCOS_DET_CODE = r'''
import sys, torch
sys.path.insert(0, %r)
from oracle import ref_shim, model as om, postproc as op
ref_shim.load_reference()
cfg = ref_shim.load_config("/root/reference/configs/fcos/iou_aware_fcos_r50_caffe_fpn_gn_1x_4gpu.py")
torch.manual_seed(0)
m, cfg = ref_shim.build_reference_detector(cfg)
sd = {k: v.clone() for k, v in m.state_dict().items()}
om.spread_fcos_weights_(sd, seed=1)
m.load_state_dict(sd)
img = torch.randn(1, 3, 128, 160)
with torch.no_grad():
    outs = m.bbox_head(m.extract_feat(img))
mine = om.fcos_detector_forward(sd, img)
worst = 0.0
for a_list, b_list in zip(outs, mine):
    for a, b in zip(a_list, b_list):
        assert a.shape == b.shape
        worst = max(worst, (a - b).abs().max().item())
print("WORST", worst)
assert worst == 0.0, worst
meta = dict(ori_shape=(128,157,3), img_shape=(128,157,3), pad_shape=(128,160,3), scale_factor=1.0, flip=False)
with torch.no_grad():
    res = m.bbox_head.get_bboxes(*outs, [torch.zeros(0,4)], [torch.zeros(0,dtype=torch.long)], [meta], cfg.test_cfg, rescale=True)
d, l = op.fcos_get_bboxes_single([c[0] for c in outs[0]], [r[0] for r in outs[1]], [q[0] for q in outs[3]],
                                 [8, 16, 32, 64, 128], meta["img_shape"], 1.0, dict(cfg.test_cfg), rescale=True, nms_mode="cpu")
assert torch.equal(d, res[0][0]) and torch.equal(l, res[0][1]), (d.shape, res[0][0].shape)
print("DETS", d.shape[0])
'''


@pytest.mark.reference
def test_oracle_fcos_detector_equals_reference():
    """caffe-style ResNet + FCOS FPN variant + IoUawareFCOSHead + get_bboxes == the reference FCOS detector."""
    out = subprocess.run([sys.executable, "-c", FCOS_DET_CODE % ROOT], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    assert "WORST 0.0" in out.stdout

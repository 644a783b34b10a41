"""Whole detector, DEFAULT scheme (fp16 + e4m3, passes = 2), against goldens written by the LIVE shimmed reference
detector (tests/golden/gen_golden_detector.py -> mmdet.models.build_detector + the reference's own forward and
get_bboxes, iou_aware_retina_head.py:390-564).  BASELINE config 2 at its full size (R50, 800x1344, 2 images) and the
R101 / ResNeXt-101 backbones of configs 3-4 (resnext.py:21-56) at 256x320."""
import pytest

import parity_util as U

pytestmark = pytest.mark.gpu


def test_r50_full_size_default_scheme_vs_reference_golden():
    r = U.check_detector_golden("r50_full")
    assert r["passes"] == 2


def test_r50_full_size_cuda_graph_replay_vs_reference_golden():
    U.check_detector_golden("r50_full", use_graph=True, verbose=False)


@pytest.mark.parametrize("name", ["r101_small", "x101_32x4d_small", "x101_64x4d_small"])
def test_deep_backbones_default_scheme_vs_reference_golden(name):
    r = U.check_detector_golden(name)
    assert r["passes"] == 2


def test_r50_full_size_bf16x3_vs_reference_golden():
    r = U.check_detector_golden("r50_full", passes=3, verbose=False)
    assert r["passes"] == 3

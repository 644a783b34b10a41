# IoU-aware RetinaNet X101-32x4d-FPN -- inference-side settings of the reference config of the same
# name (configs/iou_aware_single_stage_detector/iou_aware_retinanet_x101_32x4d_fpn_1x_4gpu.py in the reference tree); the reference file
# itself also loads unchanged through iou_aware_single_stage_object_detector_b200.Config.
backbone = dict(type='ResNeXt', depth=101, groups=32, base_width=4, num_stages=4, out_indices=(0, 1, 2, 3),
                frozen_stages=1, style='pytorch')
neck = dict(type='FPN', in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=1,
            add_extra_convs=True, num_outs=5)
bbox_head = dict(
    type='IoUawareRetinaHead', num_classes=81, in_channels=256, stacked_convs=4, feat_channels=256,
    octave_base_scale=4, scales_per_octave=3, anchor_ratios=[0.5, 1.0, 2.0],
    anchor_strides=[8, 16, 32, 64, 128], target_means=[.0, .0, .0, .0], target_stds=[1.0, 1.0, 1.0, 1.0],
    loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
    loss_bbox=dict(type='SmoothL1Loss', beta=0.11, loss_weight=1.0))
model = dict(type='RetinaNet', pretrained='open-mmlab://resnext101_32x4d', backbone=backbone, neck=neck, bbox_head=bbox_head)
train_cfg = None   # training is outside this build
test_cfg = dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05, nms=dict(type='nms', iou_thr=0.5),
                max_per_img=100)
img_norm_cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
data = dict(imgs_per_gpu=2, test=dict(img_scale=(1333, 800), size_divisor=32, flip_ratio=0))
dist_params = dict(backend='nccl')

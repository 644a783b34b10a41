# IoU-aware FCOS R50-caffe-FPN-GN -- inference-side settings of the reference config of the same name
# (configs/fcos/iou_aware_fcos_r50_caffe_fpn_gn_1x_4gpu.py in the reference tree); the reference file itself also
# loads unchanged through iou_aware_single_stage_object_detector_b200.Config.
backbone = dict(type='ResNet', depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                norm_cfg=dict(type='BN', requires_grad=False), style='caffe')
neck = dict(type='FPN', in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=1, add_extra_convs=True,
            extra_convs_on_inputs=False, num_outs=5, relu_before_extra_convs=True)      # P6 from P5, ReLU before P7
bbox_head = dict(type='IoUawareFCOSHead', num_classes=81, in_channels=256, stacked_convs=4, feat_channels=256,
                 strides=[8, 16, 32, 64, 128])
model = dict(type='FCOS', pretrained='open-mmlab://resnet50_caffe', backbone=backbone, neck=neck, bbox_head=bbox_head)
train_cfg = None   # training is outside this build
test_cfg = dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05, nms=dict(type='nms', iou_thr=0.5),
                max_per_img=100)
img_norm_cfg = dict(mean=[102.9801, 115.9465, 122.7717], std=[1.0, 1.0, 1.0], to_rgb=False)
data = dict(imgs_per_gpu=4, test=dict(img_scale=(1333, 800), size_divisor=32, flip_ratio=0))
dist_params = dict(backend='nccl')

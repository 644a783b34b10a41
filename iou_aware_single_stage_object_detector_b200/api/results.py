"""Result formats downstream of the path.

bbox2result                 mmdet/core/bbox/transforms.py:148-166  (in .transforms)
batch_bbox2result           the same per-class split for a whole batch from the padded device tensors the
                            detector returns (dets [n,K,5], labels [n,K], counts [n]): ONE device->host copy
                            instead of one per image and 80 boolean masks each
xyxy2xywh / det2json / results2json
                            mmdet/core/evaluation/coco_utils.py:78-85,103-117,140-149 (detection branch)
dump_results                ``mmcv.dump(outputs, args.out)`` of tools/test.py:176-178: pickle for .pkl/.pickle,
                            json for .json (mmcv 0.2.x picks the handler from the file extension)
"""
import json
import pickle

import numpy as np
import torch


def batch_bbox2result(dets, labels, counts, num_classes):
    """list over images of [num_classes-1 arrays (n_c,5) float32], rows of a class in detection order
    (== ``bboxes[labels == i, :]``, transforms.py:165)."""
    if isinstance(dets, torch.Tensor):
        packed = torch.cat([dets.reshape(dets.shape[0], -1), labels.to(dets.dtype), counts.to(dets.dtype)[:, None]],
                           dim=1).cpu().numpy()                      # one D2H (labels < 2^24 are exact in fp32)
        n, K = dets.shape[0], dets.shape[1]
        d = packed[:, :K * 5].reshape(n, K, 5)
        lab = packed[:, K * 5:K * 6].astype(np.int64)
        cnt = packed[:, K * 6].astype(np.int64)
    else:
        d, lab, cnt = np.asarray(dets), np.asarray(labels), np.asarray(counts)
    out = []
    for i in range(d.shape[0]):
        k = int(cnt[i])
        if k == 0:
            out.append([np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes - 1)])
            continue
        di, li = d[i, :k].astype(np.float32, copy=False), lab[i, :k]
        order = np.argsort(li, kind='stable')                         # class-major, detection order inside a class
        bounds = np.searchsorted(li[order], np.arange(num_classes))
        out.append([di[order[bounds[c]:bounds[c + 1]]] for c in range(num_classes - 1)])
    return out


def xyxy2xywh(bbox):
    _bbox = bbox.tolist()
    return [_bbox[0], _bbox[1], _bbox[2] - _bbox[0] + 1, _bbox[3] - _bbox[1] + 1]


def det2json(dataset, results):
    """dataset needs ``img_ids`` and ``cat_ids`` (CocoDataset attributes, datasets/coco.py)."""
    json_results = []
    for idx in range(len(dataset)):
        img_id = dataset.img_ids[idx]
        result = results[idx]
        for label in range(len(result)):
            bboxes = result[label]
            for i in range(bboxes.shape[0]):
                json_results.append(dict(image_id=img_id, bbox=xyxy2xywh(bboxes[i]), score=float(bboxes[i][4]),
                                         category_id=dataset.cat_ids[label]))
    return json_results


def results2json(dataset, results, out_file):
    if isinstance(results[0], list):
        json_results = det2json(dataset, results)
    elif isinstance(results[0], tuple):
        raise NotImplementedError("segm results are outside the single-stage detection path")
    elif isinstance(results[0], np.ndarray):
        raise NotImplementedError("proposal results are outside the single-stage detection path")
    else:
        raise TypeError('invalid type of results')
    dump_results(json_results, out_file)
    return json_results


def dump_results(obj, out_file):
    ext = out_file.rsplit('.', 1)[-1].lower()
    if ext in ('pkl', 'pickle'):
        with open(out_file, 'wb') as f:
            pickle.dump(obj, f, protocol=2)        # mmcv PickleHandler default
    elif ext == 'json':
        with open(out_file, 'w') as f:
            json.dump(obj, f)
    else:
        raise TypeError('Unsupported format: {}'.format(ext))

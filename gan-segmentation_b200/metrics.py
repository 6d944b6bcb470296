"""Mirror of the evaluation metric of the reference (metrics.py:497-606: ``SegmentationMetric``, pixel accuracy and
mean IoU with label -1 ignored and the background class skipped), on numpy arrays instead of NDArrays."""
from __future__ import annotations

import numpy as np


def _class_ids(pred):
    """logits [N,K,H,W] -> first-max class ids [N,H,W] (np.argmax, metrics.py:574); class ids pass through."""
    pred = np.asarray(pred)
    if pred.ndim == 4:
        return np.argmax(pred, 1).astype(np.int64)
    return pred.astype(np.int64)


def batch_pix_accuracy(output, target):
    """metrics.py:570-583: (#correct, #labelled) over the pixels whose label is not -1."""
    predict = _class_ids(output) + 1
    target = np.asarray(target).astype(np.int64) + 1
    pixel_labeled = int(np.sum(target > 0))
    pixel_correct = int(np.sum((predict == target) * (target > 0)))
    assert pixel_correct <= pixel_labeled, 'Correct area should be smaller than Labeled'
    return pixel_correct, pixel_labeled


def batch_intersection_union(output, target, nclass):
    """metrics.py:586-606: per-class intersection and union areas (histograms over 1..nclass)."""
    predict = _class_ids(output) + 1
    target = np.asarray(target).astype(np.int64) + 1
    predict = predict * (target > 0).astype(predict.dtype)
    intersection = predict * (predict == target)
    area_inter, _ = np.histogram(intersection, bins=nclass, range=(1, nclass))
    area_pred, _ = np.histogram(predict, bins=nclass, range=(1, nclass))
    area_lab, _ = np.histogram(target, bins=nclass, range=(1, nclass))
    area_union = area_pred + area_lab - area_inter
    assert (area_inter <= area_union).all(), 'Intersection area should be smaller than Union area'
    return area_inter, area_union


class SegmentationMetric:
    """``SegmentationMetric(nclass, skip_bg=True)`` (metrics.py:497-567): accumulates pixAcc and mIoU."""

    def __init__(self, nclass, skip_bg=True):
        self.name = 'pixAcc & mIoU'
        self.nclass = nclass
        self._skip_bg = skip_bg
        self.reset()

    def reset(self):
        self.total_inter = 0
        self.total_union = 0
        self.total_correct = 0
        self.total_label = 0

    def update(self, labels, preds):
        """labels: [N,H,W] int (or a list of them), preds: logits [N,K,H,W] or class ids [N,H,W] (or a list)."""
        if not isinstance(preds, (list, tuple)):
            labels, preds = [labels], [preds]
        for label, pred in zip(labels, preds):
            correct, labeled = batch_pix_accuracy(pred, label)
            inter, union = batch_intersection_union(pred, label, self.nclass)
            self.total_correct += correct
            self.total_label += labeled
            self.total_inter = self.total_inter + inter
            self.total_union = self.total_union + union

    def get(self):
        pix_acc = 1.0 * self.total_correct / (np.spacing(1) + self.total_label)
        iou = 1.0 * self.total_inter / (np.spacing(1) + self.total_union)
        iou = iou[self.total_union > 0]
        if self._skip_bg:
            iou = iou[1:]
        return ['accuracy', 'mean-iou'], [pix_acc, iou.mean()]

    def get_name_value(self):
        names, values = self.get()
        return list(zip(names, values))

"""SegmentationMetric mirror (metrics.py:497-606) against hand-computed answers."""
import numpy as np

from gan_segmentation_b200.metrics import SegmentationMetric, batch_intersection_union, batch_pix_accuracy


def test_pixacc_and_miou_known_answer():
    label = np.array([[[1, 1, 0, 0], [-1, 1, 0, 0]]])            # one ignored pixel
    pred = np.array([[[1, 0, 0, 1], [1, 1, 0, 0]]])
    assert batch_pix_accuracy(pred, label) == (5, 7)
    inter, union = batch_intersection_union(pred, label, 2)
    # class 0: pred {(0,1),(0,2),(1,2),(1,3)} vs label {(0,2),(0,3),(1,2),(1,3)} -> inter 3, union 5
    # class 1: pred {(0,0),(0,3),(1,1)} (ignored pixel dropped) vs label {(0,0),(0,1),(1,1)} -> inter 2, union 4
    assert inter.tolist() == [3, 2] and union.tolist() == [5, 4]
    m = SegmentationMetric(2, skip_bg=True)
    m.update(label, pred)
    (n0, acc), (n1, miou) = m.get_name_value()
    assert (n0, n1) == ('accuracy', 'mean-iou')
    assert abs(acc - 5 / 7) < 1e-12 and abs(miou - 0.5) < 1e-12
    m2 = SegmentationMetric(2, skip_bg=False)
    m2.update([label, label], [pred, pred])                       # lists accumulate
    assert abs(m2.get()[1][1] - (0.6 + 0.5) / 2) < 1e-12


def test_logits_are_argmaxed_first_max():
    logits = np.zeros((1, 3, 1, 2), np.float32)
    logits[0, :, 0, 0] = [0.5, 0.5, 0.1]                          # tie -> class 0
    logits[0, :, 0, 1] = [0.1, 0.7, 0.7]                          # tie -> class 1
    label = np.array([[[0, 1]]])
    assert batch_pix_accuracy(logits, label) == (2, 2)

"""Annotated-sample format of the reference: what ``seg_annotator.save_current_results``
writes (seg_annotator.py:322-337) and ``seg_datasets.CollectionDataset`` reads (seg_datasets.py:33-125).

    feat_XXXXXX.pickle   pickle.dump(list of float32 [C_i, H_i, W_i] feature maps)
    img_XXXXXX.jpg       RGB image
    mask_XXXXXX.png      painted mask: white = class 1, gray = class 0, black = ignore

Host-side I/O only (no MXNet): the items come back as numpy arrays in the reference's order and dtypes,
ready for ``SegSolver.predict`` (features) and for decoder training (mask in {1, 0, -1}).
"""
from __future__ import annotations

import os
import pickle
from os.path import join, splitext

import numpy as np


def decode_mask(gray):
    """seg_datasets.py:97-106: gray > 192 -> 1, 64..192 -> 0, < 64 -> -1 (ignore); int32."""
    g = np.asarray(gray)
    out = g.astype(np.int32)
    out[g > 192] = 1
    out[np.logical_and(g >= 64, g <= 192)] = 0
    out[g < 64] = -1
    return out


def encode_mask(labels):
    """Inverse used by the writer: 1 -> 255 (white), 0 -> 128 (gray), -1 -> 0 (black)."""
    lab = np.asarray(labels)
    out = np.zeros(lab.shape, np.uint8)
    out[lab == 1] = 255
    out[lab == 0] = 128
    return out


def save_sample(dst_dir, image_id, img_rgb, features, labels):
    """Writes feat/img/mask files named as seg_annotator.py:327-337 does."""
    import cv2
    os.makedirs(dst_dir, exist_ok=True)
    cv2.imwrite(join(dst_dir, f'img_{image_id:06d}.jpg'), np.asarray(img_rgb)[:, :, ::-1])
    m = encode_mask(labels)
    cv2.imwrite(join(dst_dir, f'mask_{image_id:06d}.png'), np.stack([m, m, m], axis=-1))     # RGB PNG like PIL's
    with open(join(dst_dir, f'feat_{image_id:06d}.pickle'), 'wb') as fp:
        pickle.dump([np.asarray(f, np.float32) for f in features], fp)


class CollectionDataset:
    """Mirror of seg_datasets.CollectionDataset (non-MXNet): lists ``feat_*.pickle`` and loads
    (img float32 [3,H,W], mask int32 [1,H,W] in {1,0,-1}, *features) -- or with ``output_idx`` the
    index first (seg_datasets.py:120-125)."""

    def __init__(self, db_dir, cfg=None, is_validation=False, output_idx=False, max_samples=None,
                 allow_missed_mask=False, load_to_memory=False):
        cfg = cfg or {}
        self._db_dir = db_dir
        self._output_idx = output_idx
        self._allow_missed_mask = allow_missed_mask
        self._preprocess_mask = cfg.get('preprocess_mask', True)
        self._not_ignore_classes = cfg.get('not_ignore_classes', None)
        names = sorted(f for f in os.listdir(db_dir) if splitext(f.lower())[1] == '.pickle' and 'feat' in f)
        if max_samples is not None:
            names = names[:max_samples]
        self._feat_names = names
        self._samples = [self.load_sample(n) for n in names] if load_to_memory else None

    def __len__(self):
        return len(self._feat_names)

    def load_sample(self, feature_name):
        import cv2
        base = splitext(feature_name)[0]
        img = cv2.imread(join(self._db_dir, base.replace('feat', 'img') + '.jpg'))
        img = img[:, :, [2, 1, 0]]                                              # bgr -> rgb (:59)
        mask = cv2.imread(join(self._db_dir, base.replace('feat', 'mask') + '.png'), 0)
        if mask is None and self._allow_missed_mask:
            mask = np.zeros(img.shape[:2])
        assert mask is not None
        with open(join(self._db_dir, feature_name), 'rb') as fp:
            features = pickle.load(fp)
        return mask, img, features

    def get_imname(self, idx):
        return splitext(self._feat_names[idx])[0].replace('feat', 'img') + '.jpg'

    def __getitem__(self, idx):
        mask, img, features = self._samples[idx] if self._samples is not None else self.load_sample(self._feat_names[idx])
        mask = decode_mask(mask) if self._preprocess_mask else mask.astype(np.int32)
        if self._not_ignore_classes is not None:
            mask[np.logical_not(np.isin(mask, self._not_ignore_classes))] = -1
        mask = mask[np.newaxis, :, :]
        img = np.transpose(img.astype(np.float32), (2, 0, 1))
        if self._output_idx:
            return (np.int32(idx), img, mask) + tuple(features)
        return (img, mask) + tuple(features)

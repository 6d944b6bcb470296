"""Synthetic-dataset sweep: what ``python3 main.py generate`` produces (reference main.py:75-104), sharded over
ranks.  For every global sample index i: ``img_{i:06d}.jpg`` (OpenCV default quality, BGR swap, main.py:102) and
``mask_{i:06d}.png`` with class ids {0,1,...} (main.py:103) -- the files
deeplabv3plus/lib/data/segmentation/ffhq_hair_segmentation.py:26,46,66-69 consumes.

The images of index i depend only on (seed, i): any number of GPUs / any batch size writes the same files.
JPEG/PNG encoding runs in a thread pool so that it overlaps the next batch's kernels and copies.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from os.path import join

import numpy as np

from .shard import shard_range, batches


def _write_pair(dst_dir, index, img_rgb, mask):
    import cv2
    cv2.imwrite(join(dst_dir, f'img_{index:06d}.jpg'), img_rgb[:, :, ::-1])      # main.py:102 (RGB -> BGR)
    cv2.imwrite(join(dst_dir, f'mask_{index:06d}.png'), mask)                    # main.py:103 (class ids)


def generate_dataset(pipeline, dst_dir, n_total, seed=0, psi=None, rank=0, world=1, workers=8, progress=None, write=True):
    """pipeline: networks.GeneratePipeline (batch = pipeline.n).  Writes this rank's shard of range(n_total).
    ``write=False`` runs the same sweep (kernels + device-to-host copies) without encoding (measures the producer alone).
    Returns the number of samples written by this rank."""
    os.makedirs(dst_dir, exist_ok=True)
    lo, hi = shard_range(n_total, rank, world)
    B = pipeline.n
    written = 0
    pending = []            # (slot, first, n) whose device->host copies are in flight
    with ThreadPoolExecutor(max_workers=workers) as pool:
        futures = []

        def drain(entry):
            slot, first, n = entry
            pipeline.wait()
            imgs = pipeline.img_host[slot].numpy()
            masks = pipeline.mask_host[slot].numpy()
            for k in range(n if write else 0):
                # copy out of the pinned slot: it is reused two batches later
                futures.append(pool.submit(_write_pair, dst_dir, first + k, imgs[k].copy(), masks[k].copy()))

        for first, n in batches(lo, hi, B):
            # a short last batch still runs at the pipeline's batch size; the surplus samples are dropped
            slot = pipeline.run(None, psi=psi, seed=seed, first_sample=first)
            if pending:
                drain(pending.pop(0))
            pending.append((slot, first, n))
            written += n
            if progress is not None:
                progress(n)
        while pending:
            drain(pending.pop(0))
        for f in futures:
            f.result()
    return written

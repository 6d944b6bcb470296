"""Annotated-sample format: writer / reader round trip and the mask thresholds of
seg_datasets.py:97-106."""
import numpy as np

from gan_segmentation_b200.seg_datasets import CollectionDataset, decode_mask, encode_mask, save_sample


def test_mask_thresholds():
    g = np.array([[0, 63, 64, 128, 192, 193, 255]], np.uint8)
    assert decode_mask(g).tolist() == [[-1, -1, 0, 0, 0, 1, 1]]
    lab = np.array([[1, 0, -1]])
    assert decode_mask(encode_mask(lab)).tolist() == lab.tolist()


def test_roundtrip(tmp_path):
    rs = np.random.RandomState(0)
    feats = [rs.randn(c, 4 << i, 4 << i).astype(np.float32) for i, c in enumerate([512, 512, 256])]
    img = rs.randint(0, 255, (16, 16, 3)).astype(np.uint8)
    lab = rs.randint(-1, 2, (16, 16))
    save_sample(str(tmp_path), 7, img, feats, lab)
    assert sorted(p.name for p in tmp_path.iterdir()) == ['feat_000007.pickle', 'img_000007.jpg', 'mask_000007.png']
    ds = CollectionDataset(str(tmp_path), {'preprocess_mask': True, 'not_ignore_classes': None}, output_idx=True)
    assert len(ds) == 1 and ds.get_imname(0) == 'img_000007.jpg'
    idx, im, mask, *f = ds[0]
    assert idx == 0 and im.shape == (3, 16, 16) and im.dtype == np.float32
    assert mask.shape == (1, 16, 16) and mask.dtype == np.int32 and np.array_equal(mask[0], lab)
    assert len(f) == 3 and all(np.array_equal(a, b) for a, b in zip(f, feats))

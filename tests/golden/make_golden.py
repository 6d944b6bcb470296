"""Regenerates tests/golden/oracle_res5_seed0.npz from the oracle (the reference itself cannot run
offline: MXNet is not installable here).  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import gan_segmentation_b200  # noqa: E402,F401
from parity_util import make_case  # noqa: E402
from oracle import generate_oracle as O  # noqa: E402

gc, dc, gp, dp, z, noise = make_case(5, 2, seed=0)
out = O.generate(gp, gc, dp, dc, z, noise)
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'oracle_res5_seed0.npz'),
                    img_f32=out['img_f32'], img_u8=out['img_u8'], logits=out['logits'], mask=out['mask'],
                    **{f'feat{i}': f[:, :8] for i, f in enumerate(out['features'])})
print('written')

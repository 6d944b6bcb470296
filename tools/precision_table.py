import sys
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import precision_study as S
from oracle import generate_oracle as O
from parity_util import make_case, psnr
torch.set_num_threads(8)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 41
gc, dc, gp, dp, z, noise = make_case(10, 1, seed=seed)
K = ['w', 'raw', 't1', 'u1', 't2', 'x']
class Only(dict):
    pass
with torch.no_grad():
    ref, rf = O.generator_forward(gp, gc, z, noise, 0.7)
    b = ref.clamp(-1, 1).numpy()
    print('error energy (mean sq err x 1e9) of the image when ONLY (key, res) is rounded to fp16; max-abs x1e3 in brackets')
    print('res ' + ' '.join(f'{k:>14s}' for k in K))
    orig_emulate = S.emulate
    for r in range(2, 11):
        row = []
        for k in K:
            # active only at res r: emulate uses r >= lo; do difference trick: run with lo=r and lo=r+1 is not additive -> patch q
            spec = {k: r}
            # patch: make rounding active only when res == r by wrapping dict.get
            class D(dict):
                def get(self, key, default=None):
                    return self[key] if key in self else default
            img, _ = S.emulate(gp, gc, z, noise, 0.7, D(spec), only_res=r)
            a = img.clamp(-1, 1).numpy(); d = a - b
            row.append(f'{float((d**2).mean())*1e9:7.1f} [{np.abs(d).max()*1e3:4.1f}]')
        print(f'{r:3d} ' + ' '.join(row), flush=True)

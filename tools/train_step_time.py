"""Time of one decoder training step on random features: the resident step (gsx_train_step) and, with --hooks, the
round-1 hook-based orchestration for comparison.
  python tools/train_step_time.py [res_log2=10] [batch=1] [steps=10] [--hooks]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gan_segmentation_b200.config import decoder_config
from gan_segmentation_b200.decoder_training import CudaBackend, DecoderTrainer, ResidentTrainer
from gan_segmentation_b200.random_init import init_decoder_params

args = [a for a in sys.argv[1:] if not a.startswith('--')]
res = int(args[0]) if len(args) > 0 else 10
n = int(args[1]) if len(args) > 1 else 1
steps = max(1, int(args[2])) if len(args) > 2 else 10
cfg = dict(decoder_config(res), use_dropout=True)
params = init_decoder_params(cfg, seed=2)
g = torch.Generator(device='cuda').manual_seed(0)
feats = [torch.randn((n, c, 4 << i, 4 << i), generator=g, device='cuda') for i, c in enumerate(cfg['in_channels'][:res - 1])]
mask = torch.randint(-1, 2, (n, 1, 4 << (res - 2), 4 << (res - 2)), generator=g, device='cuda')
if '--hooks' in sys.argv:
    drops = [(torch.rand((n, cfg['features'][i], 4 << i, 4 << i), generator=g, device='cuda') > 0.5).float() for i in range(res - 1)]
    tr = DecoderTrainer(cfg, params, CudaBackend())
    step = lambda k: tr.step(feats, mask, drops)
else:
    tr = ResidentTrainer(cfg, params, n, use_graph='--nograph' not in sys.argv)
    step = lambda k: tr.step(feats, mask, dropout_seed=k)
losses = []
for k in range(steps + 4):
    if k == 4:
        torch.cuda.synchronize(); t0 = time.time()
    l = step(k)
    if k in (0, steps + 3):
        losses.append(float(l.mean()))
torch.cuda.synchronize()
dt = (time.time() - t0) / steps
print(f'decoder training step ({"hooks" if "--hooks" in sys.argv else "resident"}), res 2^{res}, batch {n}: {dt * 1e3:.2f} ms/step '
      f'({n / dt:.1f} samples/s), loss {losses[0]:.4f} -> {losses[-1]:.4f}, peak torch memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB')

"""Time of one decoder training step (first, hook-based version) on random features.
  python tools/train_step_time.py [res_log2=10] [batch=1] [steps=3]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gan_segmentation_b200.config import decoder_config
from gan_segmentation_b200.decoder_training import CudaBackend, DecoderTrainer
from gan_segmentation_b200.random_init import init_decoder_params

res = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = decoder_config(res)
params = init_decoder_params(cfg, seed=2)
g = torch.Generator(device='cuda').manual_seed(0)
feats = [torch.randn((n, c, 4 << i, 4 << i), generator=g, device='cuda') for i, c in enumerate(cfg['in_channels'][:res - 1])]
mask = torch.randint(-1, 2, (n, 1, 4 << (res - 2), 4 << (res - 2)), generator=g, device='cuda')
drops = [(torch.rand((n, cfg['features'][i], 4 << i, 4 << i), generator=g, device='cuda') > 0.5).float() for i in range(res - 1)]
tr = DecoderTrainer(cfg, params, CudaBackend())
losses = []
for k in range(steps + 1):
    if k == 1:
        torch.cuda.synchronize(); t0 = time.time()
    losses.append(float(tr.step(feats, mask, drops).mean()))
torch.cuda.synchronize()
dt = (time.time() - t0) / steps
print(f'decoder training step, res 2^{res}, batch {n}: {dt * 1e3:.1f} ms/step ({n / dt:.2f} samples/s), loss {losses[0]:.4f} -> {losses[-1]:.4f}, '
      f'peak memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB (torch tensors only)')

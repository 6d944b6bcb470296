"""Experiment: two independent generate steps in flight on two streams (each with its own generator / decoder handles and
workspaces) against one stream.  python tools/two_streams.py [workload] [steps]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gan_segmentation_b200 import _lib as L
from gan_segmentation_b200.networks import GeneratePipeline
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'ffhq1024']
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device('cuda', 0)
B = wl['batch']
sets = []
for k in range(2):
    gc, dc, gp, dp, G, D = bench.build_models(wl, 'fp16', dev)
    pipe = GeneratePipeline(G, D, B, overlap=False)
    H, W = G.out_hw
    sets.append(dict(G=G, D=D, pipe=pipe, img=torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev),
                     mask=torch.empty((B, H, W), dtype=torch.uint8, device=dev), st=torch.cuda.Stream(),
                     z=torch.randn((B, 512), device=dev)))
psi = np.full((sets[0]['G'].num_layers,), wl['psi'], np.float32)
lib = sets[0]['G']._lib

def step(s, i):
    sp = C.c_void_p(s['st'].cuda_stream)
    L.check(lib.gsx_generate_dev(s['G']._h, s['D']._h, B, L.ptr(s['z']), L.np_ptr(psi), 7, i * B, L.ptr(s['img']), L.ptr(s['mask']),
                                 L.ptr(s['pipe'].gws), s['pipe'].gws.numel(), L.ptr(s['pipe'].dws), s['pipe'].dws.numel(), sp), 'generate')

for nstreams in (1, 2, 1, 2):
    for i in range(6):
        step(sets[i % nstreams], i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in sets[:nstreams]:
        s['st'].wait_stream(torch.cuda.current_stream())
    for i in range(steps):
        step(sets[i % nstreams], i)
    for s in sets[:nstreams]:
        torch.cuda.current_stream().wait_stream(s['st'])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f'{wl["name"][:30]} streams={nstreams}: {steps * B / ms * 1e3:.1f} samples/s, {ms / steps:.4f} ms per step', flush=True)

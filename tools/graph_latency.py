"""Small-batch latency: host-paced launches vs CUDA-graph replay of the same generate step.
  python tools/graph_latency.py [bedrooms256|ffhq1024] [batch]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_segmentation_b200.config import generator_config, decoder_config
from gan_segmentation_b200.networks import Generator, Decoder, GraphedGenerate
from gan_segmentation_b200.random_init import init_generator_params, init_decoder_params

name = sys.argv[1] if len(sys.argv) > 1 else 'bedrooms256'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
res = {'bedrooms256': 8, 'ffhq1024': 10}[name]
gc, dc = generator_config(res), decoder_config(res)
G = Generator(gc); G.set_parameters(init_generator_params(gc, seed=0))
D = Decoder(dc); D.set_parameters(init_decoder_params(dc, seed=2))


def plain(k):
    G.forward(n=n, seed=3, first_sample=k * n, return_u8=True, return_image=False, return_features=False)
    D.forward(generator=G, return_logits=False)


def timed(fn, iters=200):
    for k in range(20):
        fn(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(iters):
        fn(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


t_plain = timed(plain)
gg = GraphedGenerate(G, D, n, seed=3)
t_graph = timed(lambda k: gg.replay())
print(f'{name} batch {n}: host-paced launches {t_plain:.3f} ms/step, CUDA-graph replay {t_graph:.3f} ms/step '
      f'({n / t_graph * 1e3:.0f} samples/s)')

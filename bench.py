#!/usr/bin/env python
"""Benchmark of the generate hot path: image+mask samples/sec, StyleGAN-FFHQ 1024^2 generator + decoder.

  python bench.py --gpus N --steps K --warmup W             (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                      (CPU baseline arm: the oracle on host cores)

One step = one batch of 32 latents per GPU -> uint8 image [1024,1024,3] + uint8 mask [1024,1024] per
latent (BASELINE.json configs[1]).  Latents are index-sharded over the ranks, no collective on the
data path (weak scaling).  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'image+mask samples/sec'
UNIT = 'samples/s'
WORKLOADS = {
    'ffhq1024': dict(gan='ffhq', max_res_log2=10, base=(4, 4), batch=32, psi=0.7,
                     name='StyleGAN-FFHQ 1024^2 generator + hair decoder forward, batch 32/GPU, psi=0.7 (BASELINE configs[1])'),
    'cars512x384': dict(gan='cars', max_res_log2=9, base=(3, 4), batch=64, psi=0.7,
                        name='StyleGAN-cars 512x384 generator + decoder forward, batch 64/GPU (BASELINE configs[2])'),
    'bedrooms256': dict(gan='bedrooms', max_res_log2=8, base=(4, 4), batch=1, psi=1.0,
                        name='StyleGAN-bedrooms 256^2 generator + decoder forward, batch 1 (BASELINE configs[0])'),
    # BASELINE configs[3]: decoder training, generator frozen; handled by run_train (metric: training samples/sec)
    'ffhq_train': dict(gan='ffhq', max_res_log2=10, base=(4, 4), batch=1, psi=0.7, train=True,
                       name='FFHQ hair decoder training on 20 synthetic annotated samples, generator frozen, batch 1/GPU, '
                            'one decoder-gradient all-reduce per step (BASELINE configs[3])'),
}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d['hbm_gbs']), tensor=float(d.get('bf16_tflops_sustained', d['bf16_tflops'])), src='measured')
    return dict(hbm=6650.0, tensor=1400.0, src='fallback')      # B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '50', '-i', str(gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.06)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def pin_to_gpu_numa_node(gpu):
    """Runs this rank (and therefore its pinned host buffers, first-touch) on the CPUs NVML reports as local to the GPU:
    at 8 ranks the end-to-end path moves 8 x 128 MiB per step over PCIe, buffers on the far socket cost bandwidth."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def build_models(wl, dtype, device):
    import gan_segmentation_b200  # noqa: F401
    from gan_segmentation_b200.config import generator_config, decoder_config
    from gan_segmentation_b200.random_init import init_generator_params, init_decoder_params
    from gan_segmentation_b200.networks import Generator, Decoder
    gc = generator_config(wl['max_res_log2'], *wl['base'])
    dc = decoder_config(wl['max_res_log2'])
    gp = init_generator_params(gc, seed=0)
    dp = init_decoder_params(dc, seed=2)
    G = Generator(gc, device=device, dtype=dtype)
    G.set_parameters(gp)
    D = Decoder(dc, base_hw=wl['base'], device=device, dtype=dtype)
    D.set_parameters(dp)
    return gc, dc, gp, dp, G, D


def cpu_oracle_rate(wl, n_samples, warm=1, variants=False):
    """The oracle (PyTorch-CPU restatement of the reference; MXNet itself is not installable offline) on the
    host cores: image+mask samples/s at batch 1 (main.py:97-99 decodes one sample at a time)."""
    import numpy as np
    import torch
    from gan_segmentation_b200.config import generator_config, decoder_config, noise_shapes
    from gan_segmentation_b200.random_init import init_generator_params, init_decoder_params
    from oracle import generate_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gc = generator_config(wl['max_res_log2'], *wl['base'])
    dc = decoder_config(wl['max_res_log2'])
    gp = init_generator_params(gc, seed=0)
    dp = init_decoder_params(dc, seed=2)
    rs = np.random.RandomState(0)
    times = []
    for i in range(warm + n_samples):
        z = rs.randn(1, 512).astype(np.float32)
        noise = [rs.randn(*s).astype(np.float32) for s in noise_shapes(gc, 1)]
        t0 = time.perf_counter()
        O.generate(gp, gc, dp, dc, z, noise, psi=wl['psi'])
        t = time.perf_counter() - t0
        if i >= warm:
            times.append(t)
    out = dict(value=1.0 / statistics.median(times), unit=UNIT, cores=cores, kind='port',
               sample=f'{n_samples} latents at batch 1 after {warm} warm-up (median), oracle/generate_oracle.py '
                      f'(PyTorch CPU fp32, {cores} threads); the MXNet reference is not installable offline')
    if variants:
        # the reference's default generator batch (8, config.yml.example:5), compute only, and "as the reference
        # drives it": generator at batch 8, every feature map through host numpy (image_generator.py:103-114), then the
        # decoder one sample at a time (main.py:97-99).  One timed pass each (bounded sample).
        import torch as _t
        z = rs.randn(8, 512).astype(np.float32)
        noise = [rs.randn(*s).astype(np.float32) for s in noise_shapes(gc, 8)]
        t0 = time.perf_counter()
        O.generate(gp, gc, dp, dc, z, noise, psi=wl['psi'])
        out['batch8_value'] = 8.0 / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        with _t.no_grad():
            img, feats = O.generator_forward(gp, gc, z, noise, wl['psi'])
            img_u8 = O.transform_gan_back(img.numpy())
            host = [f.numpy().copy() for f in feats]                        # .asnumpy() of the 9 feature maps
            for k in range(8):
                per = [_t.from_numpy(np.ascontiguousarray(h[k:k + 1])) for h in host]
                O.argmax_mask(O.decoder_forward(dp, dc, per).numpy())
        out['reference_driver_value'] = 8.0 / (time.perf_counter() - t0)
        out['variants'] = 'batch8_value: batch 8 compute only; reference_driver_value: generator at batch 8, features via host numpy, decoder per sample (main.py:94-99); one pass each'
    return out, times


def run_train(args, wl, rank, world, local):
    """BASELINE configs[3]: decoder training on 20 synthetic annotated samples (features from the frozen generator at
    seeds 0..19, masks = disk 1 / ring 0 / rest -1), per-GPU batch B, one all-reduce of the flat gradient bucket per step.
    value = samples/s over all ranks, device-timed; the all-reduce is timed separately with its own events."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from gan_segmentation_b200 import _lib as L
    from gan_segmentation_b200.decoder_training import ResidentTrainer
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    B = wl['batch']
    steps, warm = max(1, args.steps), max(4, args.warmup)
    gc, dc, gp, dp, G, D = build_models(wl, args.dtype, device)
    n_samples = 20
    # the 20 annotated samples: features of the frozen generator kept on the device as fp32 NCHW (what the reference's
    # feat_*.pickle files hold, 127 MiB each), synthetic 3-valued masks
    feats, masks = [], []
    H, W = G.out_hw
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(n_samples):
        out = G.forward(n=1, seed=11, first_sample=i, psi=wl['psi'], return_image=False, return_features=True)
        feats.append([f.clone() for f in out['features']])
        rr = np.hypot(yy - H * (0.4 + 0.01 * i), xx - W * 0.5)
        masks.append(torch.from_numpy(np.where(rr < 0.25 * H, 1, np.where(rr < 0.4 * H, 0, -1)).astype(np.int32)).to(device))
    del G
    cfg = dict(dc, use_dropout=True, base_lr=1e-4)
    tr = ResidentTrainer(cfg, dp, B, device=device, base_hw=wl['base'], dtype=args.dtype)
    stream = torch.cuda.current_stream()
    order = np.random.RandomState(1).permutation(n_samples)

    def batch(i):
        idx = [int(order[((i * world + rank) * B + k) % n_samples]) for k in range(B)]
        f = [torch.cat([feats[j][l] for j in idx], 0) if B > 1 else feats[idx[0]][l] for l in range(len(feats[0]))]
        m = torch.stack([masks[j] for j in idx], 0)
        return f, m

    ar_ms = []

    def step(i, timed=False):
        f, m = batch(i)
        tr.forward_backward(f, m, dropout_seed=(i * world + rank))
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            dist.all_reduce(tr.g, op=dist.ReduceOp.SUM)
            e1.record(stream)
            if timed:
                ar_ms.append((e0, e1))
        tr.t += 1
        L.check(tr.lib.gsx_adam_step(L.ptr(tr.p), L.ptr(tr.g), L.ptr(tr.m), L.ptr(tr.v), tr.n_learn, tr.t, tr.lr, tr.beta1, tr.beta2,
                                     tr.eps, tr.wd, 1.0 / (B * world * tr.grad_scale), C.c_void_p(stream.cuda_stream)), 'adam', args.dtype)

    for i in range(warm):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = float(tr.loss.mean())
    e0.record(stream)
    for i in range(steps):
        step(warm + i, timed=True)
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = L.launch_count() - n0
    if rank == 0:
        ar = [a.elapsed_time(b) for a, b in ar_ms]
        nbytes = tr.n_learn * 4
        line = dict(metric='decoder training samples/sec', value=steps * B * world / (ms * 1e-3), unit=UNIT, n_gpus=world, steps=steps,
                    warmup=warm, ms_per_step=ms / steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype=args.dtype,
                    data='synthetic',
                    config=dict(workload=wl['name'], batch_per_gpu=B, global_batch=B * world, samples=n_samples,
                                features='resident in HBM as fp32 NCHW (the reference reloads a 127 MiB pickle per sample)',
                                optimizer='Adam 1e-4, rescale 1/global batch', dropout='Philox, in-kernel',
                                parallelism=f'data parallel over {world} GPU(s): one all-reduce of the flat {nbytes / 2**20:.1f} MiB gradient bucket per step',
                                graph='gsx_train_step replayed from a CUDA graph'),
                    clocks=clocks, gpu_launches=int(launches) * world if not tr.use_graph else None,
                    gpu_launches_note='the step is replayed from a CUDA graph (~380 kernels of this library per step); only the Adam kernel is launched directly',
                    allreduce=dict(bytes=nbytes, ms_median=statistics.median(ar) if ar else 0.0, ms_max=max(ar) if ar else 0.0),
                    loss=dict(first=l0, last=float(tr.loss.mean())),
                    e2e=dict(value=steps * B * world / (ms * 1e-3), unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0,
                             note='training consumes device-resident annotated samples; there is no per-step host transfer on this path'))
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def make_config(wl, B, world):
    """The workload description shared by both arms (the driver compares the two lines' config)."""
    return dict(workload=wl['name'], batch_per_gpu=B, global_batch=B * world, psi=wl['psi'],
                noise='on-device Philox4x32-10 keyed by (seed, global sample index, layer)',
                weights='random init of the named architecture (seed 0 / 2), pretrained .params unavailable offline',
                parallelism=f'latent-index sharding over {world} GPU(s), no collective',
                l2='per-step working set (several GiB of activations) exceeds the 126 MB L2: inputs larger than L2')


def run_reference_train(args, wl):
    """CPU arm of the training workload: oracle/train_oracle.py (fp32 autograd restatement of seg_solver.py:386-421) on the
    host cores, batch 1, random features of the generator's shapes."""
    import numpy as np
    import torch
    from gan_segmentation_b200.config import decoder_config
    from gan_segmentation_b200.random_init import init_decoder_params
    from oracle import train_oracle as T
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dc = dict(decoder_config(wl['max_res_log2']), use_dropout=False, base_lr=1e-4)
    dp = init_decoder_params(dc, seed=2)
    rs = np.random.RandomState(0)
    by, bx = wl['base']
    feats = [rs.randn(1, c, by << i, bx << i).astype(np.float32) for i, c in enumerate(dc['in_channels'])]
    H, W = by << (len(feats) - 1), bx << (len(feats) - 1)
    mask = rs.randint(-1, 2, (1, 1, H, W)).astype(np.int64)
    steps = max(1, min(args.steps, 3))
    times, state = [], None
    for i in range(1 + steps):
        t0 = time.perf_counter()
        dp, state, loss, _ = T.train_step(dp, dc, feats, mask, state)
        if i >= 1:
            times.append(time.perf_counter() - t0)
    v = 1.0 / statistics.median(times)
    base = dict(value=v, unit=UNIT, cores=cores, kind='port',
                sample=f'{steps} training steps at batch 1 after 1 warm-up (median), oracle/train_oracle.py (PyTorch CPU fp32 autograd, {cores} threads)')
    emit(dict(metric='decoder training samples/sec', value=v, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=1,
              ms_per_step=1e3 * statistics.mean(times), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
              data='synthetic', impl='reference', config=dict(workload=wl['name']), cpu_baseline=base, gpu_launches=0,
              e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0)))


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    if wl.get('train'):
        run_reference_train(args, wl)
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    steps = min(steps, 8)                                  # bounded: ~2 s per 1024^2 sample on 8 cores
    base, times = cpu_oracle_rate(wl, steps, warm=min(warm, 1))
    ms = 1e3 * statistics.mean(times)
    line = dict(metric=METRIC, value=base['value'], unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=min(warm, 1),
                ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                impl='reference', config=make_config(wl, wl['batch'], max(1, args.gpus)),
                cpu_baseline=dict(base), gpu_launches=0,
                e2e=dict(value=base['value'], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else libraries print (NCCL banner, warnings)
    was redirected to stderr at start-up."""
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # stray library output (e.g. "NCCL version ...") -> stderr
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='ffhq1024', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--dtype', default=os.environ.get('GSX_DTYPE', 'fp16'), choices=['fp16', 'bf16'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--layers-out', default='', help='write the per-layer roofline table (TSV) here')
    ap.add_argument('--opt', action='append', default=[], help='library option name=value (gsx_set_option), for A/B runs')
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl['batch'] = args.batch
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, wl, rank, world)
        return
    if wl.get('train'):
        run_train(args, wl, rank, world, local)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from gan_segmentation_b200 import _lib as L
    from gan_segmentation_b200.networks import GeneratePipeline
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the generate path has no CPU fallback')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    pin_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    steps, warm = max(1, args.steps), max(3, args.warmup)
    B = wl['batch']
    for o in args.opt:
        k, v = o.split('=')
        L.check(L.lib(args.dtype).gsx_set_option(k.encode(), int(v)), f'gsx_set_option({o})', args.dtype)
    gc, dc, gp, dp, G, D = build_models(wl, args.dtype, device)
    pipe = GeneratePipeline(G, D, B)
    lib = G._lib
    H, W = G.out_hw
    img_dev = torch.empty((B, H, W, 3), dtype=torch.uint8, device=device)
    mask_dev = torch.empty((B, H, W), dtype=torch.uint8, device=device)
    z_dev = torch.randn((B, 512), generator=torch.Generator(device=device).manual_seed(1234 + rank), device=device)
    psi = np.full((G.num_layers,), wl['psi'], np.float32)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def step_device(i):
        # latents already resident in HBM; noise from the on-device Philox stream of the global sample index
        first = (i * world + rank) * B
        L.check(lib.gsx_generate_dev(G._h, D._h, B, L.ptr(z_dev), L.np_ptr(psi), 7, first, L.ptr(img_dev), L.ptr(mask_dev),
                                     L.ptr(pipe.gws), pipe.gws.numel(), L.ptr(pipe.dws), pipe.dws.numel(), sp), 'generate', args.dtype)

    z_host = np.random.RandomState(99 + rank).randn(B, 512).astype(np.float32)

    def step_host(i):
        pipe.run(z_host, psi=psi, seed=7, first_sample=(i * world + rank) * B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()          # device-wide: drains the copy stream of the e2e pipeline too

    def timed(fn):
        for i in range(warm):
            fn(i)
        barrier()
        n0 = L.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(warm + i)
        if pipe.copy_stream is not None:
            stream.wait_stream(pipe.copy_stream)      # the last D2H copies belong to the timed region
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = L.launch_count() - n0
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches

    sampler = ClockSampler(local) if rank == 0 else None
    ms_total, launches = timed(step_device)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _ = timed(step_host)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = steps * B * world / (ms_total * 1e-3)
    e2e_value = steps * B * world / (ms_e2e * 1e-3)

    # ---- per-layer profile of one step (outside the timed region) -> dominant kernel + roofline
    pk = peaks()
    lib.gsx_profile_enable(1)
    step_device(0)
    buf = C.create_string_buffer(1 << 20)
    L.check(lib.gsx_profile_dump(buf, len(buf)), 'profile_dump', args.dtype)
    lib.gsx_profile_enable(0)
    rows = []
    for ln in buf.value.decode().strip().split('\n'):
        lab, ms, by, fl, fx, kern = (ln.split('\t') + ['', ''])[:6]
        rows.append(dict(label=lab, ms=float(ms), bytes=float(by), flops=float(fl), flops_exec=float(fx or 0), kernel=kern))
    agg = {}
    for r in rows:
        a = agg.setdefault(r['label'], dict(label=r['label'], ms=0.0, bytes=0.0, flops=0.0, flops_exec=0.0, launches=0, kernel=r['kernel']))
        a['ms'] += r['ms']; a['bytes'] += r['bytes']; a['flops'] += r['flops']; a['flops_exec'] += r['flops_exec']; a['launches'] += 1
    ridge = pk['tensor'] * 1e12 / (pk['hbm'] * 1e9)
    table = []
    for a in agg.values():
        if a['ms'] <= 0:
            continue
        ai = a['flops'] / a['bytes'] if a['bytes'] else 0.0
        bound = 'tensor' if ai > ridge else 'hbm'
        ach = a['flops'] / (a['ms'] * 1e-3) / 1e12 if bound == 'tensor' else a['bytes'] / (a['ms'] * 1e-3) / 1e9
        peak = pk['tensor'] if bound == 'tensor' else pk['hbm']
        # tensor-class layers: also the fraction on the flops the tensor cores actually execute (phase-decomposed up-convs
        # execute 2.25x fewer than the dense-equivalent count)
        frac_exec = a['flops_exec'] / (a['ms'] * 1e-3) / 1e12 / pk['tensor'] if a['flops_exec'] else 0.0
        table.append(dict(a, bound=bound, achieved=ach, peak=peak, frac=ach / peak, frac_exec=frac_exec,
                          unit='TFLOP/s' if bound == 'tensor' else 'GB/s'))
    table.sort(key=lambda t: -t['ms'])
    step_ms_profiled = sum(t['ms'] for t in table)
    if args.layers_out:
        with open(args.layers_out, 'w') as f:
            f.write('label\tkernel\tlaunches\tms\tshare\tbound\talg_GB\talg_GFLOP\texec_GFLOP\tachieved\tunit\tfrac_of_%s_peak\ttensor_frac_on_executed_flops\n' % pk['src'])
            for t in table:
                f.write(f"{t['label']}\t{t['kernel'] or '-'}\t{t['launches']}\t{t['ms']:.4f}\t{t['ms'] / step_ms_profiled:.4f}\t{t['bound']}\t"
                        f"{t['bytes'] / 1e9:.4f}\t{t['flops'] / 1e9:.2f}\t{t['flops_exec'] / 1e9:.2f}\t{t['achieved']:.1f}\t{t['unit']}\t{t['frac']:.3f}\t"
                        f"{t['frac_exec']:.3f}\n")
    top = table[0]
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload, {}).get(top['label'])
    kshare = sum(t['ms'] for t in table if t['kernel'] == top['kernel']) / step_ms_profiled if top['kernel'] else None
    roofline = dict(kernel=f"{top['kernel']} [{top['label']}]" if top['kernel'] else top['label'],
                    bound=top['bound'], achieved=top['achieved'], peak=top['peak'], unit=top['unit'], frac=top['frac'],
                    traffic=traffic, peak_source=pk['src'], share_of_step=top['ms'] / step_ms_profiled,
                    kernel_share_of_step=kshare,
                    note='achieved = algorithmic bytes (or dense-equivalent flops) of the dominant LAUNCH / its CUDA-event time; '
                         'share_of_step is that launch, kernel_share_of_step all launches of the same __global__ function',
                    whole_step=dict(hbm_frac=sum(t['bytes'] for t in table if t['bound'] == 'hbm') / 1e9 /
                                    (sum(t['ms'] for t in table if t['bound'] == 'hbm') * 1e-3 + 1e-12) / pk['hbm'],
                                    tensor_frac=sum(t['flops'] for t in table if t['bound'] == 'tensor') / 1e12 /
                                    (sum(t['ms'] for t in table if t['bound'] == 'tensor') * 1e-3 + 1e-12) / pk['tensor']))

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu, _ = cpu_oracle_rate(wl, 4 if wl['max_res_log2'] >= 10 else 10, variants=True)

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=steps, warmup=warm,
                ms_per_step=ms_total / steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype=args.dtype, data='synthetic',
                config=make_config(wl, B, world), workspace_gib=(pipe.gws.numel() + pipe.dws.numel()) / 2 ** 30,
                clocks=clocks, gpu_launches=launches * world,
                e2e=dict(value=e2e_value, unit=UNIT, ms_per_step=ms_e2e / steps, h2d_bytes_per_step=pipe.h2d_bytes * world,
                         d2h_bytes_per_step=pipe.d2h_bytes * world,
                         api='gsx_generate_host (pinned host z in, uint8 image + mask out to pinned host; D2H on a second stream, double-buffered)'),
                roofline=roofline, cpu_baseline=cpu)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

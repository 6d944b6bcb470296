"""Command-line entry point with the reference's actions and ``config.yml`` keys (reference main.py:15-104,
config.yml.example:1-8):

    python -m gan_segmentation_b200.main {train,evaluate,generate} [--config config.yml]
    python -m torch.distributed.run --nproc-per-node N -m gan_segmentation_b200.main generate     (one process per GPU)

``config.yml`` keys, as the reference reads them (main.py:32-43): BASE_DIR, GAN ('ffhq' | 'cars' | 'bedrooms'),
GAN_DIR, GAN_GPU_IDS, GAN_BATCH_SIZE_PER_GPU, SOLVER_GPU_IDS, ANNOTATION, NO_GAN, IMGS_DIR, GENERATE_NUM.

  train     SegSolver(max_res_log2, BASE_DIR/data, BASE_DIR/checkpoints, SOLVER_GPU_IDS, keep_weights=False).fit()   (:56-62)
  evaluate  ... .evaluate(BASE_DIR/eval), prints "accuracy: .., mean-iou: .., total-loss: .."                        (:63-74)
  generate  GENERATE_NUM image + mask pairs -> BASE_DIR/dataset/train_generated/img_XXXXXX.jpg, mask_XXXXXX.png      (:75-104)
            on the fused device-resident path (features never leave HBM); under torchrun the global index range is
            sharded over the ranks and every rank writes its own files (the file of index i does not depend on the split)
  annotation  the Tk GUI (seg_annotator.py) is outside this package's scope (out of scope)

Without a GPU every action fails: there is no CPU path.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from os.path import join

import numpy as np

MAX_RES_LOG2 = {'ffhq': 10, 'cars': 9, 'bedrooms': 8}          # main.py:57, 64, 77


def parse_args(argv=None):
    p = argparse.ArgumentParser(prog='gan_segmentation_b200.main')
    p.add_argument('action', nargs='?', choices=('annotation', 'train', 'evaluate', 'generate'), default='annotation')
    p.add_argument('--config', default='config.yml', help="the reference reads 'config.yml' from the working directory")
    p.add_argument('--random-init', action='store_true',
                   help='random weights of the named architecture instead of GAN_DIR/stylegan-GAN.params / the decoder '
                        'checkpoint (benchmarks: the pretrained files are not available offline)')
    p.add_argument('--psi', type=float, default=None, help='truncation psi override (default: the parameter file)')
    p.add_argument('--encode-workers', type=int, default=0, help='JPEG/PNG encoder threads per rank (0: host cores / ranks)')
    p.add_argument('--no-write', action='store_true', help='generate: run the sweep but skip encoding and file output')
    return p.parse_args(argv)


def load_config_file(path):
    """utils.load_config_file of the reference: a YAML mapping."""
    import yaml
    with open(path) as f:
        cfg = yaml.safe_load(f)
    if not isinstance(cfg, dict):
        raise ValueError(f'{path}: expected a mapping of the reference config keys')
    return cfg


def _dist():
    """(rank, world, local_rank); initialises torch.distributed when launched by torchrun."""
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return rank, world, local


def _solver(cfg, gan, gpu_ids, random_init):
    from .seg_solver import SegSolver
    root_dir = cfg['BASE_DIR']
    solver = SegSolver(MAX_RES_LOG2[gan], join(root_dir, 'data'), join(root_dir, 'checkpoints'), gpu_ids=gpu_ids,
                       keep_weights=False, verbose=int(os.environ.get('RANK', 0)) == 0)
    if random_init and not solver.is_trained:
        from .random_init import init_decoder_params
        solver.set_parameters(init_decoder_params(solver.cfg, seed=2))
        solver.is_trained = True
    return solver


def run_generate(cfg, args, rank, world, local):
    """main.py:75-104 on the fused path.  Returns (samples written by this rank, seconds)."""
    import torch
    from .config import generator_config
    from .dataset_writer import generate_dataset
    from .networks import Generator, GeneratePipeline
    gan = cfg['GAN']
    gpu_ids = [local] if world > 1 else list(cfg['GAN_GPU_IDS'])[:1] or [0]
    solver = _solver(cfg, gan, gpu_ids if world > 1 else list(cfg['SOLVER_GPU_IDS'])[:1] or gpu_ids, args.random_init)
    if not solver.is_trained:
        print('train Decoder first!')                       # main.py:82-84
        raise SystemExit(-1)
    dev = torch.device('cuda', gpu_ids[0])
    torch.cuda.set_device(dev)
    gc = generator_config(MAX_RES_LOG2[gan])               # the shipped driver hard-codes the 4x4 base (image_generator.py:58-59)
    G = Generator(gc, device=dev)
    if args.random_init:
        from .random_init import init_generator_params
        G.set_parameters(init_generator_params(gc, seed=0))
    else:
        G.load_parameters(f"{cfg['GAN_DIR']}/stylegan-{gan}.params")     # image_generator.py:21-22
    D = solver.nets[0]
    n_total = int(cfg.get('GENERATE_NUM', 10000))
    dst_dir = join(cfg['BASE_DIR'], 'dataset', 'train_generated')
    pipe = GeneratePipeline(G, D, int(cfg['GAN_BATCH_SIZE_PER_GPU']))
    workers = args.encode_workers or max(1, (os.cpu_count() or 8) // world)
    pb = None
    if rank == 0:
        try:
            from tqdm import tqdm
            pb = tqdm(total=n_total)
        except Exception:
            pb = None
    t0 = time.time()
    n = generate_dataset(pipe, dst_dir, n_total, seed=0, psi=args.psi, rank=rank, world=world, workers=workers,
                         progress=(lambda k: pb.update(k * world)) if pb is not None else None, write=not args.no_write)
    torch.cuda.synchronize(dev)
    dt = time.time() - t0
    if pb is not None:
        pb.close()
    return n, dt


def main(argv=None):
    args = parse_args(argv)
    seed = 0
    np.random.seed(seed)                                    # main.py:28-30
    cfg = load_config_file(args.config)
    gan = cfg['GAN']
    if gan not in MAX_RES_LOG2:
        raise NotImplementedError(gan)
    if args.action == 'annotation':
        print('the annotation GUI (seg_annotator.py) is not part of this package; use the reference GUI with this '
              "package's ImageGenerator / SegSolver as drop-ins (INTEGRATION.md)")
        return 2
    rank, world, local = _dist()
    if args.action == 'train':
        solver = _solver(cfg, gan, [local] if world > 1 else list(cfg['SOLVER_GPU_IDS']), False)
        solver.fit()
    elif args.action == 'evaluate':
        solver = _solver(cfg, gan, [local] if world > 1 else list(cfg['SOLVER_GPU_IDS']), args.random_init)
        if not solver.is_trained:
            print('train Decoder first!')                   # main.py:67-69
            return -1
        if rank == 0:
            result = solver.evaluate(join(cfg['BASE_DIR'], 'eval'))
            print(', '.join([f'{name}: {value:.4f}' for name, value in result]))       # main.py:71-73
    elif args.action == 'generate':
        n, dt = run_generate(cfg, args, rank, world, local)
        import torch
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([float(n), dt], dtype=torch.float64, device='cuda')
            tot = t.clone()
            dist.all_reduce(tot[:1], op=dist.ReduceOp.SUM)
            dist.all_reduce(t[1:], op=dist.ReduceOp.MAX)
            n, dt = int(tot[0].item()), float(t[1].item())
        if rank == 0:
            print(f'generated {n} image+mask pairs in {dt:.2f} s ({n / dt:.1f} samples/s, {world} GPU(s), files '
                  f"{'skipped' if args.no_write else 'written'})")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())

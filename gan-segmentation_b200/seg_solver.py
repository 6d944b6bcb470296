"""Drop-in mirror of the predict / evaluate half of the reference's SegSolver (seg_solver.py:16-132, 222-349)."""
from __future__ import annotations

import os
from os.path import join

import numpy as np
import torch

from .config import decoder_config
from .networks import Decoder
from .random_init import init_decoder_params


def list_params_files(base_dir):
    """utils.list_files_with_ext(dir, ['.params']) (utils.py:18-37): sorted os.walk order."""
    out = []
    if not os.path.isdir(base_dir):
        return out
    for root, _, fnames in sorted(os.walk(base_dir)):
        rel = os.path.relpath(root, base_dir)
        for fn in fnames:
            if os.path.splitext(fn.lower())[1] == '.params' and os.path.isfile(join(root, fn)):
                out.append(fn if rel == '.' else join(rel, fn))
    return out


class SegSolver:
    """``SegSolver(max_res_log2, path_to_data, checkpoints_dir, gpu_ids, keep_weights)`` (seg_solver.py:17).

    The generate path needs ``predict`` / ``load`` / ``save`` and the attributes callers read
    (``is_trained``, ``net``, ``cfg``, ``ctx``, ``params_file``); ``evaluate`` runs the decoder and the loss kernel
    over a directory of annotated samples; ``fit`` trains the decoder (first, functional version on the
    single-operator kernels, see ``decoder_training.py``).
    """

    def __init__(self, max_res_log2, path_to_data, checkpoints_dir, gpu_ids, keep_weights=True, *, base_hw=(4, 4),
                 verbose=True):
        self.path_to_data = path_to_data
        self.checkpoints_dir = checkpoints_dir
        self.keep_weights = keep_weights
        if len(gpu_ids) == 0:
            raise RuntimeError('SegSolver needs at least one GPU id: the B200 path has no CPU fallback')
        self.ctx = [torch.device('cuda', i) for i in gpu_ids]
        self.is_trained = False
        self.params_file = None
        self.base_hw = base_hw
        self.verbose = verbose
        self.cfg = self.get_config(max_res_log2=max_res_log2)
        self.net = self.init_net()
        self.is_trained = self.load()

    def get_config(self, max_res_log2=9):
        return decoder_config(max_res_log2)

    def init_net(self):
        """Decoder + Xavier('in', 2.34) init, fresh BatchNorm statistics (seg_solver.py:36-49)."""
        params = init_decoder_params(self.cfg, seed=self.cfg['seed'], mode='reference')
        self.nets = [Decoder(self.cfg, num_devices=len(self.ctx), base_hw=self.base_hw, device=d) for d in self.ctx]
        for net in self.nets:
            net.set_parameters(params)
        if self.verbose:
            self.print_params(params, 'decoder')
        return self.nets[0]

    def print_params(self, params, title):
        """Same table as seg_solver.py:60-81."""
        row = '{:<36}{:<16}{:<24}{:<16}'
        print(row.format(title, 'params', 'weight shape', 'dtype'))
        print(row.format('---', '---', '---', '---'))
        total = 0
        for name, p in params.items():
            n = int(np.prod(p.shape))
            total += n
            print(row.format(name, n, str(tuple(p.shape)), 'np.float32'))
        print(row.format('---', '---', '---', '---'))
        print('{:<36}{:<16}'.format('total', total))
        print('{:<36}{:<16}'.format('---', '---'))

    def set_parameters(self, params):
        for net in self.nets:
            net.set_parameters(params)

    def predict(self, features=None, *, generator=None, device_outputs=False):
        """seg_solver.py:307-329: list of [C,H,W] / [N,C,H,W] features -> float32 [N,H,W,1] class ids.
        ``generator=`` takes the features the given Generator's last forward left in HBM instead."""
        if features is None:
            out = self.net.forward(None, generator=generator, return_logits=False)
            mask = out['mask']
        else:
            feats = []
            for f in features:
                t = torch.as_tensor(f, dtype=torch.float32)
                if t.dim() == 3:
                    t = t.unsqueeze(0)
                feats.append(t)
            n = feats[0].shape[0]
            k = len(self.nets)
            if n % k and k > 1:
                raise ValueError('batch must divide the number of devices (split_and_load even_split, :317)')
            per = n // k if k > 1 else n
            masks = []
            for i, net in enumerate(self.nets):
                sl = slice(i * per, (i + 1) * per)
                masks.append(net.forward([f[sl] for f in feats], return_logits=False)['mask'])
            for d in self.ctx:
                torch.cuda.synchronize(d)
            mask = torch.cat([m.to(self.ctx[0]) for m in masks], 0)
        if device_outputs:
            return mask
        return mask.cpu().numpy().astype(np.float32)[:, :, :, None]

    def save(self, suffix=None):
        param_name = 'checkpoint_last.params' if suffix is None else f'checkpoint_{suffix}.params'
        self.params_file = param_name
        os.makedirs(self.checkpoints_dir, exist_ok=True)
        self.net.save_parameters(join(self.checkpoints_dir, param_name))

    def load(self):
        files = list_params_files(self.checkpoints_dir)
        if len(files) > 0:
            params_file = files[0]
            print(f'loading checkpoint: {params_file}')
            self.params_file = params_file
            from .params_io import load_params
            self.set_parameters(load_params(join(self.checkpoints_dir, params_file)))
            return True
        return False

    def fit(self, epoch_end_callback=None, *, max_iters=None):
        """seg_solver.py:351-465: train the decoder on the annotated samples of ``path_to_data`` (generator frozen).
        Train-mode forward (BatchNorm batch statistics, Dropout(0.5) after every cvt block), SoftmaxCE with weight
        ``mask > -1``, backward, one all-reduce of the flat gradient bucket when ``torch.distributed`` is initialised
        (one process per GPU; the reference's in-process KVStore('nccl') sum), Adam(base_lr) with
        ``rescale_grad = 1/global batch``; ``epoch_end_callback()`` after every epoch; saves ``checkpoint_last.params``
        and returns ``[]`` like the reference.  Every iteration is one ``gsx_train_step`` call (``decoder_training.ResidentTrainer``: resident blocked
        16-bit tensors, tcgen05 convs / data / weight gradients, CUDA-graph replay) + the gradient all-reduce + ``gsx_adam_step``.  ``max_iters`` bounds the run (tests)."""
        import logging
        import time
        from .decoder_training import ResidentTrainer
        from .seg_datasets import CollectionDataset
        cfg = self.cfg
        if not self.keep_weights:
            self.net = self.init_net()
        ds = CollectionDataset(self.path_to_data, cfg, max_samples=None, load_to_memory=False)
        if len(ds) <= 0:
            print('number of training samples should be > 0')
            raise SystemExit(-1)
        bs = int(cfg['train_batch_size'])
        # One process per GPU: an iteration consumes bs samples PER RANK (rank r takes the r-th slice of bs*world
        # consecutive entries of the shuffled order, the same order on every rank), gradients are summed over the ranks
        # and rescaled by 1/(bs*world).  The reference splits one batch over its contexts (seg_solver.py:389-390), which
        # cannot work with its own train_batch_size = 1 on more than one GPU.
        rank, world = 0, 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
        iters_per_epoch = int(len(ds) / (bs * world))
        if iters_per_epoch <= 0:
            raise ValueError(f'{len(ds)} training samples are fewer than one global batch ({bs} x {world} ranks)')
        if rank == 0:
            print('total train samples: {}'.format(len(ds)))
            print('batch size: {}'.format(bs))
            print('epoch size: {}'.format(iters_per_epoch))
        with torch.cuda.device(self.ctx[0]):
            trainer = ResidentTrainer(cfg, self.net.get_parameters(), bs, device=self.ctx[0], base_hw=self.base_hw)
            rng = np.random.RandomState(cfg['seed'])
            # Device-side sample cache: the reference re-reads a 127 MiB feat_*.pickle per sample per step; the
            # 20 annotated samples of the FFHQ recipe are 2.5 GiB as fp32 and stay in HBM after their first use, up to
            # cfg['cache_max_size'] GB (seg_solver.py:88) -- beyond that samples come from disk as in the reference.
            cache, cache_bytes, cache_cap = {}, [0], float(cfg.get('cache_max_size', 4) or 0) * 2 ** 30

            def sample(j):
                if j in cache:
                    return cache[j]
                it = ds[j]
                fl = [np.asarray(f, np.float32) for f in it[2:]]
                nbytes = sum(f.nbytes for f in fl)
                if cache_bytes[0] + nbytes <= cache_cap:
                    entry = (np.asarray(it[1]), [torch.from_numpy(f).to(self.ctx[0]) for f in fl])
                    cache[j] = entry
                    cache_bytes[0] += nbytes
                    return entry
                return (np.asarray(it[1]), fl)

            display = cfg['train_display_iters']
            done = 0
            for epoch in range(int(cfg['train_epochs'])):
                tic = speed_tic = time.time()
                order = rng.permutation(len(ds))                       # DataLoader(shuffle=True, last_batch='discard')
                correct = total = 0                                    # mx.metric.Accuracy between two log lines (:173-175)
                ep_correct = ep_total = 0
                for nbatch in range(1, iters_per_epoch + 1):
                    lo = ((nbatch - 1) * world + rank) * bs
                    items = [sample(int(j)) for j in order[lo:lo + bs]]
                    mask = np.stack([it[0] for it in items]).astype(np.int32)
                    nfeat = len(items[0][1])
                    feats = [torch.stack([it[1][k] for it in items]) if torch.is_tensor(items[0][1][k]) else
                             np.stack([it[1][k] for it in items]) for k in range(nfeat)]
                    # Dropout masks: Philox bits of (seed, level, sample, element) inside the kernels; one seed per (step, rank)
                    drop_seed = (int(cfg['seed']) << 40) + (done * world + rank)
                    loss = trainer.step(feats, mask, dropout_seed=drop_seed, global_batch=bs * world)
                    pred = trainer.pred                                 # argmax of this step's logits (uint8, device)
                    if pred is not None:
                        hit = int((pred.cpu().numpy().astype(np.int32) == mask.reshape(pred.shape)).sum())
                        correct += hit; total += mask.size; ep_correct += hit; ep_total += mask.size
                    done += 1
                    if display is not None and nbatch % display == 0 and rank == 0:
                        speed = 1.0 * display * bs * world / (time.time() - speed_tic)
                        logging.info('Epoch[%03d] Batch[%04d] Speed: % 9.2f samples/sec accuracy=%f total-loss=%f', epoch, nbatch,
                                     speed, correct / max(total, 1), float(loss.mean().item()))
                        correct = total = 0
                        speed_tic = time.time()
                    if max_iters is not None and done >= max_iters:
                        break
                if rank == 0:
                    logging.info('Epoch[%d] Train-accuracy=%f', epoch + 1, ep_correct / max(ep_total, 1))
                    logging.info('Epoch[%d] Learning rate=%.5f', epoch + 1, trainer.lr)
                    logging.info('Epoch[%d] Time cost=%.3f', epoch + 1, time.time() - tic)
                if epoch_end_callback is not None and rank == 0:
                    epoch_end_callback()
                if max_iters is not None and done >= max_iters:
                    break
            self.set_parameters(trainer.state())
        self.is_trained = True
        if rank == 0:                                     # every rank holds the same weights; one writer
            self.save()
        if world > 1:
            torch.distributed.barrier()
        self._release_scratch()
        return []

    def evaluate(self, input_dir, output_dir=None):
        """seg_solver.py:222-305: pixel accuracy, mean IoU (background skipped) and the mean SoftmaxCE loss over
        the annotated samples of ``input_dir`` (``feat_*.pickle`` + ``img_*.jpg`` + ``mask_*.png``).  Returns
        ``[('accuracy', a), ('mean-iou', m), ('total-loss', l)]``; with ``output_dir`` also writes, per sample,
        the image, predicted / ground-truth masks (255 / 128 / 0) and a metrics line, as the reference does.
        Batches of ``cfg['val_batch_size']``, an incomplete last batch is discarded (``last_batch='discard'``, :166);
        the reference shuffles the order, which none of the returned numbers depends on."""
        from .metrics import SegmentationMetric
        from .seg_datasets import CollectionDataset
        from .training import softmax_ce
        ds = CollectionDataset(input_dir, self.cfg, max_samples=None, load_to_memory=False, output_idx=True)
        if len(ds) <= 0:
            print('number of training samples should be > 0')
            raise ValueError
        bs = int(self.cfg.get('val_batch_size', 1))
        print('total eval samples: {}'.format(len(ds)))
        print('batch size: {}'.format(bs))
        metric = SegmentationMetric(self.cfg['num_classes'], skip_bg=True)
        total_loss, total_cnt = 0.0, 0
        if output_dir is not None:
            os.makedirs(output_dir, exist_ok=True)
        for b0 in range(0, len(ds) - bs + 1, bs):
            items = [ds[i] for i in range(b0, b0 + bs)]
            idx = [int(it[0]) for it in items]
            imgs = np.stack([it[1] for it in items])                       # [N,3,H,W] float32 RGB
            mask = np.stack([it[2] for it in items]).astype(np.int32)      # [N,1,H,W] in {1,0,-1}
            nfeat = len(items[0]) - 3
            feats = [np.stack([np.asarray(it[3 + k], np.float32) for it in items]) for k in range(nfeat)]
            out = self.net.forward(feats, return_logits=True)
            logits = out['logits']
            labels = torch.as_tensor(mask, device=logits.device)
            loss, _ = softmax_ce(logits, labels, want_grad=False)           # weight (mask > -1), :239-242
            total_loss += float(loss.mean().item())
            total_cnt += 1
            pred_ids = out['mask'].cpu().numpy().astype(np.int64)           # first-max argmax of the same logits
            metric.update(mask[:, 0], pred_ids)
            if output_dir is not None:
                import cv2
                for i in range(bs):
                    m_i = SegmentationMetric(self.cfg['num_classes'], skip_bg=True)
                    m_i.update(mask[i:i + 1, 0], pred_ids[i:i + 1])
                    metric_str = ', '.join(f'{name} {v:.3f}' for name, v in m_i.get_name_value())
                    imname = ds.get_imname(idx[i])
                    img_i = np.transpose(imgs[i], (1, 2, 0))
                    pm = pred_ids[i].astype(np.int32)
                    gm = mask[i, 0].astype(np.int32).copy()
                    pm_out = np.where(pm == 1, 255, np.where(pm == 0, 128, pm)).astype(np.int32)
                    gm_out = np.where(gm == 1, 255, np.where(gm == 0, 128, np.where(gm == -1, 0, gm))).astype(np.int32)
                    cv2.imwrite(join(output_dir, imname), img_i[:, :, ::-1])
                    cv2.imwrite(join(output_dir, imname.replace('img', 'mask').replace('.jpg', '.png')), pm_out)
                    cv2.imwrite(join(output_dir, imname.replace('img', 'gt_mask').replace('.jpg', '.png')), gm_out)
                    with open(join(output_dir, imname.replace('img', 'metrics').replace('.jpg', '.txt')), 'w') as fp:
                        fp.write(', '.join(str(w) for w in [imname, img_i.shape, pm_out.shape, gm_out.shape, metric_str]) + '\n')
        result = metric.get_name_value()
        result.append(('total-loss', total_loss / total_cnt if total_cnt > 0 else 0.0))
        self._release_scratch()
        return result

    def _release_scratch(self):
        """The single-operator hooks (loss, metrics helpers) keep their scratch blocks in a pool outside torch's caching
        allocator; hand them back when a pass over the data ends so that torch can use the memory."""
        from . import _lib as L
        torch.cuda.synchronize(self.net.device)
        L.lib(self.net.dtype).gsx_op_release_cache()

"""TEST INFRASTRUCTURE: a plain-PyTorch implementation of the backend interface of
gan_segmentation_b200.decoder_training.DecoderTrainer, with every backward formula written out by hand (no autograd).
Used on the CPU to check the trainer's orchestration and formulas against the autograd oracle; never imported by the
product."""
import numpy as np
import torch
import torch.nn.functional as F


class _Adam:
    def __init__(self, shapes, P, lr, wd):
        self.names, self.P, self.lr, self.wd, self.t = list(shapes), P, lr, wd, 0
        self.m = {k: torch.zeros(s, dtype=torch.float64) for k, s in shapes.items()}
        self.v = {k: torch.zeros(s, dtype=torch.float64) for k, s in shapes.items()}

    def apply(self, grads, batch, group=None, grad_scale=1.0):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            grads = dict(grads)
            for k in self.names:                      # (the product all-reduces ONE flat bucket, training.FlatAdam)
                g = grads[k].clone()
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
                grads[k] = g
        self.t += 1
        lr_t = self.lr * np.sqrt(1 - 0.999 ** self.t) / (1 - 0.9 ** self.t)
        for k in self.names:
            w = self.P[k].double()
            g = grads[k].double() / (batch * grad_scale) + self.wd * w
            self.m[k] = 0.9 * self.m[k] + 0.1 * g
            self.v[k] = 0.999 * self.v[k] + 0.001 * g * g
            self.P[k] = (w - lr_t * self.m[k] / (self.v[k].sqrt() + 1e-8)).float()

    def export(self, P):
        return dict(P)


class TorchBackend:
    def tensor(self, a):
        return a.float() if torch.is_tensor(a) else torch.tensor(np.asarray(a, np.float32))

    def numpy(self, t):
        return t.detach().cpu().numpy()

    def cat(self, ts, dim):
        return torch.cat(ts, dim)

    def make_adam(self, shapes, P, lr, wd):
        return _Adam(shapes, P, lr, wd)

    def conv(self, xs, w, b, k):
        return F.conv2d(torch.cat(xs, 1), w, b, 1, k // 2)

    def upconv(self, xs, w, b):
        return F.conv2d(F.interpolate(torch.cat(xs, 1), scale_factor=2, mode='nearest'), w, b, 1, 1)

    def conv_dgrad(self, dy, w, k):
        wd = w.transpose(0, 1).flip(2, 3).contiguous()
        return F.conv2d(dy, wd, None, 1, k // 2)

    def conv_wgrad(self, x, dy, k):
        p = k // 2
        xp = F.pad(x, (p, p, p, p))
        n, cin, h, w = x.shape
        dw = torch.zeros((dy.shape[1], cin, k, k))
        for ky in range(k):
            for kx in range(k):
                dw[:, :, ky, kx] = torch.einsum('nohw,nihw->oi', dy, xp[:, :, ky:ky + h, kx:kx + w])
        return dw, dy.sum(dim=(0, 2, 3))

    def upsample2(self, x):
        return F.interpolate(x, scale_factor=2, mode='nearest')

    def sumpool2(self, dy):
        return F.avg_pool2d(dy, 2) * 4.0

    def bn_lrelu_fwd(self, z, gamma, beta, drop):
        mean = z.mean(dim=(0, 2, 3))
        var = z.var(dim=(0, 2, 3), unbiased=False)
        rstd = torch.rsqrt(var + 1e-5)
        xh = (z - mean[None, :, None, None]) * rstd[None, :, None, None]
        y = F.leaky_relu(xh * gamma[None, :, None, None] + beta[None, :, None, None], 0.2)
        if drop is not None:
            y = y * drop * 2.0
        return y, dict(mean=mean, var=var, rstd=rstd)

    def bn_lrelu_bwd(self, dy, z, c, gamma, beta, drop):
        xh = (z - c['mean'][None, :, None, None]) * c['rstd'][None, :, None, None]
        pre = xh * gamma[None, :, None, None] + beta[None, :, None, None]
        g = dy * (drop * 2.0 if drop is not None else 1.0) * torch.where(pre > 0, torch.ones_like(pre), torch.full_like(pre, 0.2))
        m = z.shape[0] * z.shape[2] * z.shape[3]
        dbeta = g.sum(dim=(0, 2, 3))
        dgamma = (g * xh).sum(dim=(0, 2, 3))
        dz = (gamma * c['rstd'] / m)[None, :, None, None] * (m * g - dbeta[None, :, None, None] - xh * dgamma[None, :, None, None])
        return dz, dgamma, dbeta

    def lrelu_fwd(self, z, drop):
        y = F.leaky_relu(z, 0.2)
        return (y * drop * 2.0 if drop is not None else y), {}

    def lrelu_bwd(self, dy, z, drop):
        return dy * (drop * 2.0 if drop is not None else 1.0) * torch.where(z > 0, torch.ones_like(z), torch.full_like(z, 0.2))

    def softmax_ce(self, logits, mask):
        mask = torch.as_tensor(np.asarray(mask)).long()
        w = (mask > -1).float()
        lp = F.log_softmax(logits, dim=1)
        loss = (-torch.gather(lp, 1, mask.clamp(min=0)) * w).mean(dim=(1, 2, 3))
        onehot = torch.zeros_like(logits).scatter_(1, mask.clamp(min=0), 1.0)
        hw = logits.shape[2] * logits.shape[3]
        return loss, w * (lp.exp() - onehot) / hw, 1.0

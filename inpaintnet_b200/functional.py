"""torch.autograd bindings of the engine: the reference's forward signatures stay intact and
`loss.backward()` (utils/trainer.py:150) drives the hand-written backward kernels."""
import os

import torch

from . import engine, ops
from .ops import F32

_DEFAULT_PRECISION = os.environ.get("INPAINTNET_B200_PRECISION", "bf16")


def default_precision():
    return _DEFAULT_PRECISION


def set_default_precision(name):
    global _DEFAULT_PRECISION
    assert name in ("fp32", "bf16")
    _DEFAULT_PRECISION = name


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, tokens, arena, pre, prec, cfg, training, need_grad):
        mu, ls, saved = engine.encoder_forward(arena, pre, prec, cfg, tokens, training, need_grad)
        ctx.state = (arena, pre, prec, cfg, saved)
        return mu, ls

    @staticmethod
    def backward(ctx, dmu, dls):
        arena, pre, prec, cfg, saved = ctx.state
        if saved is None:
            raise RuntimeError("encoder backward called twice or forward ran without grad")
        ctx.state = (arena, pre, prec, cfg, None)
        engine.encoder_backward(arena, pre, prec, cfg, saved, dmu, dls)
        return (None,) * 8


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, z, tokens, arena, pre, prec, cfg, teacher_forced, training, need_grad):
        weights, samples, saved = engine.decoder_forward(arena, pre, prec, cfg, z, tokens, teacher_forced, training,
                                                         need_grad)
        ctx.state = (arena, pre, prec, cfg, saved)
        ctx.need_dz = z.requires_grad
        ctx.mark_non_differentiable(samples)
        return weights, samples

    @staticmethod
    def backward(ctx, dweights, _dsamples):
        arena, pre, prec, cfg, saved = ctx.state
        if saved is None:
            raise RuntimeError("decoder backward called twice or forward ran without grad")
        ctx.state = (arena, pre, prec, cfg, None)
        dz = engine.decoder_backward(arena, pre, prec, cfg, saved, dweights, need_dz=ctx.need_dz)
        return (None, dz) + (None,) * 8


class _ReparamFn(torch.autograd.Function):
    """z = mu + exp(log_std) * eps   (MeasureVAE/measure_vae.py:119: z_dist.rsample())"""

    @staticmethod
    def forward(ctx, mu, log_std, eps):
        mu, log_std, eps = mu.contiguous(), log_std.contiguous(), eps.contiguous()
        z = torch.empty_like(mu)
        ops.reparam_fwd(mu.data_ptr(), log_std.data_ptr(), eps.data_ptr(), mu.numel(), z.data_ptr(), 0, F32)
        ctx.save_for_backward(log_std, eps)
        return z

    @staticmethod
    def backward(ctx, dz):
        log_std, eps = ctx.saved_tensors
        dz = dz.contiguous()
        dmu = torch.zeros_like(dz)
        dls = torch.zeros_like(dz)
        ops.reparam_bwd(dz.data_ptr(), F32, log_std.data_ptr(), eps.data_ptr(), dz.numel(), dmu.data_ptr(),
                        dls.data_ptr(), F32)
        return dmu, dls, None


class _CeKlFn(torch.autograd.Function):
    """loss = mean CE(weights, targets) [+ beta * mean_b sum_j KL], accuracy; forward and backward in ONE
    pass over the logits (utils/trainer.py:271-306, MeasureVAE/vae_trainer.py:128-139)."""

    @staticmethod
    def forward(ctx, weights, targets, mu, log_std, beta):
        ops.require_cuda(weights, "weights")
        w = weights.contiguous()
        t = targets.contiguous()
        V = w.shape[-1]
        rows = w.numel() // V
        assert t.numel() == rows
        dev = w.device
        scalars = torch.zeros(4, dtype=torch.float32, device=dev)
        want = weights.requires_grad or (mu is not None and mu.requires_grad)
        dl = torch.empty_like(w) if want else None
        has_kl = mu is not None
        dmu = dls = None
        if has_kl:
            mu_c, ls_c = mu.contiguous(), log_std.contiguous()
            Bz, Z = mu_c.shape
            if want:
                dmu, dls = torch.empty_like(mu_c), torch.empty_like(ls_c)
        ops.ce_kl(w.data_ptr(), t.data_ptr(), rows, V, scalars.data_ptr(), dlogits=dl.data_ptr() if want else 0, dl_dt=F32,
                  ld_dl=V, mu=mu_c.data_ptr() if has_kl else 0, log_std=ls_c.data_ptr() if has_kl else 0,
                  Bz=Bz if has_kl else 0, Z=Z if has_kl else 0, beta=beta, dmu=dmu.data_ptr() if dmu is not None else 0,
                  dls=dls.data_ptr() if dls is not None else 0, dz_dt=F32)
        ce = scalars[0] / rows
        loss = ce + (beta / Bz) * scalars[1] if has_kl else ce
        acc = scalars[2] / rows
        ctx.grads = (dl, dmu, dls)
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    def backward(ctx, dloss, _dacc):
        dl, dmu, dls = ctx.grads
        ctx.grads = None
        return (dl * dloss if dl is not None else None, None, dmu * dloss if dmu is not None else None,
                dls * dloss if dls is not None else None, None)


def fused_ce_kl(weights, targets, mu=None, log_std=None, beta=0.001):
    """Returns (loss, accuracy) as 0-dim device tensors."""
    return _CeKlFn.apply(weights, targets, mu, log_std, beta)

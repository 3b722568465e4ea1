"""LatentRNN on the B200 hot path -- reference: LatentRNN/latent_rnn.py:11-307 (same constructor,
forward signature, attribute names and state_dict layout: 102 keys = own GRUs/linear + vae_model.*)."""
import os
import random

import torch
from torch import nn

from . import engine_latent, functional as Fn
from .arena import arena_of
from .measure_vae import MeasureVAE, _anchor
from .model_base import Model
from .ops import Precision


class _LatentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, past, future, arena, prec, model, n_gen, need_grad, target, teacher_forcing):
        weights, samples, z_out, saved = engine_latent.latent_forward(arena, prec, model, past, future, n_gen,
                                                                      model.training, need_grad, target=target,
                                                                      teacher_forcing=teacher_forcing)
        ctx.state = (arena, prec, saved)
        ctx.mark_non_differentiable(samples)
        return weights, samples, z_out

    @staticmethod
    def backward(ctx, dweights, _ds, dz_out):
        arena, prec, saved = ctx.state
        if saved is None:
            raise RuntimeError("LatentRNN backward called twice or forward ran without grad")
        ctx.state = (arena, prec, None)
        engine_latent.latent_backward(arena, prec, saved, dweights, dz_out)
        return (None,) * 10


class LatentRNN(Model):
    # LatentRNNAblations sets this to "past" / "future": the generation GRU is then seeded by that context GRU only
    type = None

    def __init__(self, dataset, vae_model: MeasureVAE, num_rnn_layers, rnn_hidden_size, dropout, rnn_class,
                 auto_reg=False, teacher_forcing=True):
        super(LatentRNN, self).__init__()
        self.dataset = dataset.__repr__()
        self.vae_model = vae_model
        self.auto_reg = auto_reg
        self.use_teacher_forcing = teacher_forcing if self.auto_reg else False
        self.teacher_forcing_prob = 0.5
        for param in self.vae_model.parameters():
            param.requires_grad = False
        print('Freeze the ', self.vae_model.__repr__(), ' model.')
        self.num_rnn_layers = num_rnn_layers
        self.rnn_hidden_size = rnn_hidden_size
        self.dropout = dropout
        self.z_dim = self.vae_model.latent_space_dim
        self.rnn_class = rnn_class
        self.bidirectional = True
        self.rnn_num_direction = 2 if self.bidirectional else 1
        self.context_rnn_past = self.rnn_class(input_size=self.z_dim, hidden_size=self.rnn_hidden_size,
                                               num_layers=self.num_rnn_layers, dropout=self.dropout,
                                               bidirectional=self.bidirectional, batch_first=True)
        self.context_rnn_future = self.rnn_class(input_size=self.z_dim, hidden_size=self.rnn_hidden_size,
                                                 num_layers=self.num_rnn_layers, dropout=self.dropout,
                                                 bidirectional=self.bidirectional, batch_first=True)
        if self.auto_reg:
            self.gen_rnn_input_dim = self.z_dim
        else:
            self.gen_rnn_input_dim = 1
            self.x_0 = nn.Parameter(data=torch.randn(1, 1, self.gen_rnn_input_dim))
        gen_hidden = self.rnn_hidden_size * (self.num_rnn_layers if self.type is None else 1)
        self.generation_rnn = self.rnn_class(input_size=self.gen_rnn_input_dim, hidden_size=gen_hidden,
                                             num_layers=self.num_rnn_layers, dropout=self.dropout,
                                             bidirectional=self.bidirectional, batch_first=True)
        self.generation_linear = nn.Linear(gen_hidden * self.rnn_num_direction, self.z_dim)
        self.xavier_initialization()
        cur_dir = os.path.dirname(os.path.realpath(__file__))
        self.filepath = os.path.join(cur_dir, 'models/', self.__repr__())
        self.precision = None
        self.vae_model._set_root(self)   # the VAE's parameters now live in this model's arena

    def set_precision(self, name):
        assert name in ("fp32", "bf16")
        self.precision = name
        self.vae_model.set_precision(name)
        return self

    def __repr__(self):
        filestr = f'LatentRNN(' \
                  f'{self.type if self.type is not None else ""}' \
                  f'{self.dataset}' \
                  f'{self.rnn_class},' \
                  f'{self.num_rnn_layers},' \
                  f'{self.rnn_hidden_size},' \
                  f'{self.dropout},' \
                  f')'
        if self.auto_reg:
            filestr += 'auto_reg'
        if self.use_teacher_forcing:
            filestr += ',tf'
        else:
            filestr += ',no_tf'
        return filestr

    def forward(self, past_context, future_context, target, measures_to_generate, train=True):
        """past_context (batch, n_past, 24), future_context (batch, n_future, 24), target (batch, n_target, 24)
        -> weights (batch, measures_to_generate, 24, num_notes), samples (batch, 1, 24*measures_to_generate),
        gen_z (batch, measures_to_generate, z_dim)            -- latent_rnn.py:110-159.
        The target-encode of latent_rnn.py:133 only feeds the teacher-forced autoregressive model (its first
        measures_to_generate - 1 measures); it is skipped everywhere else."""
        if self.use_teacher_forcing and train:                       # latent_rnn.py:142-145
            teacher_forcing = random.random() < self.teacher_forcing_prob
        else:
            teacher_forcing = False
        arena = arena_of(self)
        prec = Precision(self.precision or Fn.default_precision())
        anchor = _anchor(self)
        need_grad = torch.is_grad_enabled() and anchor is not None
        past = past_context if past_context.dtype == torch.int64 else past_context.long()
        fut = future_context if future_context.dtype == torch.int64 else future_context.long()
        tgt = None
        if self.auto_reg and teacher_forcing:
            tgt = target if target.dtype == torch.int64 else target.long()
        weights, samples, gen_z = _LatentFn.apply(anchor, past, fut, arena, prec, self, int(measures_to_generate), need_grad,
                                                  tgt, teacher_forcing)
        return weights, samples, gen_z

    def xavier_initialization(self):
        for mod in (self.context_rnn_past, self.context_rnn_future, self.generation_rnn, self.generation_linear):
            for name, param in mod.named_parameters():
                if 'weight' in name:
                    nn.init.xavier_normal_(param)


class LatentRNNAblations(LatentRNN):
    """reference: LatentRNN/latent_rnn_ablations.py:11-313 -- the generation GRU (hidden size = the context GRUs',
    :79) is seeded by the past OR the future context only (:143-146); everything else as LatentRNN.  Both
    context GRUs are still parameters (same state_dict keys); the unused one receives a zero gradient."""

    def __init__(self, dataset, vae_model: MeasureVAE, num_rnn_layers, rnn_hidden_size, dropout, rnn_class,
                 auto_reg=False, teacher_forcing=True, type='past'):
        if type not in ("past", "future"):
            raise ValueError(type)
        self.__dict__["type"] = type      # read by LatentRNN.__init__ (sizes) before nn.Module state exists
        super(LatentRNNAblations, self).__init__(dataset, vae_model, num_rnn_layers, rnn_hidden_size, dropout, rnn_class,
                                                 auto_reg=auto_reg, teacher_forcing=teacher_forcing)
        self.type = type
        # __repr__ (and so the checkpoint file name) is the reference's: 'LatentRNN(' + type + ... (:97-110)

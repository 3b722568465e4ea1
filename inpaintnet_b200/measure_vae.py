"""MeasureVAE on the B200 hot path -- same constructor / forward signatures, attribute names and
state_dict layout as the reference (MeasureVAE/measure_vae.py, encoder.py, decoder.py), but
`forward` runs the hand-written sm_100a kernels through the C-ABI instead of torch.nn.GRU/Linear.

The nn.GRU / nn.Linear / nn.Embedding sub-modules are kept as PARAMETER CONTAINERS only (they
give the reference's state_dict keys and initialisation); their own forward is never called.
"""
import os
import random

import torch
from torch import nn, distributions
from torch.nn import Parameter

from . import engine, functional as Fn, ops
from .arena import arena_of, prefix_of
from .ops import Precision
from .model_base import Model


def _anchor(module):
    """A parameter that requires grad (so autograd records our Function), or None when frozen."""
    for p in module.parameters():
        if p.requires_grad:
            return p
    return None


class Encoder(nn.Module):
    """reference: MeasureVAE/encoder.py:9-134"""

    def __init__(self, note_embedding_dim, rnn_hidden_size, num_layers, num_notes, dropout, bidirectional, z_dim,
                 rnn_class):
        super().__init__()
        self.bidirectional = bidirectional
        self.num_directions = 2 if bidirectional else 1
        self.note_embedding_dim = note_embedding_dim
        self.num_layers = num_layers
        self.rnn_hidden_size = rnn_hidden_size
        self.z_dim = z_dim
        self.dropout = dropout
        self.rnn_class = rnn_class
        self.lstm = self.rnn_class(input_size=note_embedding_dim, hidden_size=rnn_hidden_size, num_layers=num_layers,
                                   dropout=self.dropout, bidirectional=self.bidirectional, batch_first=True)
        self.num_notes = num_notes
        self.note_embedding_layer = nn.Embedding(self.num_notes, self.note_embedding_dim)
        width = self.rnn_hidden_size * self.num_directions
        self.linear_mean = nn.Sequential(nn.Linear(width * self.num_layers, width), nn.SELU(), nn.Linear(width, z_dim))
        self.linear_log_std = nn.Sequential(nn.Linear(width * self.num_layers, width), nn.SELU(), nn.Linear(width, z_dim))
        self.xavier_initialization()
        self.precision = None  # None -> functional.default_precision()
        if num_layers != 2 or not bidirectional:
            raise NotImplementedError("the fused B200 encoder implements the reference configuration "
                                      "(2 layers, bidirectional); use inpaintnet_b200.GRU for other stacks")

    def __repr__(self):
        return f'Encoder(' \
               f'{self.note_embedding_dim},' \
               f'{self.rnn_class},' \
               f'{self.num_layers},' \
               f'{self.rnn_hidden_size},' \
               f'{self.dropout},' \
               f'{self.bidirectional},' \
               f'{self.z_dim},' \
               f')'

    def xavier_initialization(self):
        for name, param in self.named_parameters():
            if 'weight' in name:
                nn.init.xavier_normal_(param)

    def _cfg(self):
        return engine.EncCfg(self.num_notes, self.note_embedding_dim, self.rnn_hidden_size, self.z_dim, self.dropout)

    def forward_params(self, score_tensor):
        """Returns (mu, log_std) -- the fused path used by MeasureVAE / LatentRNN."""
        arena = arena_of(getattr(self, "_ipn_root", self))
        pre = prefix_of(self, arena)
        prec = Precision(self.precision or Fn.default_precision())
        anchor = _anchor(self)
        need_grad = torch.is_grad_enabled() and anchor is not None
        batch_size, measure_seq_len = score_tensor.size()
        mu, log_std = Fn._EncoderFn.apply(anchor, score_tensor, arena, pre, prec, self._cfg(), self.training, need_grad)
        return mu, log_std

    def forward(self, score_tensor):
        """score_tensor (batch_size, measure_seq_len) int64 -> torch Normal distribution (encoder.py:104-134).
        The NaN guard of encoder.py:111-116 is folded into the fused Adam kernel (device flag polled by the
        trainer once per step) instead of 13 host syncs per call."""
        z_mean, z_log_std = self.forward_params(score_tensor)
        z_distribution = distributions.Normal(loc=z_mean, scale=torch.exp(z_log_std), validate_args=False)
        z_distribution.log_std = z_log_std  # kept for the fused KL term
        return z_distribution


class Decoder(nn.Module):
    """reference: MeasureVAE/decoder.py:11-54 (abstract base; only HierarchicalDecoder is ever built)"""

    def __init__(self, note_embedding_dim, num_notes, z_dim):
        super().__init__()
        self.name = 'DecoderABC'
        self.num_notes = num_notes
        self.note_embedding_dim = note_embedding_dim
        self.z_dim = z_dim
        self.note_embedding_layer = nn.Embedding(self.num_notes, self.note_embedding_dim)

    def xavier_initialization(self):
        for name, param in self.named_parameters():
            if 'weight' in name:
                nn.init.xavier_normal_(param)


class HierarchicalDecoder(Decoder):
    """reference: MeasureVAE/decoder.py:313-529"""

    def __init__(self, note_embedding_dim, num_notes, z_dim, num_layers, rnn_hidden_size, dropout, rnn_class):
        super().__init__(note_embedding_dim, num_notes, z_dim)
        self.name = 'HierarchicalDecoder'
        self.rnn_class = rnn_class
        self.num_layers = num_layers
        self.rnn_hidden_size = rnn_hidden_size
        self.dropout = dropout
        self.z_to_beat_rnn_input = nn.Sequential(nn.Linear(self.z_dim, self.rnn_hidden_size * self.num_layers), nn.SELU())
        self.beat_rnn_input_dim = 1
        self.b_0 = Parameter(data=torch.zeros(self.beat_rnn_input_dim))
        self.rnn_beat = self.rnn_class(input_size=self.beat_rnn_input_dim, hidden_size=self.rnn_hidden_size,
                                       num_layers=self.num_layers, dropout=self.dropout, batch_first=True)
        self.beat_emb_to_tick_rnn_hidden = nn.Sequential(
            nn.Linear(self.rnn_hidden_size, self.rnn_hidden_size * self.num_layers), nn.SELU())
        self.beat_emb_to_tick_rnn_input = nn.Sequential(nn.Linear(self.rnn_hidden_size, self.rnn_hidden_size), nn.SELU())
        self.x_0 = Parameter(data=torch.zeros(note_embedding_dim))
        self.rnn_tick = self.rnn_class(input_size=self.note_embedding_dim + self.rnn_hidden_size,
                                       hidden_size=self.rnn_hidden_size, num_layers=self.num_layers,
                                       dropout=self.dropout, batch_first=True)
        self.tick_emb_to_note_emb = nn.Sequential(nn.Linear(self.rnn_hidden_size, self.num_notes), nn.ReLU())
        self.use_teacher_forcing = True
        self.teacher_forcing_prob = 0.5
        self.sampling = 'argmax'
        self.xavier_initialization()
        self.precision = None
        if num_layers != 2:
            raise NotImplementedError("the fused B200 decoder implements the reference configuration (2 layers)")

    def __repr__(self):
        return f'{self.name}' \
               f'{self.note_embedding_dim},' \
               f'{self.rnn_class},' \
               f'{self.num_layers},' \
               f'{self.rnn_hidden_size},' \
               f'{self.dropout},' \
               f')'

    def _cfg(self):
        return engine.EncCfg(self.num_notes, self.note_embedding_dim, self.rnn_hidden_size, self.z_dim, self.dropout)

    def forward(self, z, score_tensor, train):
        """z (batch, z_dim), score_tensor (batch, 24) -> weights (batch, 24, num_notes) post-ReLU,
        samples (batch, 1, 24).  decoder.py:412-453: ONE python coin per batch selects teacher forcing."""
        if self.use_teacher_forcing and train:
            teacher_forced = random.random() < self.teacher_forcing_prob
        else:
            teacher_forced = False
        sampling = 'argmax' if not train else self.sampling
        if sampling != 'argmax':
            raise NotImplementedError  # decoder.py:518 (multinomial is never selected: sampling='argmax', :376)
        batch_size_z, z_dim = z.size()
        assert (z_dim == self.z_dim)
        batch_size, measure_seq_len = score_tensor.size()
        assert (batch_size == batch_size_z)
        assert (measure_seq_len == 24)
        arena = arena_of(getattr(self, "_ipn_root", self))
        pre = prefix_of(self, arena)
        prec = Precision(self.precision or Fn.default_precision())
        anchor = _anchor(self)
        need_grad = torch.is_grad_enabled() and (anchor is not None or z.requires_grad)
        if teacher_forced and score_tensor.dtype != torch.int64:
            score_tensor = score_tensor.long()
        weights, samples = Fn._DecoderFn.apply(anchor, z, score_tensor, arena, pre, prec, self._cfg(), teacher_forced,
                                               self.training, need_grad)
        return weights, samples


class MeasureVAE(Model):
    """reference: MeasureVAE/measure_vae.py:10-169"""

    def __init__(self, dataset, note_embedding_dim=10, metadata_embedding_dim=2, num_encoder_layers=2,
                 encoder_hidden_size=512, encoder_dropout_prob=0.5, latent_space_dim=256, num_decoder_layers=2,
                 decoder_hidden_size=512, decoder_dropout_prob=0.5, has_metadata=False):
        super().__init__()
        self.num_beats_per_measure = 4
        self.num_ticks_per_measure = 24
        self.num_ticks_per_beat = int(self.num_ticks_per_measure / self.num_beats_per_measure)
        self.dataset = dataset.__repr__()
        self.note_embedding_dim = note_embedding_dim
        self.metadata_embedding_dim = metadata_embedding_dim
        self.num_encoder_layers = num_encoder_layers
        self.encoder_hidden_size = encoder_hidden_size
        self.encoder_dropout_prob = encoder_dropout_prob
        self.latent_space_dim = latent_space_dim
        self.num_decoder_layers = num_decoder_layers
        self.decoder_hidden_size = decoder_hidden_size
        self.decoder_dropout_prob = decoder_dropout_prob
        self.has_metadata = has_metadata
        self.num_notes = len(dataset.note2index_dicts[0])
        print("NUMBER OF NOTES: ", self.num_notes)
        self.encoder = Encoder(note_embedding_dim=self.note_embedding_dim, rnn_hidden_size=self.encoder_hidden_size,
                               num_layers=self.num_encoder_layers, num_notes=self.num_notes,
                               dropout=self.encoder_dropout_prob, bidirectional=True, z_dim=self.latent_space_dim,
                               rnn_class=torch.nn.GRU)
        self.decoder = HierarchicalDecoder(note_embedding_dim=self.note_embedding_dim, num_notes=self.num_notes,
                                           z_dim=self.latent_space_dim, num_layers=self.num_decoder_layers,
                                           rnn_hidden_size=self.decoder_hidden_size, dropout=self.decoder_dropout_prob,
                                           rnn_class=torch.nn.GRU)
        cur_dir = os.path.dirname(os.path.realpath(__file__))
        self.filepath = os.path.join(cur_dir, 'models/', self.__repr__())
        self._set_root(self)

    def _set_root(self, root):
        # sub-modules look their parameters up in the arena of the outermost model
        object.__setattr__(self.encoder, "_ipn_root", root)
        object.__setattr__(self.decoder, "_ipn_root", root)

    def set_precision(self, name):
        """'bf16' (tcgen05 tensor cores, default) or 'fp32' (exact-parity CUDA-core mode)."""
        assert name in ("fp32", "bf16")
        self.encoder.precision = name
        self.decoder.precision = name
        return self

    def __repr__(self):
        return f'MeasureVAE(' \
               f'{self.dataset},' \
               f'{self.encoder.__repr__()},' \
               f'{self.decoder.__repr__()},' \
               f')'

    def forward(self, measure_score_tensor, train=True):
        """measure_score_tensor (batch, 24) int64 ->
        (weights, samples, z_dist, prior_dist, z_tilde, z_prior)   -- measure_vae.py:97-134"""
        seq_len = measure_score_tensor.size(1)
        assert (seq_len == self.num_ticks_per_measure)
        z_dist = self.encoder(measure_score_tensor)
        arena = arena_of(self.encoder._ipn_root)
        eps = engine.NOISE.normal(arena, tuple(z_dist.loc.shape), z_dist.loc.device)
        z_tilde = Fn._ReparamFn.apply(z_dist.loc, z_dist.log_std, eps)
        prior_dist = distributions.Normal(loc=torch.zeros_like(z_dist.loc), scale=torch.ones_like(z_dist.scale),
                                          validate_args=False)
        z_prior = engine.NOISE.normal(arena, tuple(z_dist.loc.shape), z_dist.loc.device) if engine.NOISE.eps is None \
            else torch.zeros_like(z_dist.loc)
        weights, samples = self.decoder(z=z_tilde, score_tensor=measure_score_tensor, train=train)
        return weights, samples, z_dist, prior_dist, z_tilde, z_prior

    def forward_test(self, measure_score_tensor):
        """(batch, num_measures, 24) -> weights (batch, num_measures, 24, V), samples (batch, 1, 24*num_measures);
        measure_vae.py:136-169.  The per-measure python loops of the reference are batched into one
        (batch*num_measures) encode and one decode: measures are independent."""
        batch_size, num_measures, seq_len = measure_score_tensor.size()
        assert (seq_len == self.num_ticks_per_measure)
        flat = measure_score_tensor.reshape(batch_size * num_measures, seq_len)
        z_dist = self.encoder(flat)
        arena = arena_of(self.encoder._ipn_root)
        eps = engine.NOISE.normal(arena, tuple(z_dist.loc.shape), z_dist.loc.device)
        z_tilde = Fn._ReparamFn.apply(z_dist.loc, z_dist.log_std, eps)
        w, s = self.decoder(z=z_tilde, score_tensor=flat, train=False)
        weights = w.view(batch_size, num_measures, seq_len, -1)
        samples = s.view(batch_size, 1, num_measures * seq_len)
        return weights, samples

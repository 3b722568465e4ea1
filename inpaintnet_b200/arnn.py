"""AnticipationRNN baseline on the B200 hot path -- reference:
AnticipationRNN/anticipation_rnn_gauss_reg_model.py:42-435 (ConstraintModelGaussianReg), :682-725
(AnticipationRNNBaseline) and AnticipationRNN/anticipation_rnn_trainer.py.

Constraint stack (LSTMs over the time-flipped [metadata | masked-note] embeddings) and generation stack
(LSTMs over [shifted note embedding | constraint output]) run on the fused LSTM step kernels; every
input projection, the two head linears and all weight gradients are hoisted GEMMs.  Supports the
single-voice folk dataset the reference scripts use (num_voices == 1).  Same constructor / forward
signature / state_dict keys (`linear_ouput_notes` spelled as in the reference).
"""
import os
import random

import torch
from torch import nn
from torch.nn import ModuleList, Embedding

from . import ops, functional as Fn
from .arena import arena_of
from .engine import _lin, _dgrad, _wgrad, NOISE
from .helpers import to_cuda_variable_long
from .measure_vae import _anchor
from .model_base import Model
from .ops import F32, ACT_NONE, ACT_RELU, CORE_SIMT, MUL_RELU_GRAD, MUL_KEEP_MASK, Precision
from .trainer import Trainer, LatentRNNTrainer


def _round8(x):
    return (x + 7) // 8 * 8


def _lstm_stack_fwd(arena, prec, names, P0, T, B, H, need_grad, last_y_reverse=False, step=None, table=0, tokptr=0,
                    state=None, persist=False, gates_blocked=False):
    """Runs a stack of single-layer LSTMs. names[l] = parameter prefix ('lstm_generation.0.').
    persist=False: P0 = tensor [T*B,4H], input projection of layer 0 incl. b_ih; one fused kernel per timestep.
        step=None: whole sequence; step=t: only timestep t (serial decode); step=(s0, s1): timesteps [s0, s1).
        gates_blocked: the saved gates are written in the persistent kernels' layout (the backward pass then runs
        the persistent cluster kernel although the forward ran tick by tick).
    persist=True (whole sequence only): P0 = BLOCKED input projection of layer 0 (ops.lstm_inproj_blocked: b_ih + b_hh
        folded); every layer runs as ONE persistent cluster kernel (W_hh resident in shared memory, h_t exchanged
        through distributed shared memory) and the projections between the layers are written blocked by the GEMM.
    Returns the state dict (hseq, cseq, gates, y per layer)."""
    dev, act = P0.device, prec.tdt
    L = len(names)
    if persist:
        assert step is None and state is None and not table
        gcols = ops.lstm_gates_cols(H, True)
        hseq, cseq, gates, ys, Ps = [], [], [], [], [P0]
        for l in range(L):
            h = torch.empty((T + 1) * B, H, dtype=act, device=dev)
            h[:B].zero_()
            c = torch.empty((T + 1) * B, H, dtype=torch.float32, device=dev)   # only slots 0 and T are touched
            c[:B].zero_()
            hseq.append(h)
            cseq.append(c)
            gates.append(torch.empty(T * B, gcols, dtype=act, device=dev) if need_grad else None)
            rev = last_y_reverse and l == L - 1
            # y_t == h_t: the layer output is the hseq slots 1..T unless it has to be written time-flipped
            ys.append(torch.empty(T * B, H, dtype=act, device=dev) if rev else h[B:])
            if l > 0:
                nm = names[l]
                w = arena.w(prec, nm + "weight_ih_l0")
                Pl = torch.empty(T * B, 4 * H, dtype=act, device=dev)
                ops.lstm_inproj_blocked(ys[l - 1].data_ptr(), H, H, w[0], w[1], T * B, arena.fptr(nm + "bias_ih_l0"),
                                        arena.fptr(nm + "bias_hh_l0"), H, Pl.data_ptr())
                Ps.append(Pl)
            nm = names[l]
            ops.lstm_layer_fwd(prec, T, B, H, arena.w(prec, nm + "weight_hh_l0")[0], arena.fptr(nm + "bias_hh_l0"),
                               Ps[l].data_ptr(), 4 * H, h.data_ptr(), c.data_ptr(),
                               gates=gates[l].data_ptr() if need_grad else 0, y=ys[l].data_ptr() if rev else 0, ld_y=H,
                               y_reverse_time=1 if rev else 0, P_blocked=1)
        return dict(hseq=hseq, cseq=cseq, gates=gates, y=ys, P=Ps, persist=True)
    if state is None:
        gcols = ops.lstm_gates_cols(H, True) if gates_blocked else 4 * H
        state = dict(hseq=[torch.zeros((T + 1) * B, H, dtype=act, device=dev) for _ in range(L)],
                     cseq=[torch.zeros((T + 1) * B, H, dtype=torch.float32, device=dev) for _ in range(L)],
                     gates=[torch.empty(T * B, gcols, dtype=act, device=dev) if need_grad else None for _ in range(L)],
                     y=[torch.empty(T * B, H, dtype=act, device=dev) for _ in range(L)],
                     P=[P0] + [torch.empty(T * B, 4 * H, dtype=act, device=dev) for _ in range(L - 1)],
                     persist=bool(gates_blocked and need_grad), gates_blocked=bool(gates_blocked and need_grad))
    s0, s1 = (0, 0) if step is None else (step if isinstance(step, tuple) else (step, step + 1))
    for l, nm in enumerate(names):
        if l > 0:
            rows, r0 = (T * B, 0) if step is None else ((s1 - s0) * B, s0 * B)
            _lin(prec, state["y"][l - 1].data_ptr() + prec.es * r0 * H, H, rows, H, arena.w(prec, nm + "weight_ih_l0"), 4 * H,
                 state["P"][l].data_ptr() + prec.es * r0 * 4 * H, prec.act, 4 * H, bias=arena.fptr(nm + "bias_ih_l0"))
        ops.lstm_layer_fwd(prec, T, B, H, arena.w(prec, nm + "weight_hh_l0")[0], arena.fptr(nm + "bias_hh_l0"),
                           state["P"][l].data_ptr(), 4 * H, state["hseq"][l].data_ptr(), state["cseq"][l].data_ptr(),
                           gates=state["gates"][l].data_ptr() if need_grad else 0, y=state["y"][l].data_ptr(), ld_y=H,
                           y_reverse_time=1 if (last_y_reverse and l == L - 1) else 0, s_begin=s0, s_end=s1,
                           table=table if l == 0 else 0, ld_table=4 * H, tok_scalar=tokptr if l == 0 else 0,
                           gates_blocked=1 if state.get("gates_blocked") else 0)
    return state


def _lstm_stack_bwd(arena, prec, names, state, dY_last, T, B, H, X0_parts, last_y_reverse=False):
    """Backward of _lstm_stack_fwd.  dY_last: tensor [T*B,H] gradient wrt the last layer's output (in the order it
    was written).  X0_parts: list of (X ptr, ld, K, w_ih column offset) making up layer 0's input.
    Returns dP of layer 0 [T*B,4H] (for the caller's input gradients)."""
    dev, act, es = dY_last.device, prec.tdt, prec.es
    L = len(names)
    persist = bool(state.get("persist"))
    ws = None if persist else torch.empty(3 * B * H, dtype=torch.float32, device=dev)
    dY = dY_last
    dP = None
    for l in range(L - 1, -1, -1):
        nm = names[l]
        dP = torch.empty(T * B, 4 * H, dtype=act, device=dev)
        rev = 1 if (last_y_reverse and l == L - 1) else 0
        if persist:
            ops.lstm_layer_bwd(prec, T, B, H, arena.w(prec, nm + "weight_hh_l0")[0], 0, 0, state["gates"][l].data_ptr(),
                               dY.data_ptr(), H, 0, dP.data_ptr(), 0, y_reverse_time=rev, gates_persist=1)
        else:
            ops.lstm_layer_bwd(prec, T, B, H, arena.w(prec, nm + "weight_hh_l0")[0], state["hseq"][l].data_ptr(),
                               state["cseq"][l].data_ptr(), state["gates"][l].data_ptr(), dY.data_ptr(), H, 0, dP.data_ptr(),
                               ws.data_ptr(), y_reverse_time=rev)
        if arena.wants_grad(nm + "weight_hh_l0"):
            _wgrad(prec, dP.data_ptr(), 4 * H, 4 * H, state["hseq"][l].data_ptr(), H, H, T * B, arena.gptr(nm + "weight_hh_l0"), H)
            ops.colsum(dP.data_ptr(), prec.act, 4 * H, T * B, 4 * H, arena.gptr(nm + "bias_ih_l0"),
                       out2=arena.gptr(nm + "bias_hh_l0"), cols2=4 * H)
            p = arena.by_name[nm + "weight_ih_l0"]
            full = p.shape[1]
            parts = X0_parts if l == 0 else [(state["y"][l - 1].data_ptr(), H, H, 0)]
            for (xp, ldx, K, c0) in parts:
                _wgrad(prec, dP.data_ptr(), 4 * H, 4 * H, xp, ldx, K, T * B, arena.gptr(nm + "weight_ih_l0", c0), full)
        if l > 0:
            dYn = torch.empty(T * B, H, dtype=act, device=dev)
            _dgrad(prec, dP.data_ptr(), 4 * H, T * B, 4 * H, arena.w(prec, nm + "weight_ih_l0"), H, dYn.data_ptr(), prec.act, H)
            dY = dYn
    return dP


class _ArnnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, model, arena, prec, score, metadata, constraints_loc, teacher_forcing, need_grad):
        logits, saved = model._engine_forward(arena, prec, score, metadata, constraints_loc, teacher_forcing, need_grad)
        ctx.state = (model, arena, prec, saved)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model, arena, prec, saved = ctx.state
        if saved is None:
            raise RuntimeError("ARNN backward called twice or forward ran without grad")
        ctx.state = (model, arena, prec, None)
        model._engine_backward(arena, prec, saved, dlogits)
        return (None,) * 9


class ConstraintModelGaussianReg(Model):
    def __init__(self, dataset, note_embedding_dim=20, metadata_embedding_dim=30, num_lstm_constraints_units=256,
                 num_lstm_generation_units=256, linear_hidden_size=128, num_layers=1, dropout_input_prob=0.2,
                 dropout_prob=0.5, unary_constraint=False, teacher_forcing=True):
        super(ConstraintModelGaussianReg, self).__init__()
        self.dataset = dataset
        self.use_teacher_forcing = teacher_forcing
        self.teacher_forcing_prob = 0.5
        self.num_layers = num_layers
        self.num_units_linear = linear_hidden_size
        self.unary_constraint = unary_constraint
        unary_constraint_size = 1 if self.unary_constraint else 0
        self.note_embedding_dim = note_embedding_dim
        self.num_lstm_generation_units = num_lstm_generation_units
        self.num_notes_per_voice = [len(d) for d in self.dataset.note2index_dicts]
        self.note_embeddings = ModuleList([Embedding(n + unary_constraint_size, self.note_embedding_dim)
                                           for n in self.num_notes_per_voice])
        self.metadata_embedding_dim = metadata_embedding_dim
        self.num_elements_per_metadata = [m.num_values for m in self.dataset.metadatas]
        self.num_elements_per_metadata.append(self.dataset.num_voices)
        self.metadata_embeddings = ModuleList([Embedding(n, self.metadata_embedding_dim)
                                               for n in self.num_elements_per_metadata])
        self.num_lstm_constraints_units = num_lstm_constraints_units
        self.dropout_input_prob = dropout_input_prob
        self.dropout_prob = dropout_prob
        Hc = self.num_lstm_constraints_units
        c_in = self.metadata_embedding_dim * len(self.num_elements_per_metadata) + self.note_embedding_dim * unary_constraint_size
        self.lstm_constraint = nn.ModuleList([nn.LSTM(input_size=i, hidden_size=h, num_layers=1, dropout=dropout_prob, batch_first=True)
                                              for i, h in [(c_in, Hc)] + [(Hc, Hc)] * (self.num_layers - 1)])
        self.lstm_generation = nn.ModuleList([nn.LSTM(input_size=i, hidden_size=h, num_layers=1, dropout=dropout_prob, batch_first=True)
                                              for i, h in [(self.note_embedding_dim + Hc, Hc)] + [(Hc, Hc)] * (self.num_layers - 1)])
        self.linear_1 = nn.Linear(self.num_lstm_generation_units, linear_hidden_size)
        self.linear_ouput_notes = ModuleList([nn.Linear(self.num_units_linear, n) for n in self.num_notes_per_voice])
        self.optimizer = None  # the reference builds an unused torch Adam here (arnn_model.py:145)
        cur_dir = os.path.dirname(os.path.realpath(__file__))
        self.filepath = os.path.join(cur_dir, 'models/', self.__repr__())
        self.precision = None
        if self.dataset.num_voices != 1 or not unary_constraint or num_lstm_constraints_units != num_lstm_generation_units:
            raise NotImplementedError("the B200 AnticipationRNN path implements the reference training configuration: "
                                      "one voice, unary constraints, equal constraint/generation widths")

    def set_precision(self, name):
        assert name in ("fp32", "bf16")
        self.precision = name
        return self

    def __repr__(self):
        filestr = f'AnticipationRNNReg(' \
                  f'{self.dataset.__repr__()},' \
                  f'{self.note_embedding_dim},' \
                  f'{self.metadata_embedding_dim},' \
                  f'{self.num_lstm_constraints_units},' \
                  f'{self.num_lstm_generation_units},' \
                  f'{self.num_units_linear},' \
                  f'{self.num_layers},' \
                  f'{self.dropout_input_prob},' \
                  f'{self.dropout_prob},' \
                  f'{self.unary_constraint},' \
                  f')'
        filestr += ',tf' if self.use_teacher_forcing else ',no_tf'
        return filestr

    # ------------------------------------------------------------------------------------------
    def forward(self, score_tensor, metadata_tensor, constraints_loc, start_tick=None, end_tick=None, train=True):
        """score (batch, 1, length), metadata (batch, 1, length, 3), constraints_loc (batch, 1, length) ->
        ([weights (batch, n_gap_ticks, num_notes)], None)        -- arnn_model.py:406-435"""
        if self.use_teacher_forcing and train:
            teacher_forcing = random.random() <= self.teacher_forcing_prob
        else:
            teacher_forcing = False
        arena = arena_of(self)
        prec = Precision(self.precision or Fn.default_precision())
        anchor = _anchor(self)
        need_grad = torch.is_grad_enabled() and anchor is not None
        logits = _ArnnFn.apply(anchor, self, arena, prec, score_tensor.long(), metadata_tensor.long(),
                               constraints_loc.long(), teacher_forcing, need_grad)
        gap = (constraints_loc[0, 0, :] == 0).nonzero().squeeze()
        return [logits[:, gap, :]], None

    def forward_inpaint(self, score_tensor, metadata_tensor, constraints_loc, start_tick, end_tick):
        """Inference: teacher-forced scan of the ticks before the gap, then one tick at a time through the gap
        [start_tick, end_tick) feeding back the argmax of batch element 0 (arnn_model.py:261-346).
        -> ([weights (batch, end_tick - start_tick, num_notes)], gen_score (batch, 1, length))"""
        arena = arena_of(self)
        prec = Precision(self.precision or Fn.default_precision())
        start_tick, end_tick = int(start_tick), int(end_tick)
        if not 0 <= start_tick < end_tick <= score_tensor.shape[2]:
            raise ValueError(f"forward_inpaint: bad gap [{start_tick}, {end_tick}) for length {score_tensor.shape[2]}")
        with torch.no_grad():
            logits, gen = self._engine_forward(arena, prec, score_tensor.long(), metadata_tensor.long(),
                                               constraints_loc.long(), False, False, inpaint=(start_tick, end_tick))
        return [logits], gen

    # ------------------------------------------------------------------------------------------
    def _engine_forward(self, arena, prec, score, metadata, cl, teacher_forcing, need_grad, inpaint=None):
        ops.require_cuda(score, "score tensor")
        B, _, T = score.shape
        H, E, Em, L = self.num_lstm_constraints_units, self.note_embedding_dim, self.metadata_embedding_dim, self.num_layers
        V, Lh = self.num_notes_per_voice[0], self.num_units_linear
        nmeta = len(self.num_elements_per_metadata)
        dev, act, es = score.device, prec.tdt, prec.es
        arena.refresh()
        # ---- index tensors, time-major [T, B] (tiny int tensors: plain torch)
        tok = score[:, 0].t().contiguous().to(torch.int32)
        masked = torch.where(cl[:, 0].t() > 0, tok, torch.full_like(tok, V)).contiguous()      # arnn_model.py:510-532
        md = metadata[:, 0].permute(1, 0, 2).contiguous().to(torch.int32)                       # [T, B, nmeta]
        # ---- constraint stack on the time-flipped sequence (arnn_model.py:455-475)
        Ic = Em * nmeta + E
        ldc = _round8(Ic)
        Xc = torch.zeros(T * B, ldc, dtype=act, device=dev)
        md_f, masked_f = md.flip(0).contiguous(), masked.flip(0).contiguous()
        for k in range(nmeta):
            ops.gather_cols(arena.fptr(f"metadata_embeddings.{k}.weight"), Em, md_f.data_ptr() + 4 * k, nmeta, T * B,
                            Xc.data_ptr(), prec.act, ldc, k * Em)
        ops.gather_cols(arena.fptr("note_embeddings.0.weight"), E, masked_f.data_ptr(), 1, T * B, Xc.data_ptr(), prec.act, ldc,
                        nmeta * Em)
        cnames = [f"lstm_constraint.{l}." for l in range(L)]
        gnames = [f"lstm_generation.{l}." for l in range(L)]
        Pc = torch.empty(T * B, 4 * H, dtype=act, device=dev)
        # whole-sequence stacks run the persistent cluster kernels when the shape allows (bf16 tensor-core mode,
        # H = 128 / 256, B % 128 == 0); forward_inpaint mixes a prefix scan with per-tick steps and stays per-step
        persist = inpaint is None and ops.lstm_persist_eligible(prec, B, H)
        wc0 = arena.w(prec, cnames[0] + "weight_ih_l0")
        if persist:
            ops.lstm_inproj_blocked(Xc.data_ptr(), ldc, Ic, wc0[0], wc0[1], T * B, arena.fptr(cnames[0] + "bias_ih_l0"),
                                    arena.fptr(cnames[0] + "bias_hh_l0"), H, Pc.data_ptr())
        else:
            _lin(prec, Xc.data_ptr(), ldc, T * B, Ic, wc0, 4 * H, Pc.data_ptr(), prec.act, 4 * H,
                 bias=arena.fptr(cnames[0] + "bias_ih_l0"))
        cstate = _lstm_stack_fwd(arena, prec, cnames, Pc, T, B, H, need_grad, last_y_reverse=True, persist=persist)
        cout = cstate["y"][L - 1]                                   # [T*B, H] in NORMAL time order
        # ---- generation stack
        w_e = arena.w(prec, gnames[0] + "weight_ih_l0", 0, E)
        w_c = arena.w(prec, gnames[0] + "weight_ih_l0", E, H)
        Pg = torch.empty(T * B, 4 * H, dtype=act, device=dev)
        saved_in = {}
        if inpaint is not None:
            return self._inpaint_generation(arena, prec, score, tok, cout, gnames, w_e, w_c, Pg, inpaint)
        if teacher_forcing:
            # shift-right note embeddings, zero first step, whole-timestep dropout (arnn_model.py:367-373,437-442)
            idx = torch.cat((torch.full((1, B), -1, dtype=torch.int32, device=dev), tok[:-1]), 0).contiguous()
            scale = None
            if self.training and self.dropout_input_prob > 0:
                keep = NOISE.keep_mask(arena, (T * B,), self.dropout_input_prob, dev)
                scale = keep.to(torch.float32) / (1.0 - self.dropout_input_prob)
            Xe = torch.zeros(T * B, 16, dtype=act, device=dev)
            ops.gather_cols(arena.fptr("note_embeddings.0.weight"), E, idx.data_ptr(), 1, T * B, Xe.data_ptr(), prec.act, 16, 0,
                            row_scale=scale.data_ptr() if scale is not None else 0)
            if persist:
                ops.lstm_inproj_blocked(Xe.data_ptr(), 16, E, w_e[0], w_e[1], T * B, arena.fptr(gnames[0] + "bias_ih_l0"),
                                        arena.fptr(gnames[0] + "bias_hh_l0"), H, Pg.data_ptr(),
                                        X2=cout.data_ptr(), ldx2=H, K2=H, w_ih2=w_c[0], ldw2=w_c[1])
            else:
                ops.gemm(prec.core, prec.act, T * B, 4 * H,
                         [(Xe.data_ptr(), 16, 0, w_e[0], w_e[1], 0, E), (cout.data_ptr(), H, 0, w_c[0], w_c[1], 0, H)],
                         Pg.data_ptr(), prec.act, 4 * H, bias=arena.fptr(gnames[0] + "bias_ih_l0"))
            gstate = _lstm_stack_fwd(arena, prec, gnames, Pg, T, B, H, need_grad, persist=persist)
            hid = torch.empty(T * B, Lh, dtype=act, device=dev)
            _lin(prec, gstate["y"][L - 1].data_ptr(), H, T * B, H, arena.w(prec, "linear_1.weight"), Lh, hid.data_ptr(), prec.act,
                 Lh, bias=arena.fptr("linear_1.bias"), act=ACT_RELU)
            logits = torch.empty(B, T, V, dtype=torch.float32, device=dev)
            _lin(prec, hid.data_ptr(), Lh, T * B, Lh, arena.w(prec, "linear_ouput_notes.0.weight"), V, logits.data_ptr(), F32, V,
                 bias=arena.fptr("linear_ouput_notes.0.bias"), rowmap=(1 << 30, B, 0, V, T * V))
            saved_in = dict(idx=idx, scale=scale, Xe=Xe)
        else:
            # no teacher forcing (arnn_model.py:190-259): 384 serial ticks; the token fed back to the WHOLE batch is
            # the argmax of batch element 0 (reference behaviour, :252-256); start symbol id 0.
            _lin(prec, cout.data_ptr(), H, T * B, H, w_c, 4 * H, Pg.data_ptr(), prec.act, 4 * H,
                 bias=arena.fptr(gnames[0] + "bias_ih_l0"))

            table = self._gen_table(arena, gnames, dev)
            tokens_in = torch.zeros(T + 1, dtype=torch.int32, device=dev)     # tokens_in[t] = input token of tick t
            hid = torch.empty(T * B, Lh, dtype=act, device=dev)
            logits = torch.empty(B, T, V, dtype=torch.float32, device=dev)
            gstate = None
            w1, wo = arena.w(prec, "linear_1.weight"), arena.w(prec, "linear_ouput_notes.0.weight")
            for t in range(T):
                gstate = _lstm_stack_fwd(arena, prec, gnames, Pg, T, B, H, need_grad, step=t, table=table.data_ptr(),
                                         tokptr=tokens_in.data_ptr() + 4 * t, state=gstate, gates_blocked=persist)
                yo = gstate["y"][L - 1].data_ptr() + es * t * B * H
                _lin(prec, yo, H, B, H, w1, Lh, hid.data_ptr() + es * t * B * Lh, prec.act, Lh, bias=arena.fptr("linear_1.bias"),
                     act=ACT_RELU)
                _lin(prec, hid.data_ptr() + es * t * B * Lh, Lh, B, Lh, wo, V, logits.data_ptr() + 4 * t * V, F32, T * V,
                     bias=arena.fptr("linear_ouput_notes.0.bias"))
                ops.argmax_rows(logits.data_ptr() + 4 * t * V, 1, V, tok_out=tokens_in.data_ptr() + 4 * (t + 1))
            saved_in = dict(tokens_in=tokens_in)
        saved = None
        if need_grad:
            saved = dict(B=B, T=T, Xc=Xc, ldc=ldc, Ic=Ic, md_f=md_f, masked_f=masked_f, cstate=cstate, gstate=gstate, cout=cout,
                         hid=hid, tf=teacher_forcing, **saved_in)
        return logits, saved

    def _gen_table(self, arena, gnames, dev, zero_row=False):
        """table[v] = note_embedding[v] @ W_ih[:, :E]^T of the first generation LSTM (gathered by the fed-back token);
        zero_row: one extra all-zero row (id V+1 = 'no token term')."""
        H, E, V = self.num_lstm_constraints_units, self.note_embedding_dim, self.num_notes_per_voice[0]

        def build_table():
            tab = torch.zeros(V + 2 if zero_row else V + 1, 4 * H, dtype=torch.float32, device=dev)
            ops.gemm(CORE_SIMT, F32, V + 1, 4 * H, [(arena.fptr("note_embeddings.0.weight"), E, 0,
                                                    arena.fptr(gnames[0] + "weight_ih_l0"), E + H, 0, E)],
                     tab.data_ptr(), F32, 4 * H)
            return tab

        return arena.derived(("arnn", "gen_table_z" if zero_row else "gen_table"), build_table)

    def _inpaint_generation(self, arena, prec, score, tok, cout, gnames, w_e, w_c, Pg, inpaint):
        """arnn_model.py:287-346.  The hoisted input projection covers both regimes: rows < start carry the
        shifted ground-truth embedding (with the whole-timestep input dropout in train mode, :291), row `start`
        the undropped embedding of score[start-1] (per batch element, :318), rows > start only the constraint part
        -- their token term is the fed-back argmax of batch element 0, gathered from the table inside the step kernel."""
        B, _, T = score.shape
        H, E, L = self.num_lstm_constraints_units, self.note_embedding_dim, self.num_layers
        V, Lh = self.num_notes_per_voice[0], self.num_units_linear
        dev, act, es = score.device, prec.tdt, prec.es
        s0, s1 = inpaint
        idx = torch.cat((torch.full((1, B), -1, dtype=torch.int32, device=dev), tok[:-1]), 0)
        idx[s0 + 1:] = -1
        if s0 == 0:
            idx[0] = -1
        idx = idx.contiguous()
        scale = None
        if self.training and self.dropout_input_prob > 0:
            keep = NOISE.keep_mask(arena, (T * B,), self.dropout_input_prob, dev)
            scale = keep.to(torch.float32) / (1.0 - self.dropout_input_prob)
            scale[s0 * B:] = 1.0
        Xe = torch.zeros(T * B, 16, dtype=act, device=dev)
        ops.gather_cols(arena.fptr("note_embeddings.0.weight"), E, idx.data_ptr(), 1, T * B, Xe.data_ptr(), prec.act, 16, 0,
                        row_scale=scale.data_ptr() if scale is not None else 0)
        ops.gemm(prec.core, prec.act, T * B, 4 * H,
                 [(Xe.data_ptr(), 16, 0, w_e[0], w_e[1], 0, E), (cout.data_ptr(), H, 0, w_c[0], w_c[1], 0, H)],
                 Pg.data_ptr(), prec.act, 4 * H, bias=arena.fptr(gnames[0] + "bias_ih_l0"))
        gstate = None
        if s0 > 0:
            gstate = _lstm_stack_fwd(arena, prec, gnames, Pg, T, B, H, False, step=(0, s0))
        table = self._gen_table(arena, gnames, dev, zero_row=True)
        tokens_in = torch.zeros(T + 1, dtype=torch.int32, device=dev)     # tokens_in[t] = token fed at tick t
        if s0 > 0:
            tokens_in[s0] = V + 1        # tick `start`: the token term is already in Pg (per batch element)
        n = s1 - s0
        hid = torch.empty(n * B, Lh, dtype=act, device=dev)
        logits = torch.empty(B, n, V, dtype=torch.float32, device=dev)
        w1, wo = arena.w(prec, "linear_1.weight"), arena.w(prec, "linear_ouput_notes.0.weight")
        for t in range(s0, s1):
            gstate = _lstm_stack_fwd(arena, prec, gnames, Pg, T, B, H, False, step=t, table=table.data_ptr(),
                                     tokptr=tokens_in.data_ptr() + 4 * t, state=gstate)
            yo = gstate["y"][L - 1].data_ptr() + es * t * B * H
            k = t - s0
            _lin(prec, yo, H, B, H, w1, Lh, hid.data_ptr() + es * k * B * Lh, prec.act, Lh, bias=arena.fptr("linear_1.bias"),
                 act=ACT_RELU)
            _lin(prec, hid.data_ptr() + es * k * B * Lh, Lh, B, Lh, wo, V, logits.data_ptr() + 4 * k * V, F32, n * V,
                 bias=arena.fptr("linear_ouput_notes.0.bias"))
            ops.argmax_rows(logits.data_ptr() + 4 * k * V, 1, V, tok_out=tokens_in.data_ptr() + 4 * (t + 1))
        gen = score.clone()
        gen[:, 0, s0:s1] = tokens_in[s0 + 1:s1 + 1].to(torch.int64)[None, :]
        return logits, gen

    def _engine_backward(self, arena, prec, sv, dlogits):
        B, T = sv["B"], sv["T"]
        H, E, Em, L = self.num_lstm_constraints_units, self.note_embedding_dim, self.metadata_embedding_dim, self.num_layers
        V, Lh = self.num_notes_per_voice[0], self.num_units_linear
        nmeta = len(self.num_elements_per_metadata)
        dev, act, es = dlogits.device, prec.tdt, prec.es
        arena.refresh()
        cnames = [f"lstm_constraint.{l}." for l in range(L)]
        gnames = [f"lstm_generation.{l}." for l in range(L)]
        Vp = _round8(V)
        # (B,T,V) fp32 -> time-major [T*B, Vp] act  (tiny relative to the rest; torch permute + our convert)
        dl_tm = dlogits.transpose(0, 1).contiguous()
        dl = torch.zeros(T * B, Vp, dtype=act, device=dev)
        ops.convert_2d(dl_tm.data_ptr(), F32, V, dl.data_ptr(), prec.act, Vp, T * B, V)
        del dl_tm
        hid, gstate, cstate, cout = sv["hid"], sv["gstate"], sv["cstate"], sv["cout"]
        _wgrad(prec, dl.data_ptr(), Vp, V, hid.data_ptr(), Lh, Lh, T * B, arena.gptr("linear_ouput_notes.0.weight"), Lh)
        ops.colsum(dl.data_ptr(), prec.act, Vp, T * B, V, arena.gptr("linear_ouput_notes.0.bias"))
        dhid = torch.empty(T * B, Lh, dtype=act, device=dev)
        _dgrad(prec, dl.data_ptr(), Vp, T * B, V, arena.w(prec, "linear_ouput_notes.0.weight"), Lh, dhid.data_ptr(), prec.act, Lh,
               mul=(hid.data_ptr(), prec.act, Lh, MUL_RELU_GRAD, 1.0))
        ygen = gstate["y"][L - 1]
        _wgrad(prec, dhid.data_ptr(), Lh, Lh, ygen.data_ptr(), H, H, T * B, arena.gptr("linear_1.weight"), H)
        ops.colsum(dhid.data_ptr(), prec.act, Lh, T * B, Lh, arena.gptr("linear_1.bias"))
        dYg = torch.empty(T * B, H, dtype=act, device=dev)
        _dgrad(prec, dhid.data_ptr(), Lh, T * B, Lh, arena.w(prec, "linear_1.weight"), H, dYg.data_ptr(), prec.act, H)
        # ---- generation stack
        if sv["tf"]:
            Xe, idx, scale = sv["Xe"], sv["idx"], sv["scale"]
        else:
            idx = sv["tokens_in"][:T].view(T, 1).expand(T, B).contiguous()
            scale = None
            Xe = torch.zeros(T * B, 16, dtype=act, device=dev)
            ops.gather_cols(arena.fptr("note_embeddings.0.weight"), E, idx.data_ptr(), 1, T * B, Xe.data_ptr(), prec.act, 16, 0)
        dPg = _lstm_stack_bwd(arena, prec, gnames, gstate, dYg, T, B, H,
                              [(Xe.data_ptr(), 16, E, 0), (cout.data_ptr(), H, H, E)])
        dXe = torch.empty(T * B, 16, dtype=act, device=dev)
        mul = None
        if scale is not None:
            keep16 = (scale > 0).to(torch.uint8).view(-1, 1).expand(T * B, 16).contiguous()
            mul = (keep16.data_ptr(), 2, 16, MUL_KEEP_MASK, 1.0 / (1.0 - self.dropout_input_prob))
        _dgrad(prec, dPg.data_ptr(), 4 * H, T * B, 4 * H, arena.w(prec, gnames[0] + "weight_ih_l0", 0, E), E, dXe.data_ptr(),
               prec.act, 16, mul=mul)
        Vn = V + 1
        demb = arena.gptr("note_embeddings.0.weight")
        ops.embed_grad(dXe.data_ptr(), prec.act, 16, idx.data_ptr(), T * B, E, Vn, demb, skip_id=-1)
        dcout = torch.empty(T * B, H, dtype=act, device=dev)
        _dgrad(prec, dPg.data_ptr(), 4 * H, T * B, 4 * H, arena.w(prec, gnames[0] + "weight_ih_l0", E, H), H, dcout.data_ptr(),
               prec.act, H)
        # ---- constraint stack (its last layer wrote y time-flipped; dcout is in the same (normal) order)
        Xc, ldc, Ic = sv["Xc"], sv["ldc"], sv["Ic"]
        dPc = _lstm_stack_bwd(arena, prec, cnames, cstate, dcout, T, B, H, [(Xc.data_ptr(), ldc, Ic, 0)], last_y_reverse=True)
        dXc = torch.empty(T * B, ldc, dtype=act, device=dev)
        _dgrad(prec, dPc.data_ptr(), 4 * H, T * B, 4 * H, arena.w(prec, cnames[0] + "weight_ih_l0"), Ic, dXc.data_ptr(), prec.act, ldc)
        md_f, masked_f = sv["md_f"], sv["masked_f"]
        for k in range(nmeta):
            idx_k = md_f[:, :, k].contiguous()
            ops.embed_grad(dXc.data_ptr() + es * k * Em, prec.act, ldc, idx_k.data_ptr(), T * B, Em,
                           self.num_elements_per_metadata[k], arena.gptr(f"metadata_embeddings.{k}.weight"), skip_id=-1)
        ops.embed_grad(dXc.data_ptr() + es * nmeta * Em, prec.act, ldc, masked_f.data_ptr(), T * B, E, Vn, demb, skip_id=-1)


class AnticipationRNNBaseline(ConstraintModelGaussianReg):
    """reference: anticipation_rnn_gauss_reg_model.py:682-725 (differs only in __repr__)"""

    def __repr__(self):
        return super().__repr__().replace('AnticipationRNNReg(', 'AnticipationRNNBaseline(')


class AnticipationRNNGaussianRegTrainer(Trainer):
    """reference: AnticipationRNN/anticipation_rnn_trainer.py:11-182"""

    def __init__(self, dataset, model, lr=1e-4, early_stopping=False):
        super().__init__(dataset, model, lr, early_stopping)
        self.min_num_measures_target = 2
        self.max_num_measure_target = 6
        assert (self.dataset.n_bars > self.max_num_measure_target)
        self.measure_seq_len = self.dataset.subdivision * self.dataset.num_beats_per_bar

    def loss_and_acc_for_batch(self, batch, epoch_num=None, train=True):
        score_tensor, metadata_tensor, constraints_loc, start_tick, end_tick = batch
        weights, _ = self.model(score_tensor=score_tensor, metadata_tensor=metadata_tensor, constraints_loc=constraints_loc,
                                start_tick=start_tick, end_tick=end_tick, train=train)
        targets = score_tensor[:, :, (constraints_loc[0, 0, :] == 0).nonzero().squeeze()]
        targets = targets.transpose(0, 1)
        loss = self.mean_crossentropy_loss(weights=weights, targets=targets)
        accuracy = self.mean_accuracy(weights=weights, targets=targets)
        return loss, accuracy

    def process_batch_data(self, batch):
        score_tensor, metadata_tensor = batch
        constraint_loc, start_tick, end_tick = self.get_constraints_location(score_tensor)
        return (to_cuda_variable_long(score_tensor), to_cuda_variable_long(metadata_tensor),
                to_cuda_variable_long(constraint_loc), start_tick, end_tick)

    def get_constraints_location(self, score_tensor, extra_outs=False, fix_num_target=None):
        """anticipation_rnn_trainer.py:93-128: constraints everywhere except a gap of 2..6 measures."""
        measures_tensor = LatentRNNTrainer.split_to_measures(score_tensor, self.measure_seq_len)
        num_measures = measures_tensor.size(1)
        assert (num_measures == self.dataset.n_bars)
        if fix_num_target is None:
            num_target = int(torch.randint(low=self.min_num_measures_target, high=self.max_num_measure_target + 1, size=(1,)).item())
        else:
            num_target = fix_num_target
        num_past = int(torch.randint(low=1, high=num_measures - num_target - 1, size=(1,)).item())
        start_tick = (num_past + 1) * self.measure_seq_len
        end_tick = start_tick + num_target * self.measure_seq_len
        constraints_location = torch.zeros_like(score_tensor)
        if start_tick > 0:
            constraints_location[:, :, :start_tick] = 1
        if end_tick < constraints_location.size(2) - 1:
            constraints_location[:, :, end_tick:] = 1
        return constraints_location, start_tick, end_tick

    def update_scheduler(self, epoch_num):
        return

    @staticmethod
    def mean_crossentropy_loss(weights, targets):
        """list of (batch, seq, num_notes) per voice, targets (voice, batch, seq) -- trainer.py:166-182"""
        total = 0
        for i, w in enumerate(weights):
            total = total + Fn.fused_ce_kl(w, targets[i])[0]
        return total / len(weights)

    @staticmethod
    def mean_accuracy(weights, targets):
        total = 0
        with torch.no_grad():
            for i, w in enumerate(weights):
                total = total + Fn.fused_ce_kl(w.detach(), targets[i])[1]
        return total / len(weights)


class AnticipationRNNBaselineTrainer(AnticipationRNNGaussianRegTrainer):
    pass

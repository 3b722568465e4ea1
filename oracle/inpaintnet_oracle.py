"""CPU oracle for the InpaintNet hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain restatement (torch CPU tensors, explicit per-timestep loops, no nn.GRU / nn.LSTM /
cuDNN) of the algorithm the reference delegates to torch modules.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import
this file; the product path (inpaintnet_b200/*) never does and fails loudly without its CUDA
library.

Parity status: PINNED.  The reference has no golden vectors of its own, so
the oracle is pinned against outputs of the UNMODIFIED reference modules run in the build
container (tests/golden/make_golden.py -> tests/golden/*.pt, checked by
tests/test_oracle_golden.py) and, when /root/reference is mounted, against the live reference
(tests/test_oracle_vs_reference.py).

Every function cites the reference file:line it restates (paths relative to /root/reference).
State dict keys are the reference's own.

Randomness is always INJECTED: dropout keep-masks (0/1 tensors), eps for the
reparameterisation and the teacher-forcing coin are arguments, never drawn here.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

SELU_ALPHA = 1.6732632423543772848170429916717
SELU_SCALE = 1.0507009873554804934193349852946


def selu(x: Tensor) -> Tensor:
    # torch.nn.SELU (used at MeasureVAE/encoder.py:44,50; MeasureVAE/decoder.py:337,352,357)
    return SELU_SCALE * torch.where(x > 0, x, SELU_ALPHA * (torch.exp(x) - 1.0))


def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    y = x @ w.t()
    return y if b is None else y + b


# --------------------------------------------------------------------------------------
# GRU (torch.nn.GRU semantics; gate row order [r; z; n])
# reference call sites: MeasureVAE/encoder.py:28-35,125; MeasureVAE/decoder.py:342-348,
# 361-367,470,498; LatentRNN/latent_rnn.py:53-82,188-190,231,249
# --------------------------------------------------------------------------------------

def gru_cell(xp: Tensor, h: Tensor, w_hh: Tensor, b_hh: Tensor) -> Tensor:
    """One GRU step given the input projection xp = x W_ih^T + b_ih  (B, 3H)."""
    H = h.shape[1]
    g = h @ w_hh.t() + b_hh
    r = torch.sigmoid(xp[:, :H] + g[:, :H])
    z = torch.sigmoid(xp[:, H:2 * H] + g[:, H:2 * H])
    n = torch.tanh(xp[:, 2 * H:] + r * g[:, 2 * H:])
    return (1.0 - z) * n + z * h


def gru_layer_dir(x: Tensor, h0: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor,
                  reverse: bool) -> Tuple[Tensor, Tensor]:
    """x (B,T,I) -> out (B,T,H) time-aligned, h_n (B,H)."""
    B, T, _ = x.shape
    xp = x @ w_ih.t() + b_ih  # hoisted input projection, all timesteps
    h = h0
    outs: List[Optional[Tensor]] = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        h = gru_cell(xp[:, t], h, w_hh, b_hh)
        outs[t] = h
    return torch.stack(outs, 1), h


def gru_forward(sd: Dict[str, Tensor], prefix: str, x: Tensor, h0: Tensor, num_layers: int,
                bidirectional: bool, keep_masks: Optional[List[Tensor]] = None,
                dropout_p: float = 0.0) -> Tuple[Tensor, Tensor]:
    """Multi-layer (bi)GRU, batch_first.  h0 (L*D, B, H) -> out (B,T,D*H), h_n (L*D,B,H).

    keep_masks[l] (B,T,D*H) of 0/1 is applied (scaled by 1/(1-p)) to the output of layer l
    for l < L-1 -- inter-layer dropout exactly where torch.nn.GRU applies it in train mode.
    """
    D = 2 if bidirectional else 1
    inp = x
    h_n = []
    for l in range(num_layers):
        outs = []
        for d in range(D):
            sfx = f"_l{l}" + ("_reverse" if d == 1 else "")
            o, h = gru_layer_dir(inp, h0[l * D + d], sd[prefix + "weight_ih" + sfx],
                                 sd[prefix + "weight_hh" + sfx], sd[prefix + "bias_ih" + sfx],
                                 sd[prefix + "bias_hh" + sfx], reverse=(d == 1))
            outs.append(o)
            h_n.append(h)
        inp = torch.cat(outs, 2) if D == 2 else outs[0]
        if l < num_layers - 1 and keep_masks is not None and dropout_p > 0.0:
            inp = inp * keep_masks[l] / (1.0 - dropout_p)
    return inp, torch.stack(h_n, 0)


# --------------------------------------------------------------------------------------
# MeasureVAE encoder  (MeasureVAE/encoder.py:104-134)
# --------------------------------------------------------------------------------------

def encoder_forward(sd: Dict[str, Tensor], tokens: Tensor, num_layers: int = 2,
                    keep_masks: Optional[List[Tensor]] = None, dropout_p: float = 0.0,
                    prefix: str = "encoder.") -> Tuple[Tensor, Tensor]:
    """tokens (B,24) int64 -> (mu, log_std) each (B,Z).  scale = exp(log_std) (encoder.py:133)."""
    B = tokens.shape[0]
    emb = sd[prefix + "note_embedding_layer.weight"][tokens]                 # encoder.py:93-102
    H = sd[prefix + "lstm.weight_hh_l0"].shape[1]
    h0 = torch.zeros(num_layers * 2, B, H, dtype=emb.dtype)                  # encoder.py:80-91
    _, h_n = gru_forward(sd, prefix + "lstm.", emb, h0, num_layers, True, keep_masks, dropout_p)
    hidden = h_n.transpose(0, 1).contiguous().view(B, -1)                    # encoder.py:126-127
    mu = linear(selu(linear(hidden, sd[prefix + "linear_mean.0.weight"], sd[prefix + "linear_mean.0.bias"])),
                sd[prefix + "linear_mean.2.weight"], sd[prefix + "linear_mean.2.bias"])
    log_std = linear(selu(linear(hidden, sd[prefix + "linear_log_std.0.weight"], sd[prefix + "linear_log_std.0.bias"])),
                     sd[prefix + "linear_log_std.2.weight"], sd[prefix + "linear_log_std.2.bias"])
    return mu, log_std


# --------------------------------------------------------------------------------------
# MeasureVAE hierarchical decoder  (MeasureVAE/decoder.py:392-529)
# --------------------------------------------------------------------------------------

def argmax_lowest(x: Tensor) -> Tensor:
    """Documented tie rule: the LOWEST index among maxima.
    The reference uses topk(k=1) (decoder.py:511), whose tie-breaking is implementation
    defined; on rows with a strict top-1 margin both agree."""
    return torch.argmax(x, dim=1)


def decoder_forward(sd: Dict[str, Tensor], z: Tensor, tokens: Optional[Tensor], teacher_forced: bool,
                    num_layers: int = 2, beat_keep_masks: Optional[List[Tensor]] = None,
                    tick_keep_mask: Optional[Tensor] = None, dropout_p: float = 0.0,
                    prefix: str = "decoder.") -> Tuple[Tensor, Tensor]:
    """z (B,Z) -> weights (B,24,V) post-ReLU logits, samples (B,1,24) int64.

    beat_keep_masks: [ (B,4,H) ] dropout on beat-GRU layer-0 output (train mode).
    tick_keep_mask: (B,24,H) dropout on tick-GRU layer-0 output, one independent mask per
    single-step call (decoder.py:498 calls rnn_tick with seq_len 1, 24 times).
    """
    assert num_layers == 2
    B = z.shape[0]
    H = sd[prefix + "rnn_beat.weight_hh_l0"].shape[1]
    # beat rnn (decoder.py:455-471)
    h0 = selu(linear(z, sd[prefix + "z_to_beat_rnn_input.0.weight"], sd[prefix + "z_to_beat_rnn_input.0.bias"]))
    h0 = h0.view(B, num_layers, -1).transpose(0, 1).contiguous()             # decoder.py:408-409
    beat_in = sd[prefix + "b_0"].unsqueeze(0).expand(B, 4, 1)
    beat_out, _ = gru_forward(sd, prefix + "rnn_beat.", beat_in, h0, num_layers, False,
                              beat_keep_masks, dropout_p)
    # tick rnn (decoder.py:473-529)
    emb = sd[prefix + "note_embedding_layer.weight"]
    w_ih0, w_hh0 = sd[prefix + "rnn_tick.weight_ih_l0"], sd[prefix + "rnn_tick.weight_hh_l0"]
    b_ih0, b_hh0 = sd[prefix + "rnn_tick.bias_ih_l0"], sd[prefix + "rnn_tick.bias_hh_l0"]
    w_ih1, w_hh1 = sd[prefix + "rnn_tick.weight_ih_l1"], sd[prefix + "rnn_tick.weight_hh_l1"]
    b_ih1, b_hh1 = sd[prefix + "rnn_tick.bias_ih_l1"], sd[prefix + "rnn_tick.bias_hh_l1"]
    w_v, b_v = sd[prefix + "tick_emb_to_note_emb.0.weight"], sd[prefix + "tick_emb_to_note_emb.0.bias"]
    tick_in = sd[prefix + "x_0"].unsqueeze(0).expand(B, -1)                   # decoder.py:488-492
    weights, samples = [], []
    for i in range(4):
        bo = beat_out[:, i]
        hid = selu(linear(bo, sd[prefix + "beat_emb_to_tick_rnn_hidden.0.weight"],
                          sd[prefix + "beat_emb_to_tick_rnn_hidden.0.bias"]))
        hid = hid.view(B, num_layers, -1).transpose(0, 1)                     # decoder.py:494
        h_l0, h_l1 = hid[0], hid[1]
        beat_emb = selu(linear(bo, sd[prefix + "beat_emb_to_tick_rnn_input.0.weight"],
                               sd[prefix + "beat_emb_to_tick_rnn_input.0.bias"]))  # decoder.py:495
        for j in range(6):
            t = 6 * i + j
            x = torch.cat((tick_in, beat_emb), 1)                             # decoder.py:497
            h_l0 = gru_cell(x @ w_ih0.t() + b_ih0, h_l0, w_hh0, b_hh0)
            y0 = h_l0
            if tick_keep_mask is not None and dropout_p > 0.0:
                y0 = y0 * tick_keep_mask[:, t] / (1.0 - dropout_p)
            h_l1 = gru_cell(y0 @ w_ih1.t() + b_ih1, h_l1, w_hh1, b_hh1)
            probs = torch.relu(linear(h_l1, w_v, b_v))                        # decoder.py:369-372,499
            if teacher_forced:
                idx = tokens[:, t]                                            # decoder.py:501-504
            else:
                idx = argmax_lowest(probs.detach())                           # decoder.py:510-511
            tick_in = emb[idx]                                                # decoder.py:519
            weights.append(probs)
            samples.append(idx)
    return torch.stack(weights, 1), torch.stack(samples, 1).unsqueeze(1)


# --------------------------------------------------------------------------------------
# MeasureVAE forward + loss  (MeasureVAE/measure_vae.py:97-134; MeasureVAE/vae_trainer.py:16-40,
# 128-139; utils/trainer.py:271-306)
# --------------------------------------------------------------------------------------

def mvae_forward(sd, tokens, eps, teacher_forced, train_dropout=None, dropout_p=0.5):
    """Returns weights, samples, mu, log_std, z_tilde.
    train_dropout: None (eval) or dict(enc=[(B,24,2H)], beat=[(B,4,H)], tick=(B,24,H)) keep-masks."""
    td = train_dropout or {}
    p = dropout_p if train_dropout is not None else 0.0
    mu, log_std = encoder_forward(sd, tokens, 2, td.get("enc"), p)
    z = mu + torch.exp(log_std) * eps                                          # measure_vae.py:119 (rsample)
    weights, samples = decoder_forward(sd, z, tokens, teacher_forced, 2, td.get("beat"), td.get("tick"), p)
    return weights, samples, mu, log_std, z


def mean_crossentropy_loss(weights: Tensor, targets: Tensor) -> Tensor:
    # utils/trainer.py:271-288 and :344-358 (the _alt variant only differs in rank)
    V = weights.shape[-1]
    w = weights.reshape(-1, V)
    t = targets.reshape(-1)
    lse = torch.logsumexp(w, dim=1)
    return (lse - w.gather(1, t[:, None])[:, 0]).mean()


def mean_accuracy(weights: Tensor, targets: Tensor) -> Tensor:
    # utils/trainer.py:290-306: weights.max(1) -> first maximal index
    V = weights.shape[-1]
    pred = torch.argmax(weights.reshape(-1, V), dim=1)
    return (pred == targets.reshape(-1)).float().mean()


def kld_loss(mu: Tensor, log_std: Tensor, beta: float = 0.001) -> Tensor:
    # vae_trainer.py:128-139: KL(N(mu, e^s) || N(0,1)) = 0.5 (e^{2s} + mu^2 - 1) - s
    kld = 0.5 * (torch.exp(2.0 * log_std) + mu * mu - 1.0) - log_std
    return beta * kld.sum(1).mean()


def mvae_loss(weights, tokens, mu, log_std):
    return mean_crossentropy_loss(weights, tokens) + kld_loss(mu, log_std)     # vae_trainer.py:33-36


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam default (utils/trainer.py:32-35), in place.  step is 1-based."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------
# LatentRNN  (LatentRNN/latent_rnn.py:110-263)
# --------------------------------------------------------------------------------------

def latent_rnn_forward(sd, past, future, target, n_gen, eps_past, eps_future, num_layers=2,
                       ctx_keep_masks=None, gen_keep_masks=None, dropout_p=0.0, vae_dropout=None,
                       vae_dropout_p=0.0, only=None, fed_tokens=None):
    """Non-autoregressive LatentRNN (auto_reg=False: what the evaluation scripts load,
    test_reconstruction.py:141).  past (B,np,24), future (B,nf,24) int64.
    eps_* (B,n,Z): injected rsample noise (latent_rnn.py:172 samples even in eval).
    Returns weights (B,n_gen,24,V), samples (B,1,24*n_gen), z_out (B,n_gen,Z).
    The target-encode of latent_rnn.py:133 does not influence any output in this mode.
    Train-mode dropout (keep masks, 1 = keep): ctx_keep_masks = {"past": [(B,np,2Hc)], "future": [(B,nf,2Hc)]},
    gen_keep_masks = [(B,n_gen,2Hg)] with dropout_p; vae_dropout = {"enc_past": (B*np,24,2H), "enc_future":
    (B*nf,24,2H), "dec": [(beat (B,4,H), tick (B,24,H)) per gap measure]} with vae_dropout_p.
    fed_tokens (B,n_gen,24): test hook -- decode with THESE tokens fed back instead of the oracle's own argmax (no
    gradient flows through the argmax, so this is the same function of the parameters along a given token path).
    only="past"/"future": LatentRNNAblations (latent_rnn_ablations.py:143-146), one context seeds the generation GRU."""
    B = past.shape[0]
    vp = "vae_model."
    vd = vae_dropout or {}     # train mode recurses into the frozen VAE (utils/trainer.py:78): its dropout is active too

    def z_seq(m, eps, keep):                                                   # latent_rnn.py:161-174
        n = m.shape[1]
        mu, ls = encoder_forward(sd, m.reshape(-1, 24), 2, [keep] if keep is not None else None,
                                 vae_dropout_p if keep is not None else 0.0, prefix=vp + "encoder.")
        return (mu + torch.exp(ls) * eps.reshape(-1, eps.shape[-1])).view(B, n, -1)

    zp, zf = z_seq(past, eps_past, vd.get("enc_past")), z_seq(future, eps_future, vd.get("enc_future"))
    Hc = sd["context_rnn_past.weight_hh_l0"].shape[1]
    h0 = torch.zeros(num_layers * 2, B, Hc, dtype=zp.dtype)
    km = ctx_keep_masks or {}
    _, hp = gru_forward(sd, "context_rnn_past.", zp, h0, num_layers, True, km.get("past"), dropout_p)
    _, hf = gru_forward(sd, "context_rnn_future.", zf, h0, num_layers, True, km.get("future"), dropout_p)
    ctx = torch.cat((hp, hf), 2) if only is None else (hp if only == "past" else hf)   # latent_rnn.py:140
    x = sd["x_0"].expand(B, n_gen, -1)                                         # latent_rnn.py:228
    out, _ = gru_forward(sd, "generation_rnn.", x, ctx, num_layers, True, gen_keep_masks, dropout_p)
    z_out = linear(out.reshape(B * n_gen, -1), sd["generation_linear.weight"],
                   sd["generation_linear.bias"]).view(B, n_gen, -1)           # latent_rnn.py:232-233
    ws, ss = [], []
    for i in range(n_gen):                                                     # latent_rnn.py:237-240
        bm, tm = vd["dec"][i] if "dec" in vd else (None, None)
        dkw = dict(beat_keep_masks=[bm], tick_keep_mask=tm, dropout_p=vae_dropout_p) if bm is not None else {}
        if fed_tokens is None:
            w, s = decoder_forward(sd, z_out[:, i], None, False, 2, prefix=vp + "decoder.", **dkw)
        else:
            w, s = decoder_forward(sd, z_out[:, i], fed_tokens[:, i], True, 2, prefix=vp + "decoder.", **dkw)
        ws.append(w)
        ss.append(s)
    return torch.stack(ws, 1), torch.cat(ss, 2), z_out


def latent_rnn_forward_autoreg(sd, past, future, target, n_gen, eps_past, eps_future, eps_target, eps_regen,
                               teacher_forcing, num_layers=2, only=None):
    """Autoregressive LatentRNN (auto_reg=True, the train_inpaintnet.py default; latent_rnn.py:142-153,219-261),
    eval-mode dropout.  teacher_forcing=True: the generation GRU reads [z_past[-1], z_target[:-1]] in one call
    (latent_rnn.py:148-149,230-240).  teacher_forcing=False: per gap measure one GRU call of length 1 with the
    hidden state carried, linear, argmax decode, and the decoded tokens re-encoded by the frozen VAE
    (rsample noise eps_regen[i] (B,Z)) as the next input (latent_rnn.py:246-260).
    eps_target (B,n_t,Z) is the noise of the target encode.  Returns weights, samples, z_out as above."""
    B = past.shape[0]
    vp = "vae_model."

    def z_seq(m, eps):                                                         # latent_rnn.py:161-174
        n = m.shape[1]
        mu, ls = encoder_forward(sd, m.reshape(-1, 24), 2, None, 0.0, prefix=vp + "encoder.")
        return (mu + torch.exp(ls) * eps.reshape(-1, eps.shape[-1])).view(B, n, -1)

    zp, zf = z_seq(past, eps_past), z_seq(future, eps_future)
    Hc = sd["context_rnn_past.weight_hh_l0"].shape[1]
    h0 = torch.zeros(num_layers * 2, B, Hc, dtype=zp.dtype)
    _, hp = gru_forward(sd, "context_rnn_past.", zp, h0, num_layers, True)
    _, hf = gru_forward(sd, "context_rnn_future.", zf, h0, num_layers, True)
    hidden = torch.cat((hp, hf), 2) if only is None else (hp if only == "past" else hf)   # latent_rnn.py:140
    lw, lb = sd["generation_linear.weight"], sd["generation_linear.bias"]
    ws, ss = [], []
    if teacher_forcing:
        zt = z_seq(target, eps_target)
        seed = torch.cat((zp[:, -1:], zt[:, :-1]), 1)                          # latent_rnn.py:149
        out, _ = gru_forward(sd, "generation_rnn.", seed, hidden, num_layers, True)
        z_out = linear(out.reshape(B * n_gen, -1), lw, lb).view(B, n_gen, -1)
        for i in range(n_gen):
            w, s = decoder_forward(sd, z_out[:, i], None, False, 2, prefix=vp + "decoder.")
            ws.append(w)
            ss.append(s)
    else:
        x = zp[:, -1:]                                                         # latent_rnn.py:151
        zs = []
        for i in range(n_gen):
            out, hidden = gru_forward(sd, "generation_rnn.", x, hidden, num_layers, True)
            gz = linear(out.reshape(B, -1), lw, lb)
            zs.append(gz.unsqueeze(1))
            w, s = decoder_forward(sd, gz, None, False, 2, prefix=vp + "decoder.")
            ws.append(w)
            ss.append(s)
            x = z_seq(s, eps_regen[i])                                         # latent_rnn.py:259
        z_out = torch.cat(zs, 1)
    return torch.stack(ws, 1), torch.cat(ss, 2), z_out


# --------------------------------------------------------------------------------------
# LSTM + AnticipationRNN teacher-forced forward
# (AnticipationRNN/anticipation_rnn_gauss_reg_model.py:14-39,348-404,437-532)
# --------------------------------------------------------------------------------------

def lstm_layer(x: Tensor, w_ih, w_hh, b_ih, b_hh) -> Tensor:
    """torch.nn.LSTM single layer, zero initial state, gate rows [i; f; g; o]. x (B,T,I)->(B,T,H)."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    xp = x @ w_ih.t() + b_ih
    h = torch.zeros(B, H, dtype=x.dtype)
    c = torch.zeros(B, H, dtype=x.dtype)
    outs = []
    for t in range(T):
        g = xp[:, t] + h @ w_hh.t() + b_hh
        i_, f_, g_, o_ = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
        c = f_ * c + i_ * g_
        h = o_ * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, 1)


def arnn_forward_tf(sd, score, metadata, constraints_loc, num_layers=2, keep_input_steps=None,
                    dropout_input_p=0.0):
    """Teacher-forced AnticipationRNN forward for ONE voice (num_voices == 1).
    score (B,1,T) int64, metadata (B,1,T,3) int64 [beat-marker, tick, voice-index],
    constraints_loc (B,1,T) 0/1.  Returns logits (B,T,V) (no ReLU on the last layer).
    keep_input_steps (B,T) 0/1: whole-timestep input dropout (arnn_model.py:437-442)."""
    B, _, T = score.shape
    tok = score[:, 0]
    V1 = sd["note_embeddings.0.weight"].shape[0]                               # V + 1 (mask id V)
    # mask_tensor_score (arnn_model.py:510-532): unconstrained -> extra id V
    masked = torch.where(constraints_loc[:, 0] > 0, tok, torch.full_like(tok, V1 - 1))
    emb_masked = sd["note_embeddings.0.weight"][masked]
    md = metadata[:, 0]
    embs_m = [sd[f"metadata_embeddings.{k}.weight"][md[:, :, k]] for k in range(md.shape[-1])]
    cin = torch.cat(embs_m + [emb_masked], 2)                                  # arnn_model.py:477-508
    # constraint stack on the time-flipped sequence (arnn_model.py:455-475)
    x = torch.flip(cin, [1])
    for l in range(num_layers):
        x = lstm_layer(x, sd[f"lstm_constraint.{l}.weight_ih_l0"], sd[f"lstm_constraint.{l}.weight_hh_l0"],
                       sd[f"lstm_constraint.{l}.bias_ih_l0"], sd[f"lstm_constraint.{l}.bias_hh_l0"])
    cout = torch.flip(x, [1])
    emb = sd["note_embeddings.0.weight"][tok]
    shifted = torch.cat((torch.zeros_like(emb[:, :1]), emb[:, :-1]), 1)       # arnn_model.py:367-371
    if keep_input_steps is not None and dropout_input_p > 0.0:
        shifted = shifted * keep_input_steps[:, :, None] / (1.0 - dropout_input_p)
    x = torch.cat((shifted, cout), 2)                                          # arnn_model.py:375
    for l in range(num_layers):
        x = lstm_layer(x, sd[f"lstm_generation.{l}.weight_ih_l0"], sd[f"lstm_generation.{l}.weight_hh_l0"],
                       sd[f"lstm_generation.{l}.bias_ih_l0"], sd[f"lstm_generation.{l}.bias_hh_l0"])
    hid = torch.relu(linear(x, sd["linear_1.weight"], sd["linear_1.bias"]))    # arnn_model.py:388-392
    return linear(hid, sd["linear_ouput_notes.0.weight"], sd["linear_ouput_notes.0.bias"])  # :396-400


def lstm_cell(xp, h, c, w_hh, b_hh):
    H = h.shape[1]
    g = xp + h @ w_hh.t() + b_hh
    i_, f_, g_, o_ = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
    c = f_ * c + i_ * g_
    return o_ * torch.tanh(c), c


def arnn_forward_no_tf(sd, score, metadata, constraints_loc, num_layers=2):
    """AnticipationRNN without teacher forcing (arnn_model.py:190-259), one voice.  The token fed back to the WHOLE
    batch at every tick is the argmax of batch element 0 (:252-256); the start symbol is id 0 (:220).
    Returns logits (B,T,V) and the fed-back tokens (T,)."""
    B, _, T = score.shape
    tok = score[:, 0]
    V1 = sd["note_embeddings.0.weight"].shape[0]
    masked = torch.where(constraints_loc[:, 0] > 0, tok, torch.full_like(tok, V1 - 1))
    md = metadata[:, 0]
    embs_m = [sd[f"metadata_embeddings.{k}.weight"][md[:, :, k]] for k in range(md.shape[-1])]
    x = torch.flip(torch.cat(embs_m + [sd["note_embeddings.0.weight"][masked]], 2), [1])
    for l in range(num_layers):
        x = lstm_layer(x, sd[f"lstm_constraint.{l}.weight_ih_l0"], sd[f"lstm_constraint.{l}.weight_hh_l0"],
                       sd[f"lstm_constraint.{l}.bias_ih_l0"], sd[f"lstm_constraint.{l}.bias_hh_l0"])
    cout = torch.flip(x, [1])
    H = cout.shape[2]
    hs = [torch.zeros(B, H, dtype=cout.dtype) for _ in range(num_layers)]
    cs = [torch.zeros(B, H, dtype=cout.dtype) for _ in range(num_layers)]
    cur = 0
    outs, fed = [], []
    for t in range(T):
        fed.append(cur)
        inp = torch.cat((sd["note_embeddings.0.weight"][cur].expand(B, -1), cout[:, t]), 1)
        for l in range(num_layers):
            xp = inp @ sd[f"lstm_generation.{l}.weight_ih_l0"].t() + sd[f"lstm_generation.{l}.bias_ih_l0"]
            hs[l], cs[l] = lstm_cell(xp, hs[l], cs[l], sd[f"lstm_generation.{l}.weight_hh_l0"], sd[f"lstm_generation.{l}.bias_hh_l0"])
            inp = hs[l]
        w = linear(torch.relu(linear(inp, sd["linear_1.weight"], sd["linear_1.bias"])),
                   sd["linear_ouput_notes.0.weight"], sd["linear_ouput_notes.0.bias"])
        outs.append(w)
        cur = int(torch.argmax(w[0].detach()))
    return torch.stack(outs, 1), torch.tensor(fed)


def arnn_forward_inpaint(sd, score, metadata, constraints_loc, start_tick, end_tick, num_layers=2):
    """AnticipationRNN inpainting inference (arnn_model.py:261-346), one voice, eval mode (no input dropout).
    Ticks [0, start) run teacher forced (input = embedding of score[t-1], zeros at t = 0, :286-299); ticks
    [start, end) run one at a time: the input token of tick `start` is score[start-1] per batch element, later
    ticks feed back the argmax of batch element 0 to the whole batch (:318-343).
    Returns logits (B, end-start, V) and the generated score (B,1,T)."""
    B, _, T = score.shape
    tok = score[:, 0]
    emb_w = sd["note_embeddings.0.weight"]
    V1 = emb_w.shape[0]
    masked = torch.where(constraints_loc[:, 0] > 0, tok, torch.full_like(tok, V1 - 1))
    md = metadata[:, 0]
    embs_m = [sd[f"metadata_embeddings.{k}.weight"][md[:, :, k]] for k in range(md.shape[-1])]
    x = torch.flip(torch.cat(embs_m + [emb_w[masked]], 2), [1])
    for l in range(num_layers):
        x = lstm_layer(x, sd[f"lstm_constraint.{l}.weight_ih_l0"], sd[f"lstm_constraint.{l}.weight_hh_l0"],
                       sd[f"lstm_constraint.{l}.bias_ih_l0"], sd[f"lstm_constraint.{l}.bias_hh_l0"])
    cout = torch.flip(x, [1])
    H = cout.shape[2]
    hs = [torch.zeros(B, H, dtype=cout.dtype) for _ in range(num_layers)]
    cs = [torch.zeros(B, H, dtype=cout.dtype) for _ in range(num_layers)]

    def step(note_emb, t):
        inp = torch.cat((note_emb, cout[:, t]), 1)
        for l in range(num_layers):
            xp = inp @ sd[f"lstm_generation.{l}.weight_ih_l0"].t() + sd[f"lstm_generation.{l}.bias_ih_l0"]
            hs[l], cs[l] = lstm_cell(xp, hs[l], cs[l], sd[f"lstm_generation.{l}.weight_hh_l0"], sd[f"lstm_generation.{l}.bias_hh_l0"])
            inp = hs[l]
        return inp

    for t in range(start_tick):                                               # arnn_model.py:286-299
        step(emb_w[tok[:, t - 1]] if t > 0 else torch.zeros(B, emb_w.shape[1], dtype=cout.dtype), t)
    gen = score.clone()
    outs = []
    for t in range(start_tick, end_tick):                                     # arnn_model.py:304-343
        prev = gen[:, 0, t - 1] if t > 0 else torch.zeros(B, dtype=torch.long)
        out = step(emb_w[prev], t)
        w = linear(torch.relu(linear(out, sd["linear_1.weight"], sd["linear_1.bias"])),
                   sd["linear_ouput_notes.0.weight"], sd["linear_ouput_notes.0.bias"])
        outs.append(w)
        gen[:, 0, t] = int(torch.argmax(w[0]))
    return torch.stack(outs, 1), gen

"""LatentRNN hot path (LatentRNN/latent_rnn.py:110-263) on the same kernels as MeasureVAE:
frozen-VAE encode of the context measures (one batched encoder call), two 2-layer bidirectional
context GRUs (H=512), the 2-layer bidirectional generation GRU (H=1024) whose initial state is the
concatenation of the context final states (written in place by the context kernels' epilogues), the
generation linear, and ONE batched argmax decode of all gap measures.  Time-major rows m*B + b.
"""
import torch

from . import engine, ops
from .engine import _lin, _dgrad, _wgrad, _gru_wgrads, NOISE, BIG
from .ops import F32, ACT_NONE, CORE_SIMT, ATOMIC_ADD

SFX = ("", "_reverse")


def bigru2_forward(arena, pfx, prec, H, T, B, x, hseq, training, p_drop, need_grad, finals=None, want_y1=False,
                   persistent=True):
    """2-layer bidirectional GRU over T steps of B rows.
    x = ("matrix", ptr, ld, I)  time-major [T*B, I] activations, or ("scalar", name_of_x0) for the
        rank-1 input of the non-autoregressive generation GRU (latent_rnn.py:228).
    hseq: tensor [2, 2, (T+1)*B, H] (layer, direction) whose h0 slots are already filled.
    finals[l][d] = (ptr, dt, ld, col0) destination of the final hidden state, or None."""
    dev, act = hseq.device, prec.tdt
    drop = training and p_drop > 0.0
    scale = 1.0 / (1.0 - p_drop) if drop else 1.0
    gates = torch.empty(2, 2, T * B, ops.gates_cols(H), dtype=act, device=dev) if need_grad else None
    y0 = torch.empty(T * B, 2 * H, dtype=act, device=dev)
    y1 = torch.empty(T * B, 2 * H, dtype=act, device=dev) if want_y1 else None
    mask0 = NOISE.keep_mask(arena, (T * B, 2 * H), p_drop, dev) if drop else None
    P0 = pvecs = None
    if x[0] == "matrix":
        _, xp, ldx, I = x
        P0 = torch.empty(2, T * B, 3 * H, dtype=act, device=dev)
        for d, s in enumerate(SFX):
            _lin(prec, xp, ldx, T * B, I, arena.w(prec, pfx + "weight_ih_l0" + s), 3 * H, P0[d].data_ptr(), prec.act, 3 * H,
                 bias=arena.fptr(pfx + "bias_ih_l0" + s))
    else:
        x0_name = x[1]

        def build():
            pv = torch.empty(2, 3 * H, dtype=torch.float32, device=dev)
            for d, s in enumerate(SFX):
                ops.gemm(CORE_SIMT, F32, 1, 3 * H, [(arena.fptr(x0_name), 1, 0, arena.fptr(pfx + "weight_ih_l0" + s), 1, 0, 1)],
                         pv[d].data_ptr(), F32, 3 * H, bias=arena.fptr(pfx + "bias_ih_l0" + s))
            return pv

        pvecs = arena.derived((pfx, "pvec"), build)

    def fin(l, d):
        f = finals[l][d] if finals is not None else None
        return dict(final_out=f[0], final_dt=f[1], ld_final=f[2], final_col0=f[3]) if f is not None else {}

    dirs = []
    for d, s in enumerate(SFX):
        src = dict(P=P0[d].data_ptr(), ldP=3 * H) if P0 is not None else dict(pvec=pvecs[d].data_ptr())
        dirs.append(ops.gru_dir(arena.w(prec, pfx + "weight_hh_l0" + s)[0], arena.fptr(pfx + "bias_hh_l0" + s),
                                hseq[0, d].data_ptr(), gates=gates[0, d].data_ptr() if need_grad else 0, reverse=d,
                                y_col0=d * H, **src, **fin(0, d)))
    pk0 = ops.gru_layer_fwd(prec, T, B, H, dirs, y=y0.data_ptr(), ld_y=2 * H, mask=mask0.data_ptr() if drop else 0,
                            ld_mask=2 * H, mask_scale=scale, persistent=persistent)
    P1 = torch.empty(2, T * B, 3 * H, dtype=act, device=dev)
    for d, s in enumerate(SFX):
        _lin(prec, y0.data_ptr(), 2 * H, T * B, 2 * H, arena.w(prec, pfx + "weight_ih_l1" + s), 3 * H, P1[d].data_ptr(),
             prec.act, 3 * H, bias=arena.fptr(pfx + "bias_ih_l1" + s))
    dirs = [ops.gru_dir(arena.w(prec, pfx + "weight_hh_l1" + s)[0], arena.fptr(pfx + "bias_hh_l1" + s),
                        hseq[1, d].data_ptr(), gates=gates[1, d].data_ptr() if need_grad else 0, P=P1[d].data_ptr(),
                        ldP=3 * H, reverse=d, y_col0=d * H, **fin(1, d)) for d, s in enumerate(SFX)]
    pk1 = ops.gru_layer_fwd(prec, T, B, H, dirs, y=y1.data_ptr() if want_y1 else 0, ld_y=2 * H, persistent=persistent)
    saved = dict(T=T, B=B, H=H, hseq=hseq, gates=gates, y0=y0, mask0=mask0, scale=scale, x=x, pk=(pk0, pk1)) if need_grad else None
    return y1, saved


def bigru2_backward(arena, pfx, prec, saved, dY1=None, dh_n=None, dh0=None):
    """dY1 = (ptr, ld) gradient wrt the layer-1 output sequence; dh_n[l][d] = (fp32 ptr, ld) gradient wrt the
    final states; dh0[l][d] = (ptr, dt, ld) destination of the gradient wrt the initial states."""
    T, B, H = saved["T"], saved["B"], saved["H"]
    hseq, gates, y0, mask0 = saved["hseq"], saved["gates"], saved["y0"], saved["mask0"]
    dev, act, es = hseq.device, prec.tdt, prec.es
    ws = torch.empty(2 * 2 * B * H, dtype=torch.float32, device=dev)

    def extra(l, d):
        kw = {}
        if dh_n is not None and dh_n[l][d] is not None:
            kw.update(dh_n=dh_n[l][d][0], ld_dhn=dh_n[l][d][1])
        if dh0 is not None and dh0[l][d] is not None:
            kw.update(dh0=dh0[l][d][0], dh0_dt=dh0[l][d][1], ld_dh0=dh0[l][d][2])
        return kw

    dP = torch.empty(2, T * B, 3 * H, dtype=act, device=dev)
    dGn = torch.empty(2, T * B, H, dtype=act, device=dev)
    dirs = [ops.gru_bwd_dir(arena.w(prec, pfx + "weight_hh_l1" + s)[0], hseq[1, d].data_ptr(), gates[1, d].data_ptr(),
                            dP[d].data_ptr(), dGn[d].data_ptr(), reverse=d, y_col0=d * H, **extra(1, d))
            for d, s in enumerate(SFX)]
    ops.gru_layer_bwd(prec, T, B, H, dirs, ws.data_ptr(), dY=dY1[0] if dY1 is not None else 0,
                      ld_dy=dY1[1] if dY1 is not None else 0, persistent=saved["pk"][1])
    for d, s in enumerate(SFX):
        hprev = hseq[1, d].data_ptr() + (es * B * H if d == 1 else 0)
        _gru_wgrads(arena, prec, pfx, "_l1" + s, H, T * B, dP[d].data_ptr(), dGn[d].data_ptr(), hprev, X=y0.data_ptr(),
                    ld_x=2 * H, K_in=2 * H)
    dY0 = torch.empty(T * B, 2 * H, dtype=act, device=dev)
    wf, wr = arena.w(prec, pfx + "weight_ih_l1"), arena.w(prec, pfx + "weight_ih_l1_reverse")
    ops.gemm(prec.core, prec.act, T * B, 2 * H,
             [(dP[0].data_ptr(), 3 * H, 0, wf[0], wf[1], 1, 3 * H), (dP[1].data_ptr(), 3 * H, 0, wr[0], wr[1], 1, 3 * H)],
             dY0.data_ptr(), prec.act, 2 * H)
    dirs = [ops.gru_bwd_dir(arena.w(prec, pfx + "weight_hh_l0" + s)[0], hseq[0, d].data_ptr(), gates[0, d].data_ptr(),
                            dP[d].data_ptr(), dGn[d].data_ptr(), reverse=d, y_col0=d * H, **extra(0, d))
            for d, s in enumerate(SFX)]
    ops.gru_layer_bwd(prec, T, B, H, dirs, ws.data_ptr(), dY=dY0.data_ptr(), ld_dy=2 * H,
                      mask=mask0.data_ptr() if mask0 is not None else 0, ld_mask=2 * H, mask_scale=saved["scale"],
                      persistent=saved["pk"][0])
    x = saved["x"]
    for d, s in enumerate(SFX):
        hprev = hseq[0, d].data_ptr() + (es * B * H if d == 1 else 0)
        if x[0] == "matrix":
            _gru_wgrads(arena, prec, pfx, "_l0" + s, H, T * B, dP[d].data_ptr(), dGn[d].data_ptr(), hprev, X=x[1], ld_x=x[2],
                        K_in=x[3])
        else:
            _gru_wgrads(arena, prec, pfx, "_l0" + s, H, T * B, dP[d].data_ptr(), dGn[d].data_ptr(), hprev)
            n_ih = pfx + "weight_ih_l0" + s
            if arena.wants_grad(n_ih):
                sv = torch.zeros(3 * H, dtype=torch.float32, device=dev)
                ops.colsum(dP[d].data_ptr(), prec.act, 3 * H, T * B, 3 * H, sv.data_ptr())
                ops.gemm(CORE_SIMT, F32, 3 * H, 1, [(sv.data_ptr(), 1, 0, arena.fptr(x[1]), 1, 0, 1)], arena.gptr(n_ih), F32, 1,
                         accumulate=ATOMIC_ADD, split_k=1)
                ops.gemm(CORE_SIMT, F32, 1, 1, [(sv.data_ptr(), 3 * H, 0, arena.fptr(n_ih), 3 * H, 0, 3 * H)], arena.gptr(x[1]),
                         F32, 1, accumulate=ATOMIC_ADD, split_k=32)


def _encode_z(arena, prec, model, tokens, n_noise_rows):
    """Frozen-VAE encode + rsample (latent_rnn.py:161-174) of tokens (R,24) -> z (R,Z) in the activation dtype.
    n_noise_rows: the row counts of the separate rsample draws that make up R (one NOISE.normal call each, so
    injected noise lines up with the reference's get_z_seq calls)."""
    enc = model.vae_model.encoder
    Z = model.z_dim
    mu, ls, _ = engine.encoder_forward(arena, "vae_model.encoder.", prec, enc._cfg(), tokens, enc.training, False)
    parts = [NOISE.normal(arena, (r, Z), tokens.device) for r in n_noise_rows if r > 0]
    eps = parts[0] if len(parts) == 1 else torch.cat(parts, 0)
    z = torch.empty(tokens.shape[0], Z, dtype=prec.tdt, device=tokens.device)
    ops.reparam_fwd(mu.data_ptr(), ls.data_ptr(), eps.data_ptr(), mu.numel(), 0, z.data_ptr(), prec.act)
    return z


def latent_forward(arena, prec, model, past, future, n_gen, training, need_grad, target=None, teacher_forcing=False):
    """past (B,np,24), future (B,nf,24) int64 cuda -> weights (B,n_gen,24,V), samples (B,1,24*n_gen), z_out (B,n_gen,Z).
    Three generation modes (latent_rnn.py:219-261):
      auto_reg=False                    rank-1 input x_0, all gap measures in one GRU call + one batched decode
      auto_reg=True, teacher_forcing    input = [z_past[-1], z_target[:-1]], otherwise as above
      auto_reg=True, no teacher forcing per gap measure: one GRU step, linear, argmax decode, re-encode the decoded
                                        measure (frozen VAE, fresh rsample) as the next input"""
    vae = model.vae_model
    dec = vae.decoder
    dcfg = dec._cfg()
    # LatentRNNAblations (latent_rnn_ablations.py:79,143-146): only one context GRU seeds the generation GRU, whose
    # hidden size is then that of the context GRU
    only = getattr(model, "type", None)
    Z, Hc = model.z_dim, model.rnn_hidden_size
    Hg = Hc if only is not None else Hc * model.num_rnn_layers
    assert model.num_rnn_layers == 2, "the generation GRU initial state only lines up for 2 layers (SURVEY.md a14)"
    B, n_p, _ = past.shape
    n_f = future.shape[1]
    dev, act, es = past.device, prec.tdt, prec.es
    auto_reg = bool(model.auto_reg)
    step_mode = auto_reg and not teacher_forcing
    T = n_gen
    arena.refresh()
    # ---- frozen VAE encode of all context measures in ONE batch, time-major rows m*B + b (latent_rnn.py:131-133)
    parts = [past.transpose(0, 1), future.transpose(0, 1)]
    n_t = 0
    if auto_reg and teacher_forcing and T > 1:   # only z_target[:-1] feeds the generation GRU (latent_rnn.py:149)
        assert target is not None and target.shape[1] >= T - 1, "teacher forcing needs the target measures"
        n_t = T - 1
        parts.append(target[:, :n_t].transpose(0, 1))
    ctx_tokens = torch.cat(parts, 0).reshape((n_p + n_f + n_t) * B, 24).contiguous()
    z_ctx = _encode_z(arena, prec, model, ctx_tokens, (n_p * B, n_f * B, n_t * B))
    # ---- state buffers; the generation GRU's h0 slots are filled by the context GRUs' final-state stores
    Tg = 1 if step_mode else T                      # steps per generation-GRU call
    hs_g = [torch.empty(2, 2, (Tg + 1) * B, Hg, dtype=act, device=dev) for _ in range(T if step_mode else 1)]

    def h0_ptr(i, l, d):  # forward direction: slot 0, reverse direction: slot Tg
        return hs_g[i][l, d].data_ptr() + (es * Tg * B * Hg if d == 1 else 0)

    saved_ctx = []
    ctxs = (("context_rnn_past.", (n_p, 0, 0)), ("context_rnn_future.", (n_f, n_p * B, Hc)))
    if only is not None:   # the other context GRU's output is discarded by the reference: not run at all
        ctxs = (("context_rnn_past.", (n_p, 0, 0)),) if only == "past" else (("context_rnn_future.", (n_f, n_p * B, 0)),)
    for which, (n_m, row0, col0) in ctxs:
        hs = torch.empty(2, 2, (n_m + 1) * B, Hc, dtype=act, device=dev)
        hs[:, 0, :B].zero_()
        hs[:, 1, n_m * B:].zero_()
        finals = [[(h0_ptr(0, l, d), prec.act, Hg, col0) for d in range(2)] for l in range(2)]
        _, sv = bigru2_forward(arena, which, prec, Hc, n_m, B, ("matrix", z_ctx.data_ptr() + es * row0 * Z, Z, Z), hs,
                               model.training, model.dropout, need_grad, finals=finals)
        saved_ctx.append(sv)
    z_tm = torch.empty(T * B, Z, dtype=torch.float32, device=dev)   # generated latents, time-major rows m*B + b
    w_lin, b_lin = arena.w(prec, "generation_linear.weight"), arena.fptr("generation_linear.bias")
    z_last = z_ctx[(n_p - 1) * B:n_p * B]                           # zp[:, -1] (latent_rnn.py:149,151)
    if not step_mode:
        if auto_reg:
            seed = torch.cat((z_last, z_ctx[(n_p + n_f) * B:]), 0) if n_t else z_last.contiguous()
            x = ("matrix", seed.data_ptr(), Z, Z)
        else:
            seed, x = None, ("scalar", "x_0")
        y1, saved_gen = bigru2_forward(arena, "generation_rnn.", prec, Hg, T, B, x, hs_g[0], model.training,
                                       model.dropout, need_grad, want_y1=True)
        _lin(prec, y1.data_ptr(), 2 * Hg, T * B, 2 * Hg, w_lin, Z, z_tm.data_ptr(), F32, Z, bias=b_lin)
        # ---- ONE batched argmax decode of the T*B gap measures (latent_rnn.py:237-240 loops over measures)
        weights, samples, saved_dec = engine.decoder_forward(arena, "vae_model.decoder.", prec, dcfg, z_tm, None, False,
                                                             dec.training, need_grad, batch_map=(B, T))
        steps = [dict(y1=y1, saved_gen=saved_gen, saved_dec=saved_dec, keep=seed)]
    else:
        steps, ws, ss = [], [], []
        x_in = z_last.contiguous()
        for i in range(T):                                          # latent_rnn.py:246-260
            finals = [[(h0_ptr(i + 1, l, d), prec.act, Hg, 0) for d in range(2)] for l in range(2)] if i + 1 < T else None
            y1, saved_gen = bigru2_forward(arena, "generation_rnn.", prec, Hg, 1, B, ("matrix", x_in.data_ptr(), Z, Z),
                                           hs_g[i], model.training, model.dropout, need_grad, finals=finals, want_y1=True,
                                           persistent=False)
            z_i = z_tm[i * B:(i + 1) * B]
            _lin(prec, y1.data_ptr(), 2 * Hg, B, 2 * Hg, w_lin, Z, z_i.data_ptr(), F32, Z, bias=b_lin)
            w_i, s_i, saved_dec = engine.decoder_forward(arena, "vae_model.decoder.", prec, dcfg, z_i, None, False,
                                                         dec.training, need_grad)
            ws.append(w_i)
            ss.append(s_i)
            steps.append(dict(y1=y1, saved_gen=saved_gen, saved_dec=saved_dec, keep=x_in))
            if i + 1 < T:   # the reference also re-encodes after the last measure; that result is never used
                x_in = _encode_z(arena, prec, model, s_i.view(B, 24), (B,))
        weights, samples = torch.stack(ws, 1), torch.cat(ss, 2)
    z_out = z_tm.view(T, B, Z).transpose(0, 1)
    saved = None
    if need_grad:
        saved = dict(B=B, T=T, n_p=n_p, n_f=n_f, z_ctx=z_ctx, saved_ctx=saved_ctx, steps=steps, step_mode=step_mode,
                     dcfg=dcfg, Hc=Hc, Hg=Hg, Z=Z, ctx=[(w, c[2]) for w, c in ctxs])
    return weights, samples, z_out, saved


def _linear_backward(arena, prec, dz, y1, rows, Hg, Z):
    """generation_linear backward: accumulates its parameter gradients, returns dY1 [rows, 2*Hg]."""
    dev, act = y1.device, prec.tdt
    dz_act = torch.empty(rows, Z, dtype=act, device=dev)
    ops.convert_2d(dz.data_ptr(), F32, Z, dz_act.data_ptr(), prec.act, Z, rows, Z)
    if arena.wants_grad("generation_linear.weight"):
        _wgrad(prec, dz_act.data_ptr(), Z, Z, y1.data_ptr(), 2 * Hg, 2 * Hg, rows, arena.gptr("generation_linear.weight"), 2 * Hg)
        ops.colsum(dz_act.data_ptr(), prec.act, Z, rows, Z, arena.gptr("generation_linear.bias"))
    dY1 = torch.empty(rows, 2 * Hg, dtype=act, device=dev)
    _dgrad(prec, dz_act.data_ptr(), Z, rows, Z, arena.w(prec, "generation_linear.weight"), 2 * Hg, dY1.data_ptr(), prec.act, 2 * Hg)
    return dY1


def latent_backward(arena, prec, saved, dweights, dz_out):
    """Accumulates the LatentRNN parameter gradients (the VAE is frozen: data-gradient only through the decoder;
    nothing flows through the argmax tokens that are re-encoded in the autoregressive loop)."""
    B, T, Hc, Hg, Z = saved["B"], saved["T"], saved["Hc"], saved["Hg"], saved["Z"]
    steps = saved["steps"]
    dev = steps[0]["y1"].device
    arena.refresh()

    def dz_of(k, rows, dw, g):
        dz = None
        if dw is not None:
            dz = engine.decoder_backward(arena, "vae_model.decoder.", prec, saved["dcfg"], steps[k]["saved_dec"], dw, need_dz=True)
        if g is not None:   # gradient arriving on the returned gen_z
            dz = g.contiguous() if dz is None else dz + g
        if dz is None:
            dz = torch.zeros(rows, Z, dtype=torch.float32, device=dev)
        return dz

    def new_dh0():
        return torch.empty(2, 2, B, Hg, dtype=torch.float32, device=dev)

    def as_dst(t):
        return [[(t[l, d].data_ptr(), F32, Hg) for d in range(2)] for l in range(2)]

    if not saved["step_mode"]:
        g = dz_out.transpose(0, 1).reshape(T * B, Z) if dz_out is not None else None   # (B,T,Z) -> time-major rows
        dz = dz_of(0, T * B, dweights, g)
        dY1 = _linear_backward(arena, prec, dz, steps[0]["y1"], T * B, Hg, Z)
        dh0 = new_dh0()                               # gradient wrt the concatenated context states
        bigru2_backward(arena, "generation_rnn.", prec, steps[0]["saved_gen"], dY1=(dY1.data_ptr(), 2 * Hg), dh0=as_dst(dh0))
    else:
        dh0 = None
        for i in range(T - 1, -1, -1):
            dw = dweights[:, i].contiguous() if dweights is not None else None
            g = dz_out[:, i] if dz_out is not None else None
            dz = dz_of(i, B, dw, g)
            dY1 = _linear_backward(arena, prec, dz, steps[i]["y1"], B, Hg, Z)
            dh_n = [[(dh0[l, d].data_ptr(), Hg) for d in range(2)] for l in range(2)] if dh0 is not None else None
            dh0_i = new_dh0()
            bigru2_backward(arena, "generation_rnn.", prec, steps[i]["saved_gen"], dY1=(dY1.data_ptr(), 2 * Hg), dh_n=dh_n,
                            dh0=as_dst(dh0_i))
            dh0 = dh0_i
    engine.grad_ready(arena, ("generation_rnn.", "generation_linear."))   # exchanged under the context GRUs' backward
    for k, (pfx, col0) in enumerate(saved["ctx"]):
        bigru2_backward(arena, pfx, prec, saved["saved_ctx"][k],
                        dh_n=[[(dh0[l, d].data_ptr() + 4 * col0, Hg) for d in range(2)] for l in range(2)])

"""Train-mode parity of the CPU oracle with the UNMODIFIED reference, dropout included.

torch.nn.GRU draws its inter-layer dropout masks inside ATen, out of reach of a Python patch -- but on the CPU they come
from the default generator in a fixed order, so they can be REPLAYED: save the generator state, run the reference
MeasureVAE in train mode, restore the state and draw the same Bernoulli / normal tensors in the reference's order
(MeasureVAE/measure_vae.py:104-131: encoder GRU -> z_dist.rsample() -> prior_dist.sample() -> decoder;
MeasureVAE/decoder.py:392-529: beat GRU, then 24 single-step tick-GRU calls, each with its own mask).  ATen applies the
mask to the time-major layer output, hence the (T, B, *) draw shapes.  Fed those masks, the oracle must reproduce the
reference's outputs and gradients to fp32 round-off -- which pins WHERE the oracle (and therefore the CUDA path, which is
tested against the oracle with injected masks) applies dropout: on layer 0's output of each 2-layer GRU call, never on
the last layer, with an independent mask per tick call."""
import random

import pytest
import torch

from oracle import inpaintnet_oracle as O
from oracle.ref_import import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not available")


@pytest.mark.parametrize("teacher_forced", [True, False])
@pytest.mark.parametrize("seed", [5, 6])
def test_mvae_train_mode_dropout_replayed_from_the_reference_generator(seed, teacher_forced):
    from oracle.ref_import import load_reference, FakeDataset
    R = load_reference()
    torch.manual_seed(seed)
    random.seed(seed)
    V, H, Z, B, p = 23, 24, 12, 4, 0.5
    m = R.MeasureVAE(FakeDataset(V), encoder_hidden_size=H, decoder_hidden_size=H, latent_space_dim=Z)
    with torch.no_grad():
        m.decoder.b_0.normal_()
        m.decoder.x_0.normal_()
    m.train()
    assert m.encoder.dropout == p and m.decoder.dropout == p
    m.decoder.teacher_forcing_prob = 2.0 if teacher_forced else -1.0
    tokens = torch.randint(0, V, (B, 24))
    torch.manual_seed(1000 + seed)
    state = torch.get_rng_state()
    m.zero_grad()
    w, s, zd, _, z_tilde, _ = m(tokens, train=True)
    loss = torch.nn.functional.cross_entropy(w.reshape(-1, V), tokens.reshape(-1)) + 0.01 * (zd.loc ** 2).mean()
    loss.backward()
    ref_grads = {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}
    after = torch.get_rng_state()

    # ---- replay the reference's draws
    torch.set_rng_state(state)
    keep = lambda shape: torch.empty(shape).bernoulli_(1 - p)
    enc = keep((24, B, 2 * H)).transpose(0, 1).contiguous()          # encoder GRU, layer 0 -> 1 (time-major in ATen)
    eps = torch.empty(B, Z).normal_()                                 # z_dist.rsample()
    torch.empty(B, Z).normal_()                                       # prior_dist.sample()
    beat = keep((4, B, H)).transpose(0, 1).contiguous()               # beat GRU, layer 0 -> 1
    tick = torch.cat([keep((1, B, H)).transpose(0, 1) for _ in range(24)], 1).contiguous()   # one mask per tick call
    assert torch.equal(torch.get_rng_state(), after), "the reference consumed the generator in a different order"

    sd = {k: v.detach().clone().requires_grad_() for k, v in m.state_dict().items()}
    mu, ls = O.encoder_forward(sd, tokens, 2, [enc], p)
    z = mu + torch.exp(ls) * eps
    w2, s2 = O.decoder_forward(sd, z, tokens if teacher_forced else None, teacher_forced, 2, [beat], tick, p)
    assert torch.allclose(mu, zd.loc, atol=2e-6) and torch.allclose(z, z_tilde, atol=2e-6)
    assert torch.equal(s2, s)
    assert torch.allclose(w2, w, atol=5e-6, rtol=1e-5)
    loss2 = torch.nn.functional.cross_entropy(w2.reshape(-1, V), tokens.reshape(-1)) + 0.01 * (mu ** 2).mean()
    loss2.backward()
    for k, g in ref_grads.items():
        assert sd[k].grad is not None, k
        assert torch.allclose(sd[k].grad, g, atol=2e-6, rtol=1e-4), (k, (sd[k].grad - g).abs().max().item())


def test_latent_rnn_train_mode_dropout_replayed_from_the_reference_generator():
    """LatentRNN (auto_reg=False) in train mode: `model.train()` recurses into the frozen VAE (utils/trainer.py:78), so
    dropout is active in its encoder and decoder as well as in the context / generation GRUs.  Draw order in the
    reference (LatentRNN/latent_rnn.py:131-140, 228-240): per get_z_seq call (past, future, target) the encoder GRU mask
    over all B*n measures then the rsample; the two context GRUs; the generation GRU; per gap measure the decoder's beat
    mask and its 24 tick masks."""
    from oracle.ref_import import load_reference, FakeDataset
    R = load_reference()
    torch.manual_seed(9)
    random.seed(9)
    V, H, Z, Hc, B, p = 23, 24, 12, 16, 3, 0.5
    n_p, n_t, n_f = 3, 2, 4
    ds = FakeDataset(V)
    vae = R.MeasureVAE(ds, encoder_hidden_size=H, decoder_hidden_size=H, latent_space_dim=Z)
    with torch.no_grad():
        vae.decoder.b_0.normal_()
        vae.decoder.x_0.normal_()
    m = R.LatentRNN(ds, vae, 2, Hc, p, torch.nn.GRU, auto_reg=False, teacher_forcing=True)
    m.train()
    assert vae.training and vae.encoder.dropout == p
    score = torch.randint(0, V, (B, n_p + n_t + n_f, 24))
    past, target, future = (x.contiguous() for x in (score[:, :n_p], score[:, n_p:n_p + n_t], score[:, n_p + n_t:]))
    torch.manual_seed(2024)
    state = torch.get_rng_state()
    w, s, gz = m(past, future, target, n_t, train=True)
    after = torch.get_rng_state()

    torch.set_rng_state(state)
    keep = lambda shape: torch.empty(shape).bernoulli_(1 - p)
    enc, eps = {}, {}
    for name, n in (("past", n_p), ("future", n_f), ("target", n_t)):
        enc[name] = keep((24, B * n, 2 * H)).transpose(0, 1).contiguous()      # rows b*n + m, as view(-1, 24) orders them
        eps[name] = torch.empty(B * n, Z).normal_().view(B, n, Z)
    ctx = {name: [keep((n, B, 2 * Hc)).transpose(0, 1).contiguous()] for name, n in (("past", n_p), ("future", n_f))}
    gen = [keep((n_t, B, 4 * Hc)).transpose(0, 1).contiguous()]                 # generation GRU: hidden 2*Hc, bidirectional
    dec = []
    for _ in range(n_t):
        beat = keep((4, B, H)).transpose(0, 1).contiguous()
        tick = torch.cat([keep((1, B, H)).transpose(0, 1) for _ in range(24)], 1).contiguous()
        dec.append((beat, tick))
    assert torch.equal(torch.get_rng_state(), after), "the reference consumed the generator in a different order"

    sd = {k: v.detach() for k, v in m.state_dict().items()}
    w2, s2, z2 = O.latent_rnn_forward(sd, past, future, target, n_t, eps["past"], eps["future"], ctx_keep_masks=ctx,
                                      gen_keep_masks=gen, dropout_p=p,
                                      vae_dropout=dict(enc_past=enc["past"], enc_future=enc["future"], dec=dec), vae_dropout_p=p)
    assert torch.allclose(z2, gz, atol=5e-6, rtol=1e-5)
    assert torch.equal(s2, s)
    assert torch.allclose(w2.reshape(w.shape), w, atol=1e-5, rtol=1e-5)


def test_arnn_train_mode_input_dropout_replayed_from_the_reference_generator():
    """AnticipationRNN, teacher forced, train mode: the only dropout the reference applies is the whole-timestep input
    dropout (`nn.Dropout2d` on x[:, :, :, None], arnn_model.py:142,437-442; `lstm_with_activations` is never given its
    dropout_layer, and the single-layer nn.LSTMs ignore their own `dropout`).  Its noise tensor is (B, T, 1, 1)."""
    from oracle.ref_import import load_reference, FakeDataset
    R = load_reference()
    random.seed(21)
    torch.manual_seed(21)
    V, B, T, p = 20, 3, 4 * 24, 0.2
    m = R.ConstraintModelGaussianReg(
        dataset=FakeDataset(V), note_embedding_dim=10, metadata_embedding_dim=2, num_lstm_constraints_units=32,
        num_lstm_generation_units=32, linear_hidden_size=32, num_layers=2, dropout_input_prob=p, dropout_prob=0.2,
        unary_constraint=True, teacher_forcing=True)
    m.train()
    score = torch.randint(0, V, (B, 1, T))
    t = torch.arange(T)
    md = torch.stack([(t // 6) % 4 == 0, t % 6, torch.zeros_like(t)], 1).long().view(1, 1, T, 3).expand(B, 1, T, 3).contiguous()
    cl = torch.ones(B, 1, T).long()
    cl[:, :, 24:48] = 0
    torch.manual_seed(77)
    state = torch.get_rng_state()
    weights, _ = m._forward_tf(score, md, cl)
    after = torch.get_rng_state()
    torch.set_rng_state(state)
    keep = torch.empty(B, T, 1, 1).bernoulli_(1 - p)[:, :, 0, 0]
    assert torch.equal(torch.get_rng_state(), after), "the reference consumed the generator in a different order"
    assert 0 < keep.sum() < keep.numel()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    logits = O.arnn_forward_tf(sd, score, md, cl, 2, keep_input_steps=keep, dropout_input_p=p)
    assert torch.allclose(logits, weights[0], atol=5e-6, rtol=1e-5)

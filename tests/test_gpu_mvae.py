"""GPU parity of the MeasureVAE hot path (encoder, hierarchical decoder, fused loss, backward, fused
Adam) against the golden vectors of the unmodified reference and against the CPU oracle.

Tolerances (BASELINE.json north_star): logits / latents within 1e-3 relative in fp32 mode and 2e-2 in
bf16 mode; argmax decodes bit-exact in fp32 mode (on rows with a strict top-1 margin)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from inpaintnet_b200 import engine, functional as Fn
from inpaintnet_b200.measure_vae import MeasureVAE
from inpaintnet_b200.trainer import VAETrainer
from inpaintnet_b200.data import SyntheticFolkDataset
from oracle import inpaintnet_oracle as O
from tests.golden import recipe

G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def load(name):
    return torch.load(os.path.join(G, name + ".pt"), weights_only=False)


def build(fx, prec):
    sd = fx["state_dict"] if "state_dict" in fx else recipe.make_state_dict(
        recipe.mvae_spec(fx["V"], 10, fx["H"], fx["Z"]), fx["seed"])
    ds = SyntheticFolkDataset(num_notes=fx["V"])
    m = MeasureVAE(ds, encoder_hidden_size=fx["H"], decoder_hidden_size=fx["H"], latent_space_dim=fx["Z"])
    m.load_state_dict(sd)
    m.to(DEV)
    m.set_precision(prec)
    return m, sd


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


CASES = [("mvae_h32", "fp32"), ("mvae_h64", "fp32"), ("mvae_default", "fp32"), ("mvae_h64", "bf16"),
         ("mvae_default", "bf16")]


@pytest.mark.parametrize("name,prec", CASES)
@pytest.mark.parametrize("mode", ["tf", "argmax"])
def test_forward_backward_vs_reference_golden(name, prec, mode):
    fx = load(name)
    g = fx[mode]
    model, sd = build(fx, prec)
    model.eval()  # dropout off, as in the golden run; TF chosen through teacher_forcing_prob
    model.decoder.teacher_forcing_prob = 2.0 if mode == "tf" else -1.0
    tokens = fx["tokens"].to(DEV)
    model.zero_grad()
    with engine.inject_noise(eps=[fx["eps"]]):
        weights, samples, z_dist, prior, z_tilde, _ = model(tokens, train=True)
    tol = 1e-3 if prec == "fp32" else 2e-2
    assert rel_err(z_dist.loc.cpu(), g["mu"]) < tol
    assert rel_err(z_dist.log_std.cpu(), g["log_std"]) < tol
    assert rel_err(z_tilde.detach().cpu(), g["z"]) < tol
    s_lib = samples.cpu()
    if mode == "tf":
        assert torch.equal(s_lib, g["samples"])
    elif prec == "fp32":
        strict = (g["margin"] > 1e-4).unsqueeze(1)
        assert bool(((s_lib == g["samples"]) | ~strict).all()), "argmax decode differs on a strict-margin row"
        assert bool(strict.all()) and torch.equal(s_lib, g["samples"])
    same_path = torch.equal(s_lib, g["samples"])
    if same_path:
        assert rel_err(weights.detach().cpu(), g["weights"]) < tol
    else:
        # bf16 argmax: the token feedback may legitimately leave the reference's path -- but only AT a near-tie of the
        # reference's own logits.  Per measure: up to and including the first differing tick the logits must match
        # within the tolerance (same inputs so far), and at that tick the reference's top-1 margin must be small.
        gs, gw, w_lib = g["samples"][:, 0], g["weights"], weights.detach().cpu()
        diff = s_lib[:, 0] != gs
        n_div = 0
        for b in range(gs.shape[0]):
            if not bool(diff[b].any()):
                assert rel_err(w_lib[b], gw[b]) < tol
                continue
            n_div += 1
            t0 = int(diff[b].float().argmax())
            assert rel_err(w_lib[b, :t0 + 1], gw[b, :t0 + 1]) < tol, (b, t0)
            top2 = gw[b, t0].topk(2).values
            assert float(top2[0] - top2[1]) < 5e-2 * float(gw[b].abs().max().clamp_min(1e-6)), (b, t0, top2)
        agree = (s_lib == g["samples"]).float().mean().item()
        print(f"bf16 argmax [{name}]: {n_div}/{gs.shape[0]} measures leave the reference token path at a near-tie; "
              f"token agreement {agree:.3f}")
        assert agree > 0.9, agree   # measured: 0.986 (mvae_h64: 1 of 6 measures diverges, at a near-tie); mvae_default stays on the path
        return
    loss, acc = Fn.fused_ce_kl(weights, tokens, z_dist.loc, z_dist.log_std, beta=0.001)
    assert abs(loss.item() - g["loss"]) < (2e-4 if prec == "fp32" else 2e-2)
    assert abs(acc.item() - g["acc"]) < (1e-6 if prec == "fp32" else 0.05)
    loss.backward()
    torch.cuda.synchronize()
    bad = []
    for k, gg in g["grads"].items():
        mine = dict(model.named_parameters())[k].grad
        assert mine is not None, k
        mine = mine.detach().float().cpu()
        if isinstance(gg, dict):
            ref_norm = gg["norm"]
            err = abs(mine.norm().item() - ref_norm) / max(ref_norm, 1e-8)
            head_err = (mine.reshape(-1)[:32] - gg["head"]).abs().max().item() / max(gg["head"].abs().max().item(), 1e-8)
            if err > (2e-3 if prec == "fp32" else 0.1) or head_err > (5e-3 if prec == "fp32" else 0.25):
                bad.append((k, err, head_err))
        else:
            err = (mine - gg).abs().max().item() / max(gg.abs().max().item(), 1e-8)
            if err > (2e-3 if prec == "fp32" else 0.15):
                bad.append((k, err))
    assert not bad, bad


def _lib_masks(B, H, enc, beat, tick):
    """oracle-layout keep masks -> library layouts (time-major / decoder row order)"""
    e = enc.transpose(0, 1).reshape(24 * B, 2 * H)
    b = beat.transpose(0, 1).reshape(4 * B, H)
    t = tick.view(B, 4, 6, H).permute(2, 1, 0, 3).reshape(24 * B, H)
    return [e.to(torch.uint8), b.to(torch.uint8), t.to(torch.uint8)]


@pytest.mark.parametrize("mode", ["tf", "argmax"])
def test_train_mode_dropout_vs_oracle_fp32(mode):
    fx = load("mvae_h32")
    model, sd = build(fx, "fp32")
    model.train()
    model.decoder.teacher_forcing_prob = 2.0 if mode == "tf" else -1.0
    B, H = fx["B"], fx["H"]
    g = torch.Generator().manual_seed(11)
    enc = (torch.rand(B, 24, 2 * H, generator=g) > 0.5).float()
    beat = (torch.rand(B, 4, H, generator=g) > 0.5).float()
    tick = (torch.rand(B, 24, H, generator=g) > 0.5).float()
    sdr = {k: v.clone().requires_grad_() for k, v in sd.items()}
    w_ref, s_ref, mu_ref, ls_ref, z_ref = O.mvae_forward(sdr, fx["tokens"], fx["eps"], teacher_forced=(mode == "tf"),
                                                         train_dropout=dict(enc=[enc], beat=[beat], tick=tick))
    O.mvae_loss(w_ref, fx["tokens"], mu_ref, ls_ref).backward()
    tokens = fx["tokens"].to(DEV)
    model.zero_grad()
    with engine.inject_noise(masks=_lib_masks(B, H, enc, beat, tick), eps=[fx["eps"]]):
        weights, samples, z_dist, *_ = model(tokens, train=True)
    assert torch.equal(samples.cpu(), s_ref)
    assert rel_err(weights.detach().cpu(), w_ref.detach()) < 1e-3
    loss, acc = Fn.fused_ce_kl(weights, tokens, z_dist.loc, z_dist.log_std, beta=0.001)
    loss.backward()
    torch.cuda.synchronize()
    for k, p in model.named_parameters():
        ref = sdr[k].grad
        err = (p.grad.cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-8)
        assert err < 2e-3, (k, err)


def test_trainer_step_fused_adam_matches_oracle():
    """One full train step through the reference-shaped Trainer API == oracle forward/backward + Adam."""
    fx = load("mvae_h32")
    model, sd = build(fx, "fp32")
    ds = SyntheticFolkDataset(num_notes=fx["V"])
    trainer = VAETrainer(ds, model, lr=1e-3)
    model.eval()
    model.decoder.teacher_forcing_prob = 2.0
    tokens = fx["tokens"].to(DEV)
    sdr = {k: v.clone().requires_grad_() for k, v in sd.items()}
    w, s, mu, ls, z = O.mvae_forward(sdr, fx["tokens"], fx["eps"], teacher_forced=True)
    O.mvae_loss(w, fx["tokens"], mu, ls).backward()
    trainer.zero_grad()
    with engine.inject_noise(eps=[fx["eps"]]):
        loss, acc = trainer.loss_and_acc_for_batch(tokens, 0, train=True)
    loss.backward()
    trainer.step()
    trainer.check_device_flags()
    torch.cuda.synchronize()
    new = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    for k in sd:
        p = sd[k].clone()
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        O.adam_step(p, sdr[k].grad, m, v, 1, lr=1e-3)
        assert torch.allclose(new[k], p, atol=2e-5, rtol=1e-4), k


def test_state_dict_layout_roundtrip():
    fx = load("mvae_h32")
    model, sd = build(fx, "fp32")
    out = model.state_dict()
    assert list(out.keys()) == list(recipe.mvae_spec(fx["V"], 10, fx["H"], fx["Z"]).keys()) or set(out) == set(sd)
    for k, v in sd.items():
        assert out[k].dtype == torch.float32 and tuple(out[k].shape) == tuple(v.shape)
        assert torch.equal(out[k].cpu(), v)


def test_token_range_guard():
    fx = load("mvae_h32")
    model, _ = build(fx, "fp32")
    ds = SyntheticFolkDataset(num_notes=fx["V"])
    trainer = VAETrainer(ds, model)
    bad = fx["tokens"].clone()
    bad[0, 3] = fx["V"] + 5
    with torch.no_grad():
        model(bad.to(DEV), train=False)
    with pytest.raises(ValueError):
        trainer.check_device_flags()


@pytest.mark.parametrize("prec,B,H", [("fp32", 12, 32), ("bf16", 512, 64), ("bf16", 1024, 512)])
@pytest.mark.parametrize("mode", ["tf", "argmax"])
def test_pipelined_microbatches_match_single_stream_step(prec, B, H, mode):
    """VAETrainer.microbatches = 2 (two half batches on two streams, one coin, gradients accumulated in the shared
    arena, weight-gradient GEMMs on per-stream side streams) must take the same training step as the whole batch
    on one stream: same loss, same gradients, same parameters after Adam -- twice in a row (cache / stream reuse)."""
    V, Z = 47, 32
    fx = dict(V=V, H=H, Z=Z, seed=1234)
    g = torch.Generator().manual_seed(5)
    tokens = [torch.randint(0, V, (B, 24), generator=g).to(DEV) for _ in range(2)]
    eps = [torch.randn(B, Z, generator=g) for _ in range(2)]
    out = {}
    for n in (1, 2):
        m, _ = build(fx, prec)
        m.eval()                                     # dropout off (masks would be drawn per micro-batch)
        m.decoder.teacher_forcing_prob = 2.0 if mode == "tf" else -1.0
        tr = VAETrainer(SyntheticFolkDataset(num_notes=V), m, lr=1e-3)
        tr.microbatches = n
        res = []
        for s in range(2):
            tr.zero_grad()
            with engine.inject_noise(eps=list(eps[s].chunk(n))):
                loss, acc = tr.loss_and_acc_for_batch(tokens[s], 0, train=True)
            loss.backward()
            torch.cuda.synchronize()
            grads = {k: p.grad.detach().float().clone() for k, p in m.named_parameters()}
            tr.step()
            torch.cuda.synchronize()
            res.append((loss.item(), acc.item(), grads, {k: p.detach().float().clone() for k, p in m.named_parameters()}))
        out[n] = res
    for s in range(2):
        l1, a1, g1, p1 = out[1][s]
        l2, a2, g2, p2 = out[2][s]
        if mode == "argmax" and prec == "bf16" and abs(l1 - l2) > 2e-3:
            pytest.skip("argmax decode diverged between the two runs (bf16 tie)")   # deterministic kernels: not expected
        assert abs(l1 - l2) < (1e-5 if prec == "fp32" else 2e-3), (s, l1, l2)
        assert abs(a1 - a2) < 1e-6 + (0 if prec == "fp32" else 1e-3)
        for k in g1:   # bf16: activations of a micro-batch are rounded like those of the full batch, sums differ in order
            err = (g1[k] - g2[k]).norm().item() / max(g1[k].norm().item(), 1e-8)
            # second step: the parameters already differ by the first step's rounding (bf16: small noisy gradients)
            assert err < (2e-4 if prec == "fp32" else (3e-2 if s == 0 else 0.15)), (s, k, err)
        if s == 0:
            for k in p1:
                assert (p1[k] - p2[k]).abs().mean().item() < (1e-6 if prec == "fp32" else 2e-4), (s, k)


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_full_size_batch_against_oracle_on_row_subsets(prec):
    """BASELINE.json configs[1] size (4096 measures, reference default hyper-parameters, V=64): measures are
    independent, so rows taken from different 128-row tiles of the full-size run must match the CPU oracle run on
    just those rows (logits / latents: 1e-3 relative in fp32 mode, 2e-2 in bf16 mode), teacher forced and argmax."""
    V, H, Z, B = 64, 512, 256, 4096
    fx = dict(V=V, H=H, Z=Z, seed=4321)
    m, sd = build(fx, prec)
    m.eval()
    g = torch.Generator().manual_seed(17)
    tokens = torch.randint(0, V, (B, 24), generator=g)
    eps = torch.randn(B, Z, generator=g)
    rows = torch.cat([torch.arange(a, a + 8) for a in (0, 120, 1021, 2048, 3333, 4088)])
    tol = 1e-3 if prec == "fp32" else 2e-2
    for mode in ("tf", "argmax"):
        m.decoder.teacher_forcing_prob = 2.0 if mode == "tf" else -1.0
        with torch.no_grad(), engine.inject_noise(eps=[eps]):
            w, s, zd, _, z, _ = m(tokens.to(DEV), train=True)
        w_ref, s_ref, mu_ref, ls_ref, _ = O.mvae_forward(sd, tokens[rows], eps[rows], teacher_forced=(mode == "tf"))
        assert rel_err(zd.loc.cpu()[rows], mu_ref) < tol
        assert rel_err(zd.log_std.cpu()[rows], ls_ref) < tol
        if mode == "tf":
            assert rel_err(w.cpu()[rows], w_ref) < tol
        else:   # free-running decode: compare every row up to its first differing token (all of it in fp32 mode)
            same = (s.cpu()[rows, 0] == s_ref[:, 0])
            n_ok = same.long().cumprod(1).sum(1)               # ticks before the first differing argmax token
            top2 = w_ref.topk(2, dim=2).values
            strict = (top2[..., 0] - top2[..., 1]) > (1e-4 if prec == "fp32" else 5e-2)
            for r in range(len(rows)):
                k = int(n_ok[r])
                if k < 24:
                    assert not bool(strict[r, k]), (mode, int(rows[r]), k)   # only a near-tie may flip a token
                k = min(k + 1, 24)   # the logits of the first differing tick still come from identical inputs
                assert rel_err(w.cpu()[rows[r], :k], w_ref[r, :k]) < tol, (mode, int(rows[r]))

"""GPU parity of the LatentRNN path (context GRUs, generation GRU, batched argmax decode, backward through
the frozen decoder) against golden vectors of the unmodified reference."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from inpaintnet_b200 import engine, functional as Fn
from inpaintnet_b200.data import SyntheticFolkDataset
from inpaintnet_b200.latent_rnn import LatentRNN
from inpaintnet_b200.measure_vae import MeasureVAE
from tests.golden import recipe

G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def build(fx, prec):
    if "state_dict" in fx:
        sd = fx["state_dict"]
    else:
        sd = recipe.make_state_dict(recipe.latent_rnn_spec(fx["Z"], fx["Hc"]), fx["seed"] + 10)
        sd.update({"vae_model." + k: v for k, v in recipe.make_state_dict(
            recipe.mvae_spec(fx["V"], 10, fx["H"], fx["Z"]), fx["seed"]).items()})
    ds = SyntheticFolkDataset(num_notes=fx["V"])
    vae = MeasureVAE(ds, encoder_hidden_size=fx["H"], decoder_hidden_size=fx["H"], latent_space_dim=fx["Z"])
    m = LatentRNN(ds, vae, 2, fx["Hc"], 0.5, torch.nn.GRU, auto_reg=False)
    m.load_state_dict(sd)
    m.to(DEV)
    m.set_precision(prec)
    return m


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


@pytest.mark.parametrize("name,prec", [("latent_h32", "fp32"), ("latent_default", "fp32"), ("latent_default", "bf16")])
def test_latent_rnn_vs_reference_golden(name, prec):
    fx = torch.load(os.path.join(G, name + ".pt"), weights_only=False)
    m = build(fx, prec)
    m.eval()
    B, n_p, Z = fx["eps_past"].shape
    n_f = fx["eps_future"].shape[1]
    n_t = fx["target"].shape[1]
    eps = [fx["eps_past"].transpose(0, 1).reshape(n_p * B, Z), fx["eps_future"].transpose(0, 1).reshape(n_f * B, Z)]
    m.zero_grad()
    with engine.inject_noise(eps=eps):
        weights, samples, gen_z = m(fx["past"].to(DEV), fx["future"].to(DEV), fx["target"].to(DEV), n_t, train=True)
    tol = 1e-3 if prec == "fp32" else 2e-2
    assert weights.shape == fx["weights"].shape and samples.shape == fx["samples"].shape
    assert rel_err(gen_z.detach().cpu(), fx["gen_z"]) < tol
    strict = (fx["margin"] > (1e-4 if prec == "fp32" else 5e-2)).reshape(B, -1)
    same = samples.cpu()[:, 0] == fx["samples"][:, 0]
    if prec == "fp32":
        assert bool((same | ~strict).all()), "argmax decode differs on a strict-margin row"
    if not bool(same.all()):
        assert same.float().mean().item() > 0.5
        return
    assert rel_err(weights.detach().cpu(), fx["weights"]) < tol
    loss, acc = Fn.fused_ce_kl(weights, fx["target"].to(DEV))
    assert abs(loss.item() - fx["loss"]) < (2e-4 if prec == "fp32" else 2e-2)
    loss.backward()
    torch.cuda.synchronize()
    bad = []
    params = dict(m.named_parameters())
    for k, gg in fx["grads"].items():
        mine = params[k].grad
        assert mine is not None, k
        mine = mine.detach().float().cpu()
        if isinstance(gg, dict):
            err = abs(mine.norm().item() - gg["norm"]) / max(gg["norm"], 1e-8)
            if err > (3e-3 if prec == "fp32" else 0.1):
                bad.append((k, err))
        else:
            err = (mine - gg).abs().max().item() / max(gg.abs().max().item(), 1e-8)
            if err > (3e-3 if prec == "fp32" else 0.15):
                bad.append((k, err))
    assert not bad, bad
    for k, p in params.items():
        if k.startswith("vae_model."):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k   # frozen VAE receives no gradient


def test_inference_no_grad_matches_grad_path():
    fx = torch.load(os.path.join(G, "latent_h32.pt"), weights_only=False)
    m = build(fx, "fp32")
    m.eval()
    B, n_p, Z = fx["eps_past"].shape
    n_f = fx["eps_future"].shape[1]
    eps = [fx["eps_past"].transpose(0, 1).reshape(n_p * B, Z), fx["eps_future"].transpose(0, 1).reshape(n_f * B, Z)]
    with torch.no_grad(), engine.inject_noise(eps=eps):
        w, s, z = m(fx["past"].to(DEV), fx["future"].to(DEV), fx["target"].to(DEV), fx["target"].shape[1], train=False)
    assert torch.equal(s.cpu(), fx["samples"])
    assert rel_err(w.cpu(), fx["weights"]) < 1e-3

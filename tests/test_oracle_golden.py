"""Pins the CPU oracle (oracle/inpaintnet_oracle.py) against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""
import os

import pytest
import torch

from oracle import inpaintnet_oracle as O
from tests.golden import recipe

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(G, name + ".pt"), weights_only=False)


def mvae_sd(fx):
    if "state_dict" in fx:
        return fx["state_dict"]
    return recipe.make_state_dict(recipe.mvae_spec(fx["V"], 10, fx["H"], fx["Z"]), fx["seed"])


@pytest.mark.parametrize("name", ["mvae_h32", "mvae_h64", "mvae_default"])
@pytest.mark.parametrize("mode", ["tf", "argmax"])
def test_mvae_forward_loss_grads(name, mode):
    fx = load(name)
    sd = {k: v.clone().requires_grad_() for k, v in mvae_sd(fx).items()}
    g = fx[mode]
    w, s, mu, ls, z = O.mvae_forward(sd, fx["tokens"], fx["eps"], teacher_forced=(mode == "tf"))
    assert torch.allclose(mu, g["mu"], atol=2e-6, rtol=1e-5)
    assert torch.allclose(ls, g["log_std"], atol=2e-6, rtol=1e-5)
    assert torch.allclose(z, g["z"], atol=2e-6, rtol=1e-5)
    # argmax parity is defined on strict-margin rows; golden rows all have margin > 1e-4
    strict = g["margin"] > 1e-5
    assert strict.all()
    assert torch.equal(s, g["samples"])
    assert torch.allclose(w, g["weights"], atol=5e-6, rtol=1e-5)
    loss = O.mvae_loss(w, fx["tokens"], mu, ls)
    assert abs(loss.item() - g["loss"]) < 1e-5
    assert abs(O.mean_accuracy(w, fx["tokens"]).item() - g["acc"]) < 1e-6
    loss.backward()
    for k, gg in g["grads"].items():
        mine = sd[k].grad
        if isinstance(gg, dict):
            assert abs(mine.norm().item() - gg["norm"]) <= 1e-4 * max(gg["norm"], 1e-6) + 1e-7, k
            assert torch.allclose(mine.reshape(-1)[:32], gg["head"], atol=1e-6, rtol=1e-3), k
        else:
            assert torch.allclose(mine, gg, atol=1e-6, rtol=1e-3), k


def test_adam_matches_torch():
    fx = load("adam")
    p = fx["p0"].clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for i, (gr, pref) in enumerate(zip(fx["grads"], fx["ps"])):
        O.adam_step(p, gr, m, v, i + 1)
        assert torch.allclose(p, pref, atol=1e-7, rtol=1e-6)


@pytest.mark.parametrize("name", ["latent_h32", "latent_default", "latent_abl_future_h32"])
def test_latent_rnn_forward(name):
    fx = load(name)
    if "state_dict" in fx:
        sd = fx["state_dict"]
    else:
        sd = recipe.make_state_dict(recipe.latent_rnn_spec(fx["Z"], fx["Hc"]), fx["seed"] + 10)
        sd.update({"vae_model." + k: v for k, v in recipe.make_state_dict(
            recipe.mvae_spec(fx["V"], 10, fx["H"], fx["Z"]), fx["seed"]).items()})
    sd = {k: v.clone().requires_grad_(not k.startswith("vae_model.")) for k, v in sd.items()}
    n_gen = fx["target"].shape[1]
    w, s, z = O.latent_rnn_forward(sd, fx["past"], fx["future"], fx["target"], n_gen,
                                   fx["eps_past"], fx["eps_future"], only=fx.get("abl_type"))
    assert torch.allclose(z, fx["gen_z"], atol=5e-6, rtol=1e-5)
    strict = (fx["margin"] > 1e-4).reshape(s.shape[0], -1)
    same = (s[:, 0] == fx["samples"][:, 0])
    assert bool((same | ~strict).all())
    if bool(same.all()):
        assert torch.allclose(w, fx["weights"], atol=1e-5, rtol=1e-4)
        loss = O.mean_crossentropy_loss(w, fx["target"])
        assert abs(loss.item() - fx["loss"]) < 1e-5
        loss.backward()
        for k, gg in fx["grads"].items():
            mine = sd[k].grad
            if isinstance(gg, dict):
                assert abs(mine.norm().item() - gg["norm"]) <= 1e-3 * max(gg["norm"], 1e-6) + 1e-7, k
            else:
                assert torch.allclose(mine, gg, atol=1e-6, rtol=2e-3), k


@pytest.mark.parametrize("name", ["latent_ar_tf_h32", "latent_ar_notf_h32", "latent_abl_past_ar_tf_h32"])
def test_latent_rnn_autoregressive_forward(name):
    """auto_reg=True (the train_inpaintnet.py default), teacher-forced and free-running branches."""
    fx = load(name)
    sd = {k: v.clone().requires_grad_(not k.startswith("vae_model.")) for k, v in fx["state_dict"].items()}
    n_gen = fx["target"].shape[1]
    w, s, z = O.latent_rnn_forward_autoreg(sd, fx["past"], fx["future"], fx["target"], n_gen, fx["eps_past"],
                                           fx["eps_future"], fx["eps_target"], fx["eps_regen"], fx["teacher_forcing"],
                                           only=fx.get("abl_type"))
    assert torch.equal(s, fx["samples"])
    assert torch.allclose(z, fx["gen_z"], atol=5e-6, rtol=1e-5)
    assert torch.allclose(w, fx["weights"], atol=1e-5, rtol=1e-4)
    loss = O.mean_crossentropy_loss(w, fx["target"])
    assert abs(loss.item() - fx["loss"]) < 1e-5
    loss.backward()
    for k, gg in fx["grads"].items():
        assert torch.allclose(sd[k].grad, gg, atol=1e-6, rtol=2e-3), k


def test_arnn_teacher_forced_logits():
    fx = load("arnn_h32")
    logits = O.arnn_forward_tf(fx["state_dict"], fx["score"], fx["metadata"], fx["constraints_loc"])
    assert torch.allclose(logits, fx["logits"], atol=2e-6, rtol=1e-5)


def test_arnn_no_teacher_forcing_logits():
    fx = load("arnn_h32")
    logits, fed = O.arnn_forward_no_tf(fx["state_dict"], fx["score"], fx["metadata"], fx["constraints_loc"])
    # gen_chorale[:, 0, t+1] holds the token fed at tick t+1 (the last argmax is never fed)
    assert torch.equal(fed[1:], fx["gen_no_tf"][0, 0, 1:fed.numel()])
    assert torch.allclose(logits, fx["logits_no_tf"], atol=5e-6, rtol=1e-5)


def test_arnn_forward_inpaint():
    fx = load("arnn_inpaint_h32")
    logits, gen = O.arnn_forward_inpaint(fx["state_dict"], fx["score"], fx["metadata"], fx["constraints_loc"],
                                         fx["start"], fx["end"])
    assert torch.equal(gen, fx["gen"])
    assert torch.allclose(logits, fx["logits"], atol=2e-6, rtol=1e-5)

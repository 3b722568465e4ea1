"""Generates tests/golden/*.pt from the UNMODIFIED reference (needs /root/reference).

Run in the build container:   python tests/golden/make_golden.py
Every fixture holds inputs + outputs of the reference's own modules (fp32 CPU, eval-mode
dropout, injected eps); see SURVEY.md section 8(c) for why the noise is injected.
"""
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.ref_import import load_reference, FakeDataset  # noqa: E402
from tests.golden import recipe  # noqa: E402

R = load_reference()


class InjectedNoise:
    """Replaces Normal.rsample by loc + scale * eps (eps popped from a queue)."""

    def __init__(self, eps_list):
        self.eps = list(eps_list)

    def __enter__(self):
        self.orig = torch.distributions.Normal.rsample
        outer = self

        def rsample(dist, sample_shape=torch.Size()):
            e = outer.eps.pop(0)
            return dist.loc + dist.scale * e.view_as(dist.loc)

        torch.distributions.Normal.rsample = rsample
        return self

    def __exit__(self, *a):
        torch.distributions.Normal.rsample = self.orig


def grads_summary(model, full):
    out = {}
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach()
        out[n] = g.clone() if full else dict(norm=g.norm().item(), sum=g.sum().item(),
                                             head=g.reshape(-1)[:32].clone())
    return out


def mvae_case(name, V, H, Z, B, seed, store_weights, full_grads):
    ds = FakeDataset(V)
    model = R.MeasureVAE(ds, encoder_hidden_size=H, decoder_hidden_size=H, latent_space_dim=Z)
    sd = recipe.make_state_dict(recipe.mvae_spec(V, 10, H, Z), seed)
    model.load_state_dict(sd)
    model.eval()  # dropout off; TF selected through the `train` argument below
    tokens = recipe.make_tokens(B, 24, V, seed + 1)
    eps = recipe.make_normal((B, Z), seed + 2)
    fx = dict(V=V, H=H, Z=Z, B=B, seed=seed, tokens=tokens, eps=eps)
    if store_weights:
        fx["state_dict"] = sd
    for mode in ("tf", "argmax"):
        model.zero_grad()
        model.decoder.teacher_forcing_prob = 2.0 if mode == "tf" else -1.0
        with InjectedNoise([eps]):
            weights, samples, z_dist, prior, z_tilde, _ = model(tokens, train=True)
        ce = R.VAETrainer.mean_crossentropy_loss(weights=weights, targets=tokens)
        kld = R.VAETrainer.compute_kld_loss(z_dist, prior)
        loss = ce + kld
        acc = R.VAETrainer.mean_accuracy(weights=weights, targets=tokens)
        loss.backward()
        # margin of the top-1 logit per row (argmax parity is defined on strict-margin rows)
        top2 = weights.detach().topk(2, dim=2).values
        fx[mode] = dict(weights=weights.detach().clone(), samples=samples.clone(),
                        mu=z_dist.loc.detach().clone(), log_std=z_dist.scale.log().detach().clone(),
                        z=z_tilde.detach().clone(), ce=ce.item(), kld=kld.item(), loss=loss.item(),
                        acc=acc.item(), margin=(top2[..., 0] - top2[..., 1]).clone(),
                        grads=grads_summary(model, full_grads))
    # eval forward (train=False -> argmax decode)
    with InjectedNoise([eps]):
        w, s, zd, *_ = model(tokens, train=False)
    assert torch.equal(s, fx["argmax"]["samples"])
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "loss tf", fx["tf"]["loss"], "argmax", fx["argmax"]["loss"],
          "min margin", fx["argmax"]["margin"].min().item())


def adam_case():
    """3 torch.optim.Adam steps (utils/trainer.py:32-35 settings) on a small tensor."""
    g = torch.Generator().manual_seed(7)
    p = torch.randn(1000, generator=g).requires_grad_()
    opt = torch.optim.Adam([p], lr=1e-4)
    p0 = p.detach().clone()
    grads, ps = [], []
    for i in range(3):
        gr = torch.randn(1000, generator=g) * (10.0 ** (i - 1))
        p.grad = gr.clone()
        opt.step()
        grads.append(gr)
        ps.append(p.detach().clone())
    torch.save(dict(p0=p0, grads=grads, ps=ps), os.path.join(HERE, "adam.pt"))


def latent_case(name, V, H, Z, Hc, B, n_past, n_tgt, n_fut, seed, store_weights, auto_reg=False, teacher_forcing=False,
                abl_type=None):
    """auto_reg=True: latent_rnn.py:142-153,219-261; teacher_forcing picks the branch (the coin is forced through
    teacher_forcing_prob); eps_regen[i] is the rsample noise of the re-encode after gap measure i (no-TF branch)."""
    ds = FakeDataset(V)
    vae = R.MeasureVAE(ds, encoder_hidden_size=H, decoder_hidden_size=H, latent_space_dim=Z)
    vsd = recipe.make_state_dict(recipe.mvae_spec(V, 10, H, Z), seed)
    vae.load_state_dict(vsd)
    if abl_type is None:
        model = R.LatentRNN(ds, vae, 2, Hc, 0.5, torch.nn.GRU, auto_reg=auto_reg, teacher_forcing=auto_reg)
    else:   # latent_rnn_ablations.py: context from one side only, generation GRU hidden size = Hc
        import importlib
        abl = importlib.import_module("LatentRNN.latent_rnn_ablations")
        model = abl.LatentRNNAblations(ds, vae, 2, Hc, 0.5, torch.nn.GRU, auto_reg=auto_reg, teacher_forcing=auto_reg,
                                       type=abl_type)
    model.teacher_forcing_prob = 2.0 if teacher_forcing else -1.0
    lsd = recipe.make_state_dict(recipe.latent_rnn_spec(Z, Hc, auto_reg=auto_reg, ablation=abl_type is not None), seed + 10)
    sd = dict(lsd)
    sd.update({"vae_model." + k: v for k, v in vsd.items()})
    model.load_state_dict(sd)
    model.eval()
    n = n_past + n_tgt + n_fut
    score = recipe.make_tokens(B, 24 * n, V, seed + 1).view(B, n, 24)
    past, target, future = score[:, :n_past].contiguous(), score[:, n_past:n_past + n_tgt].contiguous(), \
        score[:, n_past + n_tgt:].contiguous()
    eps_p = recipe.make_normal((B, n_past, Z), seed + 2)
    eps_f = recipe.make_normal((B, n_fut, Z), seed + 3)
    eps_t = recipe.make_normal((B, n_tgt, Z), seed + 4)
    eps_r = recipe.make_normal((n_tgt, B, Z), seed + 5)
    model.zero_grad()
    with InjectedNoise([eps_p.reshape(-1, Z), eps_f.reshape(-1, Z), eps_t.reshape(-1, Z)] + [eps_r[i] for i in range(n_tgt)]):
        weights, samples, gen_z = model(past, future, target, n_tgt, train=True)
    loss = R.LatentRNNTrainer.mean_crossentropy_loss_alt(weights=weights, targets=target)
    acc = R.LatentRNNTrainer.mean_accuracy_alt(weights=weights, targets=target)
    loss.backward()
    top2 = weights.detach().topk(2, dim=3).values
    fx = dict(V=V, H=H, Z=Z, Hc=Hc, B=B, seed=seed, past=past, future=future, target=target,
              eps_past=eps_p, eps_future=eps_f, eps_target=eps_t, eps_regen=eps_r, auto_reg=auto_reg,
              teacher_forcing=teacher_forcing, abl_type=abl_type,
              weights=weights.detach().clone(), samples=samples.clone(), gen_z=gen_z.detach().clone(),
              loss=loss.item(), acc=acc.item(), margin=(top2[..., 0] - top2[..., 1]).clone(),
              grads=grads_summary(model, store_weights))
    if store_weights:
        fx["state_dict"] = sd
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "loss", fx["loss"], "min margin", fx["margin"].min().item())


def arnn_case(name, V, B, seed):
    ds = FakeDataset(V)
    random.seed(seed)
    torch.manual_seed(seed)
    model = R.ConstraintModelGaussianReg(
        dataset=ds, note_embedding_dim=10, metadata_embedding_dim=2, num_lstm_constraints_units=32,
        num_lstm_generation_units=32, linear_hidden_size=32, num_layers=2, dropout_input_prob=0.2,
        dropout_prob=0.2, unary_constraint=True, teacher_forcing=True)
    model.eval()
    T = 16 * 24
    score = recipe.make_tokens(B, T, V, seed + 1).view(B, 1, T)
    t = torch.arange(T)
    metadata = torch.stack([(t // 6) % 4 == 0, t % 6, torch.zeros_like(t)], 1).long()
    metadata = metadata.view(1, 1, T, 3).expand(B, 1, T, 3).contiguous()
    start, end = 5 * 24, 9 * 24
    cl = torch.ones(B, 1, T).long()
    cl[:, :, start:end] = 0
    weights, _ = model._forward_tf(score, metadata, cl)
    with torch.no_grad():
        w_notf, gen = model._forward_no_tf(score, metadata, cl)
    fx = dict(V=V, B=B, score=score, metadata=metadata, constraints_loc=cl,
              state_dict={k: v.clone() for k, v in model.state_dict().items()},
              logits=weights[0].detach().clone(), logits_no_tf=w_notf[0].detach().clone(), gen_no_tf=gen.clone())
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, weights[0].shape)


def arnn_inpaint_case(name, V, B, seed, start, end):
    """ConstraintModelGaussianReg.forward_inpaint (arnn_model.py:261-346): prefix scan + per-tick generation."""
    ds = FakeDataset(V)
    random.seed(seed)
    torch.manual_seed(seed)
    model = R.ConstraintModelGaussianReg(
        dataset=ds, note_embedding_dim=10, metadata_embedding_dim=2, num_lstm_constraints_units=32,
        num_lstm_generation_units=32, linear_hidden_size=32, num_layers=2, dropout_input_prob=0.2,
        dropout_prob=0.2, unary_constraint=True, teacher_forcing=True)
    model.eval()
    T = 8 * 24
    score = recipe.make_tokens(B, T, V, seed + 1).view(B, 1, T)
    t = torch.arange(T)
    metadata = torch.stack([(t // 6) % 4 == 0, t % 6, torch.zeros_like(t)], 1).long()
    metadata = metadata.view(1, 1, T, 3).expand(B, 1, T, 3).contiguous()
    cl = torch.ones(B, 1, T).long()
    cl[:, :, start:end] = 0
    with torch.no_grad():
        w, gen = model.forward_inpaint(score, metadata, cl, start, end)
    top2 = w[0][0].topk(2, dim=1).values      # the fed-back token is the argmax of batch element 0
    fx = dict(V=V, B=B, score=score, metadata=metadata, constraints_loc=cl, start=start, end=end,
              state_dict={k: v.clone() for k, v in model.state_dict().items()},
              logits=w[0].detach().clone(), gen=gen.clone(), margin=(top2[:, 0] - top2[:, 1]).clone())
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, w[0].shape, "min margin", fx["margin"].min().item())


if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 else ""   # optional name prefix: regenerate just those fixtures
    cases = [
        ("mvae_h32", lambda n: mvae_case(n, V=20, H=32, Z=16, B=5, seed=100, store_weights=True, full_grads=True)),
        ("mvae_h64", lambda n: mvae_case(n, V=47, H=64, Z=32, B=6, seed=200, store_weights=True, full_grads=True)),
        ("mvae_default", lambda n: mvae_case(n, V=64, H=512, Z=256, B=4, seed=300, store_weights=False, full_grads=False)),
        ("adam", lambda n: adam_case()),
        ("latent_h32", lambda n: latent_case(n, V=20, H=32, Z=16, Hc=32, B=3, n_past=3, n_tgt=2, n_fut=3, seed=400,
                                             store_weights=True)),
        ("latent_default", lambda n: latent_case(n, V=64, H=512, Z=256, Hc=512, B=2, n_past=6, n_tgt=4, n_fut=6, seed=500,
                                                 store_weights=False)),
        ("arnn_h32", lambda n: arnn_case(n, V=20, B=2, seed=600)),
        ("latent_ar_tf_h32", lambda n: latent_case(n, V=20, H=32, Z=16, Hc=32, B=3, n_past=3, n_tgt=3, n_fut=2, seed=700,
                                                   store_weights=True, auto_reg=True, teacher_forcing=True)),
        ("latent_ar_notf_h32", lambda n: latent_case(n, V=20, H=32, Z=16, Hc=32, B=3, n_past=2, n_tgt=3, n_fut=3, seed=828,
                                                     store_weights=True, auto_reg=True, teacher_forcing=False)),
        ("latent_abl_future_h32", lambda n: latent_case(n, V=20, H=32, Z=16, Hc=32, B=3, n_past=3, n_tgt=2, n_fut=3, seed=900,
                                                        store_weights=True, abl_type="future")),
        ("latent_abl_past_ar_tf_h32", lambda n: latent_case(n, V=20, H=32, Z=16, Hc=32, B=3, n_past=3, n_tgt=3, n_fut=2,
                                                            seed=1000, store_weights=True, auto_reg=True, teacher_forcing=True,
                                                            abl_type="past")),
        ("arnn_inpaint_h32", lambda n: arnn_inpaint_case(n, V=20, B=3, seed=1100, start=3 * 24, end=5 * 24)),
    ]
    for name, fn in cases:
        if name.startswith(only):
            torch.manual_seed(0)
            random.seed(0)
            fn(name)

"""GPU parity of the AnticipationRNN (LSTM) path against the golden vectors of the unmodified reference
(teacher-forced and no-teacher-forcing forward) and against the oracle's autograd (gradients)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from inpaintnet_b200.arnn import ConstraintModelGaussianReg
from inpaintnet_b200.data import SyntheticFolkDataset
from inpaintnet_b200 import functional as Fn
from oracle import inpaintnet_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def build(fx, prec):
    ds = SyntheticFolkDataset(num_notes=fx["V"])
    m = ConstraintModelGaussianReg(ds, note_embedding_dim=10, metadata_embedding_dim=2, num_lstm_constraints_units=32,
                                   num_lstm_generation_units=32, linear_hidden_size=32, num_layers=2, dropout_input_prob=0.2,
                                   dropout_prob=0.2, unary_constraint=True, teacher_forcing=True)
    m.load_state_dict(fx["state_dict"])
    m.to(DEV).set_precision(prec)
    m.eval()
    return m


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_teacher_forced_forward_and_gradients(prec):
    fx = torch.load(os.path.join(G, "arnn_h32.pt"), weights_only=False)
    m = build(fx, prec)
    m.teacher_forcing_prob = 2.0
    score, md, cl = fx["score"].to(DEV), fx["metadata"].to(DEV), fx["constraints_loc"].to(DEV)
    gap = (fx["constraints_loc"][0, 0] == 0).nonzero().squeeze()
    m.zero_grad()
    weights, _ = m(score, md, cl, train=True)
    tol = 1e-3 if prec == "fp32" else 3e-2
    assert rel_err(weights[0].detach().cpu(), fx["logits"][:, gap]) < tol
    targets = score[:, 0, gap.to(DEV)]
    loss, acc = Fn.fused_ce_kl(weights[0], targets)
    loss.backward()
    torch.cuda.synchronize()
    sd = {k: v.clone().requires_grad_() for k, v in fx["state_dict"].items()}
    logits = O.arnn_forward_tf(sd, fx["score"], fx["metadata"], fx["constraints_loc"])
    ref = O.mean_crossentropy_loss(logits[:, gap], fx["score"][:, 0, gap])
    ref.backward()
    assert abs(loss.item() - ref.item()) < (1e-4 if prec == "fp32" else 2e-2)
    bad = []
    for k, p in m.named_parameters():
        g_ref = sd[k].grad
        err = (p.grad.cpu() - g_ref).abs().max().item() / max(g_ref.abs().max().item(), 1e-8)
        if err > (3e-3 if prec == "fp32" else 0.2):
            bad.append((k, err))
    assert not bad, bad


def test_no_teacher_forcing_forward_fp32():
    fx = torch.load(os.path.join(G, "arnn_h32.pt"), weights_only=False)
    m = build(fx, "fp32")
    score, md, cl = fx["score"].to(DEV), fx["metadata"].to(DEV), fx["constraints_loc"].to(DEV)
    gap = (fx["constraints_loc"][0, 0] == 0).nonzero().squeeze()
    with torch.no_grad():
        weights, _ = m(score, md, cl, train=False)
    assert rel_err(weights[0].cpu(), fx["logits_no_tf"][:, gap]) < 1e-3


def test_no_teacher_forcing_backward_runs_and_matches_oracle():
    fx = torch.load(os.path.join(G, "arnn_h32.pt"), weights_only=False)
    m = build(fx, "fp32")
    m.teacher_forcing_prob = -1.0
    score, md, cl = fx["score"].to(DEV), fx["metadata"].to(DEV), fx["constraints_loc"].to(DEV)
    gap = (fx["constraints_loc"][0, 0] == 0).nonzero().squeeze()
    m.zero_grad()
    weights, _ = m(score, md, cl, train=True)
    loss, _ = Fn.fused_ce_kl(weights[0], score[:, 0, gap.to(DEV)])
    loss.backward()
    torch.cuda.synchronize()
    sd = {k: v.clone().requires_grad_() for k, v in fx["state_dict"].items()}
    logits, fed = O.arnn_forward_no_tf(sd, fx["score"], fx["metadata"], fx["constraints_loc"])
    O.mean_crossentropy_loss(logits[:, gap], fx["score"][:, 0, gap]).backward()
    for k, p in m.named_parameters():
        g_ref = sd[k].grad
        err = (p.grad.cpu() - g_ref).abs().max().item() / max(g_ref.abs().max().item(), 1e-8)
        assert err < 3e-3, (k, err)

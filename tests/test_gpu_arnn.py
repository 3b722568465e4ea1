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


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_forward_inpaint_vs_reference_golden(prec):
    """forward_inpaint (arnn_model.py:261-346): teacher-forced prefix scan, then per-tick generation through the gap."""
    fx = torch.load(os.path.join(G, "arnn_inpaint_h32.pt"), weights_only=False)
    m = build(fx, prec)
    score, md, cl = fx["score"].to(DEV), fx["metadata"].to(DEV), fx["constraints_loc"].to(DEV)
    weights, gen = m.forward_inpaint(score, md, cl, fx["start"], fx["end"])
    assert weights[0].shape == fx["logits"].shape and gen.shape == fx["gen"].shape
    if prec == "fp32":
        assert torch.equal(gen.cpu(), fx["gen"])                      # argmax tokens bit-exact (min margin 8e-3)
        assert rel_err(weights[0].cpu(), fx["logits"]) < 1e-3
    else:
        same = (gen.cpu() == fx["gen"])[0, 0, fx["start"]:fx["end"]]
        n_ok = int(same.long().cumprod(0).sum())                      # ticks before the first differing fed-back token
        assert n_ok >= 1
        assert rel_err(weights[0].cpu()[:, :n_ok], fx["logits"][:, :n_ok]) < 3e-2
    # outside the gap the score is returned untouched
    assert torch.equal(gen.cpu()[:, :, :fx["start"]], fx["score"][:, :, :fx["start"]])
    assert torch.equal(gen.cpu()[:, :, fx["end"]:], fx["score"][:, :, fx["end"]:])


def test_forward_inpaint_gap_at_start():
    """start_tick = 0: no prefix; the first input is the start symbol (id 0) as in _forward_no_tf."""
    fx = torch.load(os.path.join(G, "arnn_inpaint_h32.pt"), weights_only=False)
    m = build(fx, "fp32")
    cl = torch.ones_like(fx["constraints_loc"])
    cl[:, :, :24] = 0
    weights, gen = m.forward_inpaint(fx["score"].to(DEV), fx["metadata"].to(DEV), cl.to(DEV), 0, 24)
    logits, gen_ref = O.arnn_forward_inpaint(fx["state_dict"], fx["score"], fx["metadata"], cl, 0, 24)
    assert torch.equal(gen.cpu(), gen_ref)
    assert rel_err(weights[0].cpu(), logits) < 1e-3


def test_tester_inpainting_loss_matches_oracle():
    """AnticipationRNNTester.loss_and_acc_test (anticipation_rnn_tester.py:44-86): fixed gap = measures [8, 10)."""
    from inpaintnet_b200.tester import AnticipationRNNTester
    fx = torch.load(os.path.join(G, "arnn_inpaint_h32.pt"), weights_only=False)
    m = build(fx, "fp32")
    ds = SyntheticFolkDataset(num_notes=fx["V"], num_sequences=6, seed=3)
    tester = AnticipationRNNTester(ds, m)
    loader = ds.data_loaders(batch_size=3, split=(0.0, 0.0))[2]
    loss, acc = tester.loss_and_acc_test(loader)
    ref_loss, n = 0.0, 0
    for score, md in loader:
        cl, s0, s1 = tester.get_constraints_location(score.long(), is_stochastic=False)
        assert (s0, s1) == (8 * 24, 10 * 24)
        logits, _ = O.arnn_forward_inpaint(fx["state_dict"], score.long(), md.long(), cl, s0, s1)
        ref_loss += O.mean_crossentropy_loss(logits, score.long()[:, 0, s0:s1]).item()
        n += 1
    assert n == 2
    assert abs(float(loss) - ref_loss / n) < 1e-3
    assert 0.0 <= float(acc) <= 1.0

"""GPU checks of the HBM-bound helper kernels: Philox RNG, fused CE+KL, fused Adam, embeddings."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from inpaintnet_b200 import ops, functional as Fn
from inpaintnet_b200.ops import F32, BF16
from oracle import inpaintnet_oracle as O

DEV = "cuda"


def test_rng_mask_and_normal_statistics():
    n = 1 << 20
    m = torch.empty(n, dtype=torch.uint8, device=DEV)
    ops.rng_keep_mask(123, 0, n, 0.5, m.data_ptr())
    m2 = torch.empty(n, dtype=torch.uint8, device=DEV)
    ops.rng_keep_mask(123, 0, n, 0.5, m2.data_ptr())
    m3 = torch.empty(n, dtype=torch.uint8, device=DEV)
    ops.rng_keep_mask(123, n, n, 0.2, m3.data_ptr())
    e = torch.empty(n, device=DEV)
    ops.rng_normal(7, 0, n, e.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(m, m2)
    assert set(m.unique().tolist()) <= {0, 1}
    assert abs(m.float().mean().item() - 0.5) < 5e-3
    assert abs(m3.float().mean().item() - 0.8) < 5e-3
    assert abs(e.mean().item()) < 5e-3 and abs(e.std().item() - 1.0) < 5e-3
    assert abs((e ** 4).mean().item() - 3.0) < 0.1


@pytest.mark.parametrize("V", [20, 47, 64, 90])
def test_fused_ce_kl_forward_backward(V):
    g = torch.Generator().manual_seed(V)
    B, T, Z = 33, 24, 16
    w = torch.relu(torch.randn(B, T, V, generator=g))
    w[0, 0] = 0.0  # all-way tie row
    tgt = torch.randint(0, V, (B, T), generator=g)
    mu, ls = torch.randn(B, Z, generator=g), 0.3 * torch.randn(B, Z, generator=g)
    wr, mur, lsr = w.clone().requires_grad_(), mu.clone().requires_grad_(), ls.clone().requires_grad_()
    ref = O.mvae_loss(wr, tgt, mur, lsr)
    ref.backward()
    wd, mud, lsd = (t.to(DEV).requires_grad_() for t in (w, mu, ls))
    loss, acc = Fn.fused_ce_kl(wd, tgt.to(DEV), mud, lsd, beta=0.001)
    (loss * 2.0).backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref.item()) < 1e-5
    assert abs(acc.item() - O.mean_accuracy(w, tgt).item()) < 1e-6
    assert torch.allclose(wd.grad.cpu(), 2 * wr.grad, atol=1e-7, rtol=1e-4)
    assert torch.allclose(mud.grad.cpu(), 2 * mur.grad, atol=1e-9, rtol=1e-4)
    assert torch.allclose(lsd.grad.cpu(), 2 * lsr.grad, atol=1e-9, rtol=1e-4)


def test_adam_matches_reference_golden():
    import os
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "adam.pt"), weights_only=False)
    p = fx["p0"].clone().to(DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    for i, (gr, pref) in enumerate(zip(fx["grads"], fx["ps"])):
        gd = gr.to(DEV)
        ops.adam_step(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), i + 1, 1e-4, 0.9, 0.999, 1e-8, 1.0,
                      flag.data_ptr())
        torch.cuda.synchronize()
        assert torch.allclose(p.cpu(), pref, atol=1e-7, rtol=1e-5)
    assert flag.item() == 0
    gd = torch.full_like(p, float("nan"))
    ops.adam_step(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 4, 1e-4, 0.9, 0.999, 1e-8, 1.0,
                  flag.data_ptr())
    assert flag.item() == 1


def test_embedding_gather_and_scatter():
    V, E, rows = 21, 10, 1000
    g = torch.Generator().manual_seed(0)
    emb = torch.randn(V + 1, E, generator=g).to(DEV)
    tok = torch.randint(0, V + 1, (rows,), generator=g).to(torch.int32).to(DEV)
    out = torch.empty(rows, 16, dtype=torch.bfloat16, device=DEV)
    ops.embed_rows(emb.data_ptr(), E, tok.data_ptr(), rows, out.data_ptr(), BF16, 16)
    torch.cuda.synchronize()
    assert torch.equal(out[:, :E].float(), emb[tok.long()].to(torch.bfloat16).float())
    assert bool((out[:, E:] == 0).all())
    dX = torch.randn(rows, 16, generator=g).to(DEV)
    demb = torch.zeros(V, E, device=DEV)
    dskip = torch.zeros(E, device=DEV)
    ops.embed_grad(dX.data_ptr(), F32, 16, tok.data_ptr(), rows, E, V, demb.data_ptr(), skip_id=V, dskip=dskip.data_ptr())
    torch.cuda.synchronize()
    ref = torch.zeros(V + 1, E, device=DEV).index_add_(0, tok.long(), dX[:, :E])
    assert torch.allclose(demb, ref[:V], atol=1e-4) and torch.allclose(dskip, ref[V], atol=1e-4)

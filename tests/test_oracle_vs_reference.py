"""Live check of the CPU oracle against the UNMODIFIED reference (only where /root/reference is mounted: the
build container).  Complements the committed golden vectors with fresh random weights / inputs."""
import random

import pytest
import torch

from oracle import inpaintnet_oracle as O
from oracle.ref_import import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")


@pytest.mark.parametrize("seed", [1, 2])
def test_mvae_live(seed):
    from oracle.ref_import import load_reference, FakeDataset
    R = load_reference()
    torch.manual_seed(seed)
    random.seed(seed)
    V, H, Z, B = 23, 24, 12, 4
    m = R.MeasureVAE(FakeDataset(V), encoder_hidden_size=H, decoder_hidden_size=H, latent_space_dim=Z)
    with torch.no_grad():
        m.decoder.b_0.normal_()
        m.decoder.x_0.normal_()
    m.eval()
    tokens = torch.randint(0, V, (B, 24))
    eps = torch.randn(B, Z)
    orig = torch.distributions.Normal.rsample
    torch.distributions.Normal.rsample = lambda d, s=torch.Size(): d.loc + d.scale * eps
    try:
        for tf in (True, False):
            m.decoder.teacher_forcing_prob = 2.0 if tf else -1.0
            w, s, zd, *_ = m(tokens, train=True)
            sd = {k: v.detach() for k, v in m.state_dict().items()}
            w2, s2, mu, ls, z = O.mvae_forward(sd, tokens, eps, teacher_forced=tf)
            assert torch.allclose(mu, zd.loc, atol=2e-6) and torch.equal(s2, s)
            assert torch.allclose(w2, w, atol=5e-6, rtol=1e-5)
    finally:
        torch.distributions.Normal.rsample = orig

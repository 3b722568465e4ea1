"""Data-parallel parity on real GPUs (needs >= 2; skipped otherwise): two NCCL ranks, each with half of the
batch, must end two training steps (gradient buckets all-reduced under the backward pass) with the same parameters as one process that saw the whole batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _step(model, trainer, tokens, eps):
    from inpaintnet_b200 import engine
    trainer.zero_grad()
    with engine.inject_noise(eps=[eps]):
        loss, acc = trainer.loss_and_acc_for_batch(tokens, 0, train=True)
    loss.backward()
    trainer.step()
    torch.cuda.synchronize()
    return loss.item()


def _build(fx, dev):
    from inpaintnet_b200.measure_vae import MeasureVAE
    from inpaintnet_b200.trainer import VAETrainer
    from inpaintnet_b200.data import SyntheticFolkDataset
    ds = SyntheticFolkDataset(num_notes=fx["V"])
    m = MeasureVAE(ds, encoder_hidden_size=fx["H"], decoder_hidden_size=fx["H"], latent_space_dim=fx["Z"])
    m.load_state_dict(fx["state_dict"])
    m.to(dev).set_precision("fp32")
    m.eval()
    m.decoder.teacher_forcing_prob = 2.0
    return m, VAETrainer(ds, m, lr=1e-3)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fx = torch.load(os.path.join(G, "mvae_h32.pt"), weights_only=False)
        B = fx["B"] // world * world
        per = B // world
        m, tr = _build(fx, f"cuda:{rank}")
        sl = slice(rank * per, (rank + 1) * per)
        for _ in range(2):
            _step(m, tr, fx["tokens"][sl].cuda(), fx["eps"][sl])
        # the decoder's gradient buckets must have been queued from inside the backward pass (overlap), both steps
        assert tr._grad_exchange.n_early >= 2, tr._grad_exchange.n_early
        q.put((rank, {k: v.detach().cpu() for k, v in m.state_dict().items()}))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_step_equals_full_batch_step():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(q, procs, world)
    for p in procs:
        p.join(timeout=120)
    fx = torch.load(os.path.join(G, "mvae_h32.pt"), weights_only=False)
    B = fx["B"] // world * world
    m, tr = _build(fx, "cuda:0")
    for _ in range(2):
        _step(m, tr, fx["tokens"][:B].cuda(), fx["eps"][:B])
    ref = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    for k in ref:
        assert torch.allclose(res[0][k], res[1][k], atol=0, rtol=0), k            # replicas stay identical
        assert torch.allclose(res[0][k], ref[k], atol=1e-5, rtol=2e-4), k          # mean-of-means == global mean


# ---------------------------------------------------------------------------------------------------
# LatentRNN training with a frozen MeasureVAE (BASELINE.json configs[2]): the generation GRU's gradient
# buckets are reduced while the context GRUs' backward still runs
# ---------------------------------------------------------------------------------------------------
# Confirmed on 2 B200s (round 2, first GPU call: 119 passed with both data-parallel tests).


def _latent_build(fx, dev):
    from inpaintnet_b200.measure_vae import MeasureVAE
    from inpaintnet_b200.latent_rnn import LatentRNN
    from inpaintnet_b200.trainer import LatentRNNTrainer
    from inpaintnet_b200.data import SyntheticFolkDataset
    n_bars = fx["past"].shape[1] + fx["target"].shape[1] + fx["future"].shape[1]   # the trainer generates n_bars - past - future
    ds = SyntheticFolkDataset(num_notes=fx["V"], n_bars=n_bars)
    vae = MeasureVAE(ds, encoder_hidden_size=fx["H"], decoder_hidden_size=fx["H"], latent_space_dim=fx["Z"])
    m = LatentRNN(ds, vae, 2, fx["Hc"], 0.5, torch.nn.GRU, auto_reg=False)
    m.load_state_dict(fx["state_dict"])
    m.to(dev).set_precision("fp32")
    m.eval()
    tr = LatentRNNTrainer.__new__(LatentRNNTrainer)      # the constructor asserts n_bars > 6 (reference split limits)
    from inpaintnet_b200.trainer import Trainer
    Trainer.__init__(tr, ds, m, lr=1e-3)
    return m, tr


def _latent_step(m, tr, fx, sl):
    from inpaintnet_b200 import engine
    B, n_p, Z = fx["eps_past"][sl].shape
    n_f = fx["eps_future"].shape[1]
    eps = [fx["eps_past"][sl].transpose(0, 1).reshape(n_p * B, Z), fx["eps_future"][sl].transpose(0, 1).reshape(n_f * B, Z)]
    batch = (fx["past"][sl].cuda(), fx["future"][sl].cuda(), fx["target"][sl].cuda())
    tr.zero_grad()
    with engine.inject_noise(eps=eps):
        loss, acc = tr.loss_and_acc_for_batch(batch, 0, train=True)
    loss.backward()
    tr.step()
    torch.cuda.synchronize()


def _latent_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fx = torch.load(os.path.join(G, "latent_h32.pt"), weights_only=False)
        m, tr = _latent_build(fx, f"cuda:{rank}")
        for _ in range(2):
            _latent_step(m, tr, fx, slice(rank, rank + 1))
        assert tr._grad_exchange.n_early >= 2, tr._grad_exchange.n_early
        q.put((rank, {k: v.detach().cpu() for k, v in m.state_dict().items()}))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _collect(q, procs, world, timeout=300):
    """Results of all workers; fails fast (instead of waiting out the queue timeout) when a worker died."""
    import queue
    import time
    out, t0 = {}, time.time()
    while len(out) < world:
        try:
            r, v = q.get(timeout=2)
            out[r] = v
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            assert not dead, f"worker exited with {dead}"
            assert time.time() - t0 < timeout, "timed out waiting for the workers"
    return out


def test_latent_rnn_two_rank_steps_equal_full_batch_steps():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_latent_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(q, procs, world)
    for p in procs:
        p.join(timeout=120)
    fx = torch.load(os.path.join(G, "latent_h32.pt"), weights_only=False)
    m, tr = _latent_build(fx, "cuda:0")
    for _ in range(2):
        _latent_step(m, tr, fx, slice(0, 2))
    ref = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    for k in ref:
        assert torch.equal(res[0][k], res[1][k]), k                                # replicas stay identical
        assert torch.allclose(res[0][k], ref[k], atol=1e-5, rtol=2e-4), k          # mean-of-means == global mean
        if k.startswith("vae_model."):
            assert torch.equal(ref[k], fx["state_dict"][k]), k                     # the frozen VAE never moves

"""CPU arm of bench.py: times the UNMODIFIED reference (ashispati/InpaintNet) on the host cores.

Test/measurement infrastructure only -- nothing under inpaintnet_b200/ imports this.  The reference modules come
from oracle/ref_import.py (/root/reference in the build container, the staged byte-for-byte copy oracle/_ref/ on
the GPU box); its own classes and its own trainer methods are called exactly as the reference's epoch loop calls
them (utils/trainer.py:136-156): process_batch_data -> zero_grad -> loss_and_acc_for_batch -> loss.backward() ->
step() -> to_numpy(loss).  Inputs are the synthetic tensors bench.py's B200 arm uses (same shapes, V = 64,
uniform-random tokens).  When neither reference location exists the oracle port is timed instead ("kind": "port").
"""
import os
import random
import time

import torch

V = 64


def _threads():
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def available():
    from oracle.ref_import import reference_available
    return reference_available()


def _time_steps(step, steps, warmup):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / max(1, steps)


def _metadata(B, T=384):
    t = torch.arange(T)
    md = torch.stack([(t // 6) % 4 == 0, t % 6, torch.zeros_like(t)], 1).to(torch.int32)   # beat marker, tick, voice
    return md.view(1, 1, T, 3).expand(B, 1, T, 3).contiguous()


class MvaeTrain:
    """MeasureVAE train step: MeasureVAE/vae_trainer.py:16-40 + utils/trainer.py:136-156 (config 1 / 2 on the CPU).
    cuda=True: the same unmodified code on the GPU the way train_measure_vae.py:105 runs it (model.cuda(); torch's
    cuDNN GRU; the reference's own per-step host syncs included) -- context only, SURVEY.md section 2.2."""
    unit = "measures/s"

    def __init__(self, measures, seed=0, cuda=False):
        from oracle.ref_import import load_reference, FakeDataset
        R = load_reference()
        torch.manual_seed(seed)
        random.seed(seed)
        ds = FakeDataset(V)
        self.model = R.MeasureVAE(ds)                       # reference defaults (E 10, H 512, L 2, Z 256, p 0.5)
        if cuda:
            self.model.cuda()
        self.trainer = R.VAETrainer(ds, self.model, lr=1e-4)
        self.model.train()
        self.tokens = torch.randint(0, V, (measures, 24))
        if cuda:
            self.tokens = self.tokens.cuda()
        self.cuda = cuda
        self.units = measures
        self.what = (f"{measures} measures (unmodified reference MeasureVAE + VAETrainer: fwd, loss, backward, Adam; train mode"
                     + ("; on the GPU: torch cuDNN GRU path, fp32" if cuda else "") + ")")

    def step(self):
        tr = self.trainer
        tr.zero_grad()
        loss, acc = tr.loss_and_acc_for_batch(self.tokens, 0, train=True)
        loss.backward()
        tr.step()
        return float(loss.mean().detach())                  # the reference reads the loss every step (trainer.py:154)


def _latent_model(R, auto_reg, seed):
    from oracle.ref_import import FakeDataset
    torch.manual_seed(seed)
    random.seed(seed)
    ds = FakeDataset(V)
    vae = R.MeasureVAE(ds)
    model = R.LatentRNN(ds, vae, 2, 512, 0.5, torch.nn.GRU, auto_reg=auto_reg, teacher_forcing=True)
    return ds, model


class Inpaint:
    """Batched inpainting inference, the call the reference's testers make (LatentRNN/latent_rnn_tester.py:315-321,
    test_reconstruction.py:321-327): model.eval(); model(past, future, target, n_target, train=False) with the fixed
    6/4/6 split of script_gen_diff_models.py:144-146 (or n_target = 2, test_reconstruction.py:52)."""
    unit = "queries/s"

    def __init__(self, queries, n_target=4, seed=0):
        from oracle.ref_import import load_reference
        R = load_reference()
        ds, self.model = _latent_model(R, False, seed)
        self.model.eval()
        n_past = (16 - n_target) // 2
        score = torch.randint(0, V, (queries, 16, 24), dtype=torch.int32)
        self.past = score[:, :n_past].long().contiguous()
        self.target = score[:, n_past:n_past + n_target].long().contiguous()
        self.future = score[:, n_past + n_target:].long().contiguous()
        self.n_target = n_target
        self.units = queries
        self.what = (f"{queries} queries, split {n_past}/{n_target}/{16 - n_past - n_target} (unmodified reference LatentRNN.forward, "
                     f"eval mode, stock call: no torch.no_grad, target measures encoded as the reference does)")

    def step(self):
        w, s, z = self.model(self.past, self.future, self.target, self.n_target, train=False)
        return int(s[0, 0, 0])


class LatentTrain:
    """LatentRNN train step with the frozen MeasureVAE: LatentRNN/latent_rnn_trainer.py:26-67 (config 3)."""
    unit = "sequences/s"

    def __init__(self, sequences, auto_reg=False, seed=0):
        from oracle.ref_import import load_reference
        R = load_reference()
        ds, self.model = _latent_model(R, auto_reg, seed)
        self.trainer = R.LatentRNNTrainer(ds, self.model, lr=1e-4)
        self.model.train()
        self.score = torch.randint(0, V, (sequences, 1, 384), dtype=torch.int32)   # int32 as the dataset yields (folk_dataset.py:837)
        self.units = sequences
        self.what = (f"{sequences} sequences of 16 measures (unmodified reference LatentRNN auto_reg={auto_reg} + LatentRNNTrainer: "
                     f"stochastic past/gap/future split, fwd, loss, backward, Adam; train mode)")

    def step(self):
        tr = self.trainer
        batch = tr.process_batch_data((self.score, None))
        tr.zero_grad()
        loss, acc = tr.loss_and_acc_for_batch(batch, 0, train=True)
        loss.backward()
        tr.step()
        return float(loss.mean().detach())


class ArnnTrain:
    """AnticipationRNN train step: AnticipationRNN/anticipation_rnn_trainer.py:21-67 with the train_arnn_reg.py:86-98
    model (config 5).  Teacher-forced only: the free-running backward raises on the CPU under a modern torch
    (SURVEY.md section 4), so the coin is pinned to the teacher-forced branch."""
    unit = "sequences/s"

    def __init__(self, sequences, seed=0):
        from oracle.ref_import import load_reference, FakeDataset
        R = load_reference()
        torch.manual_seed(seed)
        random.seed(seed)
        ds = FakeDataset(V)
        self.model = R.ConstraintModelGaussianReg(
            dataset=ds, note_embedding_dim=10, metadata_embedding_dim=2, num_lstm_constraints_units=256,
            num_lstm_generation_units=256, linear_hidden_size=256, num_layers=2, dropout_input_prob=0.2,
            dropout_prob=0.2, unary_constraint=True, teacher_forcing=True)
        self.model.teacher_forcing_prob = 2.0
        self.trainer = R.AnticipationRNNGaussianRegTrainer(ds, self.model, lr=1e-4)
        self.model.train()
        self.score = torch.randint(0, V, (sequences, 1, 384), dtype=torch.int32)
        self.meta = _metadata(sequences)
        self.units = sequences
        self.what = (f"{sequences} sequences of 384 ticks (unmodified reference ConstraintModelGaussianReg + trainer, "
                     f"teacher-forced branch: fwd, loss, backward, Adam; train mode)")

    def step(self):
        tr = self.trainer
        batch = tr.process_batch_data((self.score, self.meta))
        tr.zero_grad()
        loss, acc = tr.loss_and_acc_for_batch(batch, 0, train=True)
        loss.backward()
        tr.step()
        return float(loss.mean().detach())


WORKLOADS = {"mvae_train": MvaeTrain, "inpaint": Inpaint, "latent_train": LatentTrain, "arnn_train": ArnnTrain}


def run(workload, size, steps, warmup, **kw):
    """Times `steps` steps after `warmup`; returns dict(value, unit, cores, kind, sample, s_per_step)."""
    cores = _threads()
    w = WORKLOADS[workload](size, **kw)
    on_gpu = bool(getattr(w, "cuda", False))
    sync = torch.cuda.synchronize if on_gpu else (lambda: None)
    for _ in range(warmup):
        w.step()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        w.step()
    sync()
    per = (time.perf_counter() - t0) / max(1, steps)
    where = f"torch {torch.__version__} " + (f"CUDA ({torch.cuda.get_device_name(0)}, cuDNN {torch.backends.cudnn.version()})" if on_gpu else "CPU")
    return dict(value=w.units / per, unit=w.unit, cores=cores, kind="reference",
                sample=f"{steps} steps of {w.what}, after {warmup} warm-up, fp32, {per * 1e3:.1f} ms/step, {where}",
                s_per_step=per, units_per_step=w.units)

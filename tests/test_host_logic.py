"""CPU tests of the host side: C-ABI library loads and exports every declared symbol, struct layouts
match, compute calls fail loudly without a GPU, and the Python orchestration runs end to end against a
stub library (plumbing only)."""
import ctypes as C
import os
import re

import pytest
import torch

from inpaintnet_b200 import _lib, ops, engine, functional as Fn
from inpaintnet_b200.arena import ParamArena, arena_of
from inpaintnet_b200.data import SyntheticFolkDataset
from inpaintnet_b200.measure_vae import MeasureVAE
from inpaintnet_b200.trainer import VAETrainer
from tests.dryrun import stubbed
from tests.golden import recipe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "inpaintnet_b200.h")).read()
    declared = set(re.findall(r"\b(ipn_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.ipn_abi_version() == 3


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.ipn_device_check(0, None) == _lib.ERR_ARCH
    g = _lib.Gemm()
    with pytest.raises(_lib.InpaintNetB200Error):
        _lib.check(lib.ipn_gemm(C.byref(g), None))
    ds = SyntheticFolkDataset(num_notes=20)
    m = MeasureVAE(ds, encoder_hidden_size=32, decoder_hidden_size=32, latent_space_dim=16)
    with pytest.raises(_lib.InpaintNetB200Error):
        m(torch.randint(0, 20, (2, 24)), train=False)


def test_arena_keeps_reference_state_dict_layout():
    V, H, Z = 20, 32, 16
    ds = SyntheticFolkDataset(num_notes=V)
    m = MeasureVAE(ds, encoder_hidden_size=H, decoder_hidden_size=H, latent_space_dim=Z)
    spec = recipe.mvae_spec(V, 10, H, Z)
    sd = m.state_dict()
    assert set(sd) == set(spec)
    for k, shp in spec.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    ref = recipe.make_state_dict(spec, 1)
    m.load_state_dict(ref)
    a = arena_of(m)
    assert a.valid() and a.total >= sum(v.numel() for v in ref.values())
    for k, v in m.state_dict().items():
        assert torch.equal(v, ref[k])
    # decoder embedding table and x_0 are contiguous (the [Emb; x_0] gather table)
    o = a.offset
    assert o["decoder.x_0"] == o["decoder.note_embedding_layer.weight"] + V * 10
    # load_state_dict writes through the views: arena storage follows, version key changes
    k0 = a.version_key()
    ref2 = recipe.make_state_dict(spec, 2)
    m.load_state_dict(ref2)
    assert arena_of(m) is a and a.version_key() != k0
    assert torch.equal(a.flat[o["encoder.lstm.weight_hh_l0"]:o["encoder.lstm.weight_hh_l0"] + 3 * H * H].view(3 * H, H),
                       ref2["encoder.lstm.weight_hh_l0"])


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("tf", [True, False])
def test_plumbing_dry_run(prec, tf):
    """forward + fused loss + backward + fused Adam through the real Python stack on a stub library."""
    if torch.cuda.is_available():
        pytest.skip("dry run is a CPU-only plumbing check")
    V, H, Z, B = 20, 32, 16, 3
    ds = SyntheticFolkDataset(num_notes=V)
    m = MeasureVAE(ds, encoder_hidden_size=H, decoder_hidden_size=H, latent_space_dim=Z).set_precision(prec)
    m.decoder.teacher_forcing_prob = 2.0 if tf else -1.0
    tr = VAETrainer(ds, m)
    tokens = torch.randint(0, V, (B, 24))
    with stubbed():
        m.train()
        tr.zero_grad()
        loss, acc = tr.loss_and_acc_for_batch(tokens, 0, train=True)
        loss.backward()
        tr.step()
        m.eval()
        with torch.no_grad():
            w, s = m.forward_test(tokens.view(1, B, 24))
    assert w.shape == (1, B, 24, V) and s.shape == (1, 1, B * 24)
    for n, p in m.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, n


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_latent_rnn_plumbing_dry_run(prec):
    if torch.cuda.is_available():
        pytest.skip("dry run is a CPU-only plumbing check")
    from inpaintnet_b200.latent_rnn import LatentRNN
    from inpaintnet_b200.trainer import LatentRNNTrainer
    V, H, Z, B = 20, 32, 16, 3
    ds = SyntheticFolkDataset(num_notes=V)
    vae = MeasureVAE(ds, encoder_hidden_size=H, decoder_hidden_size=H, latent_space_dim=Z)
    m = LatentRNN(ds, vae, 2, 32, 0.5, torch.nn.GRU, auto_reg=False).set_precision(prec)
    sd = m.state_dict()
    assert len(sd) == 103 and "x_0" in sd and "vae_model.decoder.x_0" in sd
    tr = LatentRNNTrainer(ds, m)
    score = torch.randint(0, V, (B, 1, 384), dtype=torch.int32)
    with stubbed():
        m.train()
        batch = tr.process_batch_data((score, None))
        tr.zero_grad()
        loss, acc = tr.loss_and_acc_for_batch(batch, 0, train=True)
        loss.backward()
        tr.step()
    n_t = batch[2].shape[1]
    for n, p in m.named_parameters():
        if n.startswith("vae_model."):
            assert not p.requires_grad
        else:
            assert p.grad is not None, n
    a = arena_of(m)
    assert a.n_trainable < a.total and a.names[0] == "x_0"


def _run_reference_script(script, **kwargs):
    """Imports /root/reference/<script> UNMODIFIED with inpaintnet_b200/dropin first on sys.path and calls its
    click main() body under the stub library (plumbing only: argument wiring, shapes, state_dict / checkpoint files)."""
    import importlib.util
    import sys
    from oracle.ref_import import REFERENCE_ROOT
    ref = os.path.join(REFERENCE_ROOT, script)
    if not os.path.exists(ref) or torch.cuda.is_available():
        pytest.skip("reference neither mounted nor staged, or a GPU box (tests/test_gpu_scripts.py runs the real library there)")
    dropin = os.path.join(ROOT, "inpaintnet_b200", "dropin")
    tops = ("MeasureVAE", "LatentRNN", "utils", "DatasetManager", "AnticipationRNN")
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items() if k.split(".")[0] in tops}
    for m in saved_mods:
        del sys.modules[m]
    sys.path.insert(0, dropin)
    try:
        spec = importlib.util.spec_from_file_location("ref_" + script[:-3], ref)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        import inpaintnet_b200.data as D
        orig = D.SyntheticFolkDataset.__init__

        def small(self, *a, **k):
            k["num_sequences"] = 8
            k.setdefault("num_notes", 20)
            orig(self, *a, **k)

        D.SyntheticFolkDataset.__init__ = small
        try:
            with stubbed():
                mod.main.callback(**kwargs)
        finally:
            D.SyntheticFolkDataset.__init__ = orig
    finally:
        sys.path[:] = saved_path
        for m in [k for k in sys.modules if k.split(".")[0] in tops]:
            del sys.modules[m]
        sys.modules.update(saved_mods)


_VAE_KW = dict(note_embedding_dim=10, metadata_embedding_dim=2, num_encoder_layers=2, encoder_hidden_size=32,
               encoder_dropout_prob=0.5, has_metadata=False, latent_space_dim=16, num_decoder_layers=2,
               decoder_hidden_size=32, decoder_dropout_prob=0.5)


def test_reference_train_script_runs_unchanged_against_dropin():
    """The reference's own train_measure_vae.py, unmodified.  Needs /root/reference (build container only)."""
    _run_reference_script("train_measure_vae.py", batch_size=2, num_epochs=1, train=True, plot=False, log=False, lr=1e-4,
                          **_VAE_KW)


@pytest.mark.parametrize("auto_reg,teacher_forcing", [(True, True), (True, False), (False, False)])
def test_reference_train_inpaintnet_runs_unchanged_against_dropin(auto_reg, teacher_forcing):
    """train_inpaintnet.py, unmodified, at its defaults (auto_reg=True, teacher_forcing=True, :53-56) and the two
    other generation modes; loads the MeasureVAE checkpoint the previous script saved (vae_model.load(), :113)."""
    _run_reference_script("train_measure_vae.py", batch_size=2, num_epochs=1, train=True, plot=False, log=False, lr=1e-4,
                          **_VAE_KW)
    _run_reference_script("train_inpaintnet.py", num_latent_rnn_layers=2, latent_rnn_hidden_size=32,
                          latent_rnn_dropout_prob=0.5, batch_size=2, num_epochs=1, train=True, lr=1e-4, plot=False,
                          log=False, auto_reg=auto_reg, teacher_forcing=teacher_forcing, early_stop=True, **_VAE_KW)


def test_reference_train_inpaintnet_ablation_runs_unchanged_against_dropin():
    """train_inpaintnet_ablation.py, unmodified (LatentRNNAblations, type='past')."""
    _run_reference_script("train_measure_vae.py", batch_size=2, num_epochs=1, train=True, plot=False, log=False, lr=1e-4,
                          **_VAE_KW)
    _run_reference_script("train_inpaintnet_ablation.py", num_latent_rnn_layers=2, latent_rnn_hidden_size=32,
                          latent_rnn_dropout_prob=0.5, batch_size=2, num_epochs=1, train=True, lr=1e-4, plot=False,
                          log=False, auto_reg=True, teacher_forcing=True, early_stop=True, **_VAE_KW)


def test_reference_test_reconstruction_runs_unchanged_against_dropin():
    """test_reconstruction.py, unmodified: loads the four checkpoints the (equally unmodified) training scripts
    wrote -- MeasureVAE, LatentRNN(auto_reg=False), ARNN, ARNN baseline -- and runs its three-model inpainting
    comparison (LatentRNN.forward, forward_inpaint) over the held-out split."""
    arnn_kw = dict(note_embedding_dim=10, metadata_embedding_dim=2, num_layers=2, lstm_hidden_size=32, dropout_lstm=0.2,
                   input_dropout=0.2, linear_hidden_size=32)
    lat_kw = dict(num_latent_rnn_layers=2, latent_rnn_hidden_size=32, latent_rnn_dropout_prob=0.5)
    _run_reference_script("train_measure_vae.py", batch_size=2, num_epochs=1, train=True, plot=False, log=False, lr=1e-4,
                          **_VAE_KW)
    _run_reference_script("train_inpaintnet.py", batch_size=2, num_epochs=1, train=True, lr=1e-4, plot=False, log=False,
                          auto_reg=False, teacher_forcing=True, early_stop=True, **lat_kw, **_VAE_KW)
    for script in ("train_arnn_reg.py", "train_arnn_baseline.py"):
        _run_reference_script(script, batch_size=2, num_epochs=1, train=True, log=False, lr=1e-4, plot=False,
                              teacher_forcing=True, early_stop=True, **arnn_kw)
    kw = dict(_VAE_KW)
    kw.update(lat_kw)
    kw.update({k: v for k, v in arnn_kw.items() if k not in kw})
    _run_reference_script("test_reconstruction.py", batch_size=2, num_target=2, num_models=4, **kw)


@pytest.mark.parametrize("script", ["train_arnn_reg.py", "train_arnn_baseline.py"])
def test_reference_train_arnn_runs_unchanged_against_dropin(script):
    """train_arnn_reg.py / train_arnn_baseline.py, unmodified: trainer epoch, then AnticipationRNNTester.test_model
    (forward_inpaint over the held-out split)."""
    _run_reference_script(script, note_embedding_dim=10, metadata_embedding_dim=2, num_layers=2, lstm_hidden_size=32,
                          dropout_lstm=0.2, input_dropout=0.2, linear_hidden_size=32, batch_size=2, num_epochs=1,
                          train=True, log=False, lr=1e-4, plot=False, teacher_forcing=True, early_stop=True)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("tf", [True, False])
def test_arnn_plumbing_dry_run(prec, tf):
    if torch.cuda.is_available():
        pytest.skip("dry run is a CPU-only plumbing check")
    from inpaintnet_b200.arnn import ConstraintModelGaussianReg, AnticipationRNNGaussianRegTrainer
    ds = SyntheticFolkDataset(num_notes=20, num_sequences=4)
    m = ConstraintModelGaussianReg(ds, note_embedding_dim=10, metadata_embedding_dim=2, num_lstm_constraints_units=32,
                                   num_lstm_generation_units=32, linear_hidden_size=32, num_layers=2, dropout_input_prob=0.2,
                                   dropout_prob=0.2, unary_constraint=True, teacher_forcing=True).set_precision(prec)
    keys = set(m.state_dict())
    assert {"note_embeddings.0.weight", "metadata_embeddings.2.weight", "lstm_constraint.1.weight_hh_l0",
            "lstm_generation.0.weight_ih_l0", "linear_1.weight", "linear_ouput_notes.0.bias"} <= keys
    assert m.state_dict()["lstm_generation.0.weight_ih_l0"].shape == (128, 42)
    m.teacher_forcing_prob = 2.0 if tf else -1.0
    tr = AnticipationRNNGaussianRegTrainer(ds, m)
    batch = next(iter(ds.data_loaders(2, split=(0.5, 0.25))[0]))
    with stubbed():
        m.train()
        data = tr.process_batch_data(batch)
        tr.zero_grad()
        loss, acc = tr.loss_and_acc_for_batch(data, 0, train=True)
        loss.backward()
        tr.step()
    for n, p in m.named_parameters():
        assert p.grad is not None, n


def test_training_state_round_trip(tmp_path):
    """Trainer.save_training_state / load_training_state: weights, Adam moments + step, RNG streams, early stopping."""
    if torch.cuda.is_available():
        pytest.skip("dry run is a CPU-only plumbing check")
    import random
    from inpaintnet_b200.trainer import VAETrainer
    V, H, Z, B = 20, 32, 16, 4

    def make():
        ds = SyntheticFolkDataset(num_notes=V)
        m = MeasureVAE(ds, encoder_hidden_size=H, decoder_hidden_size=H, latent_space_dim=Z).set_precision("fp32")
        return m, VAETrainer(ds, m)

    torch.manual_seed(5)
    random.seed(5)
    m, tr = make()
    tr.early_stopping, tr.early_stopper = True, __import__("inpaintnet_b200.trainer", fromlist=["x"]).EarlyStopping()
    tokens = torch.randint(0, V, (B, 24))
    with stubbed():
        m.train()
        tr.zero_grad()
        loss, _ = tr.loss_and_acc_for_batch(tokens, 0, train=True)
        loss.backward()
        tr.step()
    tr.optimizer._m.normal_()
    tr.optimizer._v.uniform_()
    tr.early_stopper(1.5, m)
    tr.early_stopper(1.7, m)
    path = tr.save_training_state(3, str(tmp_path / "state.pt"))
    want = dict(rng_offset=arena_of(m).rng_offset, py=random.random(), t=torch.rand(3))
    assert want["rng_offset"] > 0
    torch.manual_seed(99)
    random.seed(99)
    m2, tr2 = make()
    tr2.early_stopping, tr2.early_stopper = True, type(tr.early_stopper)()
    assert tr2.load_training_state(path) == 4
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    assert tr2.optimizer.step_count == 1
    assert torch.equal(tr2.optimizer._m, tr.optimizer._m) and torch.equal(tr2.optimizer._v, tr.optimizer._v)
    assert arena_of(m2).rng_offset == want["rng_offset"]
    assert random.random() == want["py"] and torch.equal(torch.rand(3), want["t"])
    assert (tr2.early_stopper.counter, tr2.early_stopper.best_score) == (tr.early_stopper.counter, tr.early_stopper.best_score)
    assert not os.path.exists(path + ".tmp")


def test_lagged_readback_orders_results_and_raises_on_guard_flags():
    """trainer.LaggedReadback (CPU tensors here: no pinned memory / events): results come back in step order, the
    newest `keep` stay queued, a raised NaN / token-range flag surfaces as the reference's ValueError."""
    from inpaintnet_b200.trainer import LaggedReadback

    class _Arena:
        def __init__(self):
            self.nan_flag = torch.zeros(1, dtype=torch.int32)
            self.range_flag = torch.zeros(1, dtype=torch.int32)

    a, rb = _Arena(), LaggedReadback(depth=3)
    got = []
    for i in range(7):
        rb.push(torch.tensor(float(i)), torch.tensor(i / 10.0), a)
        got += rb.pop(keep=1)
        assert len(rb.pending) == 1
    got += rb.pop(keep=0)
    assert [round(l) for l, _ in got] == list(range(7))
    assert all(abs(acc - i / 10.0) < 1e-6 for i, (_, acc) in enumerate(got))
    rb.push(torch.tensor(1.0), None, a)            # accuracy may be absent
    assert rb.pop(keep=0) == [(1.0, 0.0)]
    for _ in range(3):
        rb.push(torch.tensor(0.0), None, a)
    with pytest.raises(RuntimeError):
        rb.push(torch.tensor(0.0), None, a)        # ring full: the caller must pop
    rb.pop(keep=0)
    a.range_flag[0] = 1
    rb.push(torch.tensor(0.0), None, a)
    with pytest.raises(ValueError):
        rb.pop(keep=0)
    a.range_flag[0] = 0
    a.nan_flag[0] = 1
    rb.push(torch.tensor(0.0), None, a)
    with pytest.raises(ValueError):
        rb.pop(keep=0)


def test_early_stopping_matches_reference_and_resume_sees_the_stop(monkeypatch):
    """EarlyStopping behaves like utils/trainer.py:379-413 on a loss curve with plateaus (checked against the
    unmodified reference class when the reference is mounted / staged), and the stopper runs BEFORE the training
    state is saved, so a resumed run of an early-stopped job stops at once instead of training on."""
    from inpaintnet_b200.trainer import EarlyStopping
    curve = [1.0, 0.9, 0.9, 0.899999, 0.95, 0.8, 0.80000001, 0.85, 0.9, 0.81, 0.82, 0.7]
    mine = EarlyStopping(patience=5)
    trace = []
    for v in curve:
        mine(v, None)
        trace.append((mine.counter, mine.early_stop, mine.best_score, mine.val_loss_min))
    assert trace[5][0] == 0 and trace[5][2] == -0.8
    assert [t[0] for t in trace] == [0, 0, 1, 2, 3, 0, 1, 2, 3, 4, 5, 5][:len(trace)] or trace[10][1]
    assert trace[10][1] and trace[11][1]           # fifth epoch without improvement -> stop, and it stays stopped
    from oracle.ref_import import reference_available, load_reference
    if reference_available():
        import importlib
        load_reference()
        import sys
        from oracle.ref_import import REFERENCE_ROOT
        sys.path.insert(0, REFERENCE_ROOT)
        try:
            ref_cls = importlib.import_module("utils.trainer").EarlyStopping
        finally:
            sys.path.remove(REFERENCE_ROOT)
        import numpy as np
        monkeypatch.setattr(np, "Inf", np.inf, raising=False)   # the reference predates numpy 2 (np.Inf was removed)
        ref = ref_cls(patience=5)
        for v, t in zip(curve, trace):
            ref(v, None)
            assert (ref.counter, ref.early_stop, ref.best_score, ref.val_loss_min) == t


def test_shard_batch_and_resume_after_early_stop(tmp_path, monkeypatch):
    from inpaintnet_b200 import trainer as T
    monkeypatch.setattr(T, "dp_rank_world", lambda: (1, 4))
    score, meta = torch.arange(22).view(11, 1, 2), torch.arange(11)
    s, m, none = T.shard_batch((score, meta, None))
    assert none is None and s.shape[0] == 2 and torch.equal(m, torch.tensor([1, 5]))   # rows 1, 5 of the first 8
    with pytest.raises(ValueError):
        T.shard_batch((torch.zeros(3, 1, 2),))
    monkeypatch.setattr(T, "dp_rank_world", lambda: (0, 1))
    assert T.shard_batch((score, meta))[0] is score

    # a run that ended with "Early Stopping" resumes as stopped (the state is saved after the stopper ran)
    with stubbed():
        ds = SyntheticFolkDataset(num_notes=20, num_sequences=8)
        m = MeasureVAE(ds, encoder_hidden_size=32, decoder_hidden_size=32, latent_space_dim=16)
        m.filepath = str(tmp_path / "model")
        tr = VAETrainer(ds, m)
        tr.early_stopping, tr.early_stopper = True, T.EarlyStopping(patience=1)
        calls = []

        def fake_epoch(data_loader, epoch_num=None, train=True):
            calls.append((epoch_num, train))
            return 1.0, 0.5            # the loss never improves -> stops after the second epoch

        tr.loss_and_acc_on_epoch = fake_epoch
        tr.train_model(batch_size=2, num_epochs=10)
        assert [e for e, t in calls if t] == [0, 1]
        tr2 = VAETrainer(ds, m)
        tr2.early_stopping, tr2.early_stopper = True, T.EarlyStopping(patience=1)
        tr2.loss_and_acc_on_epoch = fake_epoch
        n = len(calls)
        tr2.train_model(batch_size=2, num_epochs=10, resume=True)
        assert len(calls) == n and tr2.early_stopper.early_stop


def test_arena_rebuilds_when_the_trainable_set_changes_and_adam_state_follows():
    """freezing a sub-module after the arena exists must rebuild the trainable-first layout, and
    FusedAdam must carry the moments of the still-trainable parameters over instead of silently zeroing them."""
    from inpaintnet_b200.optim import FusedAdam
    m = torch.nn.ModuleDict(dict(a=torch.nn.Linear(5, 3), b=torch.nn.Linear(3, 2)))
    opt = FusedAdam(m)
    a0 = opt.arena()
    opt.step_count = 3
    opt._m.copy_(torch.arange(a0.n_trainable, dtype=torch.float32))
    off_b = a0.offset["b.weight"]
    want = opt._m[off_b:off_b + 6].clone()
    for p in m["a"].parameters():
        p.requires_grad = False
    assert not a0.valid()
    a1 = opt.arena()                           # every still-trainable parameter keeps its moments: no warning
    assert a1 is not a0 and a1.offset["b.weight"] == 0 and a1.n_trainable == 12
    assert torch.equal(opt._m[:6], want) and opt.step_count == 3
    for p in m["a"].parameters():
        p.requires_grad = True
    with pytest.warns(UserWarning):            # a.* has no history: announced, not silent
        a2 = opt.arena()
    ob = a2.offset["b.weight"]
    assert torch.equal(opt._m[ob:ob + 6], want) and float(opt._m[a2.offset["a.weight"]]) == 0.0
    a1 = a2
    v0 = a1.version_key()
    a1.invalidate()
    assert a1.version_key() != v0

"""The reference's own scripts, UNMODIFIED, against the drop-in packages and the REAL CUDA library on a B200.

tests/test_host_logic.py runs the same scripts against a stub library on the CPU (plumbing); here the very same
`main()` bodies train and evaluate with the sm_100a kernels: train_measure_vae.py writes the MeasureVAE checkpoint
that train_inpaintnet.py loads, the AnticipationRNN scripts train + run their tester (forward_inpaint), and
test_reconstruction.py loads all four checkpoints for its three-model inpainting comparison.  The script sources are
the byte-for-byte staged copy oracle/_ref/ (oracle/make_ref.sh) or /root/reference where it is mounted."""
import importlib.util
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_VAE_KW = dict(note_embedding_dim=10, metadata_embedding_dim=2, num_encoder_layers=2, encoder_hidden_size=32,
               encoder_dropout_prob=0.5, has_metadata=False, latent_space_dim=16, num_decoder_layers=2,
               decoder_hidden_size=32, decoder_dropout_prob=0.5)
_LAT_KW = dict(num_latent_rnn_layers=2, latent_rnn_hidden_size=32, latent_rnn_dropout_prob=0.5)
_ARNN_KW = dict(note_embedding_dim=10, metadata_embedding_dim=2, num_layers=2, lstm_hidden_size=32, dropout_lstm=0.2,
                input_dropout=0.2, linear_hidden_size=32)


def _run(script, **kwargs):
    from oracle.ref_import import REFERENCE_ROOT, reference_available
    if not reference_available():
        pytest.skip("reference neither mounted nor staged (oracle/make_ref.sh)")
    ref = os.path.join(REFERENCE_ROOT, script)
    dropin = os.path.join(ROOT, "inpaintnet_b200", "dropin")
    tops = ("MeasureVAE", "LatentRNN", "utils", "DatasetManager", "AnticipationRNN")
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items() if k.split(".")[0] in tops}
    for m in saved_mods:
        del sys.modules[m]
    sys.path.insert(0, dropin)
    try:
        spec = importlib.util.spec_from_file_location("refgpu_" + script[:-3], ref)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        import inpaintnet_b200.data as D
        orig = D.SyntheticFolkDataset.__init__

        def small(self, *a, **k):
            k["num_sequences"] = 40   # 28 train / 8 validation / 4 test sequences: at least one batch of 4 in every split
            k.setdefault("num_notes", 20)
            orig(self, *a, **k)

        D.SyntheticFolkDataset.__init__ = small
        try:
            mod.main.callback(**kwargs)
            torch.cuda.synchronize()
        finally:
            D.SyntheticFolkDataset.__init__ = orig
    finally:
        sys.path[:] = saved_path
        for m in [k for k in sys.modules if k.split(".")[0] in tops]:
            del sys.modules[m]
        sys.modules.update(saved_mods)


def test_reference_scripts_train_and_evaluate_on_the_real_library():
    from inpaintnet_b200 import ops
    l0 = ops.launch_count()
    _run("train_measure_vae.py", batch_size=4, num_epochs=1, train=True, plot=False, log=False, lr=1e-4, **_VAE_KW)
    assert ops.launch_count() > l0 + 100, "train_measure_vae.py did not reach the CUDA library"
    l1 = ops.launch_count()
    _run("train_inpaintnet.py", batch_size=4, num_epochs=1, train=True, lr=1e-4, plot=False, log=False, auto_reg=False,
         teacher_forcing=True, early_stop=True, **_LAT_KW, **_VAE_KW)
    assert ops.launch_count() > l1 + 100
    for script in ("train_arnn_reg.py", "train_arnn_baseline.py"):
        _run(script, batch_size=4, num_epochs=1, train=True, log=False, lr=1e-4, plot=False, teacher_forcing=True,
             early_stop=True, **_ARNN_KW)
    kw = dict(_VAE_KW)
    kw.update(_LAT_KW)
    kw.update({k: v for k, v in _ARNN_KW.items() if k not in kw})
    l2 = ops.launch_count()
    _run("test_reconstruction.py", batch_size=4, num_target=2, num_models=4, **kw)
    assert ops.launch_count() > l2 + 100


def test_reference_train_inpaintnet_script_default_autoregressive_mode():
    """train_inpaintnet.py at its defaults (auto_reg=True, teacher_forcing=True: train_inpaintnet.py:53-56)."""
    _run("train_measure_vae.py", batch_size=4, num_epochs=1, train=True, plot=False, log=False, lr=1e-4, **_VAE_KW)
    _run("train_inpaintnet.py", batch_size=4, num_epochs=1, train=True, lr=1e-4, plot=False, log=False, auto_reg=True,
         teacher_forcing=True, early_stop=True, **_LAT_KW, **_VAE_KW)

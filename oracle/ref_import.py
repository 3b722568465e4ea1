"""Import harness for the UNMODIFIED reference (test infrastructure only).

Imports the reference from /root/reference when it is mounted (the build container), else from the byte-for-byte
staged copy oracle/_ref/ that oracle/make_ref.sh writes (git-ignored; it travels to the GPU box with the snapshot).
Used by tests/golden/make_golden.py to generate the committed golden fixtures, by the `-m "not gpu"` tests that
validate the oracle restatement, and by oracle/ref_bench.py (the CPU arm of bench.py).
Nothing on the product path imports this file.

Stubs the four packages the reference imports but this image lacks
(music21, glob2, tensorboard_logger, matplotlib) - SURVEY.md section 8(c).
"""
import os
import sys
import types
from unittest.mock import MagicMock

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    cands = [os.environ.get("INPAINTNET_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "MeasureVAE")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "MeasureVAE"))


class FakeMetadata:
    def __init__(self, num_values):
        self.num_values = num_values
        self.is_global = False


class FakeDataset:
    """Synthetic stand-in exposing the dataset attributes the hot path touches
    (reference: MeasureVAE/measure_vae.py:45,56; LatentRNN/latent_rnn_trainer.py:22-24;
    AnticipationRNN/anticipation_rnn_gauss_reg_model.py:70-87)."""

    def __init__(self, num_notes=64, n_bars=16):
        self.note2index_dicts = [{i: i for i in range(num_notes)}]
        self.index2note_dicts = [{i: i for i in range(num_notes)}]
        self.n_bars = n_bars
        self.subdivision = 6
        self.num_beats_per_bar = 4
        self.num_voices = 1
        self.NOTES = 0
        self.metadatas = [FakeMetadata(6), FakeMetadata(6)]
        self.seq_size_in_beats = n_bars * 4

    def __repr__(self):
        return f"FakeDataset({len(self.note2index_dicts[0])},{self.n_bars})"

    def empty_score_tensor(self, score_length):
        import torch
        return torch.zeros(self.num_voices, score_length).long()


_loaded = {}


def load_reference():
    """Returns a namespace with the reference classes. Idempotent."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    for m in ["music21", "music21.abcFormat", "music21.interval", "glob2", "tensorboard_logger",
              "matplotlib", "matplotlib.pyplot"]:
        if m not in sys.modules:
            sys.modules[m] = MagicMock()
    # the reference packages are called MeasureVAE/LatentRNN/utils/...; make sure OUR drop-in
    # packages of the same name are not already imported in this interpreter.
    for m in list(sys.modules):
        top = m.split(".")[0]
        if top in ("MeasureVAE", "LatentRNN", "AnticipationRNN", "utils", "DatasetManager"):
            f = getattr(sys.modules[m], "__file__", "") or ""
            if not f.startswith(REFERENCE_ROOT):
                raise RuntimeError(f"module {m} already imported from {f}; cannot load reference")
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        from MeasureVAE.measure_vae import MeasureVAE
        from MeasureVAE.vae_trainer import VAETrainer
        from LatentRNN.latent_rnn import LatentRNN
        from LatentRNN.latent_rnn_trainer import LatentRNNTrainer
        from AnticipationRNN.anticipation_rnn_gauss_reg_model import ConstraintModelGaussianReg
        from AnticipationRNN.anticipation_rnn_trainer import AnticipationRNNGaussianRegTrainer
        from utils.trainer import Trainer
    finally:
        sys.path.remove(REFERENCE_ROOT)
    _loaded.update(dict(MeasureVAE=MeasureVAE, VAETrainer=VAETrainer, LatentRNN=LatentRNN,
                        LatentRNNTrainer=LatentRNNTrainer,
                        ConstraintModelGaussianReg=ConstraintModelGaussianReg,
                        AnticipationRNNGaussianRegTrainer=AnticipationRNNGaussianRegTrainer,
                        Trainer=Trainer, FakeDataset=FakeDataset))
    return types.SimpleNamespace(**_loaded)

# The reference's test_reconstruction.py star-imports this module and then uses MeasureVAE, LatentRNN,
# ConstraintModelGaussianReg, AnticipationRNNBaseline, Trainer, torch and tqdm without importing them itself
# (MeasureVAE/vae_tester.py:1-14 re-exports part of that set; the rest is missing upstream): export them all here.
import torch  # noqa: F401
from inpaintnet_b200.tester import VAETester  # noqa: F401
from inpaintnet_b200.helpers import *  # noqa: F401,F403
from inpaintnet_b200.measure_vae import MeasureVAE  # noqa: F401
from inpaintnet_b200.latent_rnn import LatentRNN, LatentRNNAblations  # noqa: F401
from inpaintnet_b200.arnn import ConstraintModelGaussianReg, AnticipationRNNBaseline  # noqa: F401
from inpaintnet_b200.trainer import Trainer, VAETrainer, tqdm  # noqa: F401

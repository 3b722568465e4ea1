from inpaintnet_b200.latent_rnn import LatentRNNAblations  # noqa: F401

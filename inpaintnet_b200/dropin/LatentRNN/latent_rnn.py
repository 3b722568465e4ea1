from inpaintnet_b200.latent_rnn import LatentRNN  # noqa: F401

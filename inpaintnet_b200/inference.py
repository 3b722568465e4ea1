"""Batched inpainting inference with the launch sequence captured in a CUDA graph.

`LatentRNN.forward(past, future, target, n, train=False)` (LatentRNN/latent_rnn.py:110-159, the call the reference's
testers and generation scripts make) issues ~250 kernel launches per query batch: the context encode, two context
GRUs, the generation GRU and 24 serial decoder ticks.  Shapes are static for a fixed (batch, split), so the whole
sequence is captured ONCE and replayed with a single launch per batch -- on an 8-GPU box eight processes no longer
compete for host cores to issue launches (weak-scaling efficiency of the inference path).  Nothing about the
mathematics changes: the graph is a recording of exactly the eager call; the reparameterisation noise of the
context latents (the reference draws fresh `rsample` noise on every call, also in eval mode: latent_rnn.py:172) is
regenerated into static buffers by one Philox launch before every replay.
"""
import torch

from . import engine
from .arena import arena_of


class GraphedInpainter:
    """model: an eval-mode non-autoregressive LatentRNN on a CUDA device (the model the evaluation scripts load,
    test_reconstruction.py:141).  queries / (n_past, n_target, n_future): the fixed batch size and split.

        inp = GraphedInpainter(model, 8192, 6, 4, 6)
        weights, samples, gen_z = inp(score)        # score: (queries, n_bars, 24) or (queries, 1, n_bars * 24) int

    The returned tensors are the graph's static outputs: they are overwritten by the next call."""

    def __init__(self, model, queries, n_past, n_target, n_future, warmup=2):
        if model.auto_reg:
            raise NotImplementedError("GraphedInpainter captures the non-autoregressive generation path")
        if model.training:
            raise ValueError("GraphedInpainter needs model.eval() (no dropout masks are drawn inside the graph)")
        self.model = model
        self.split = (int(n_past), int(n_target), int(n_future))
        n_bars = sum(self.split)
        arena = arena_of(model)
        dev = arena.device
        if dev.type != "cuda":
            raise ValueError("GraphedInpainter needs the model on a CUDA device")
        Z = model.z_dim
        self.score = torch.zeros(queries, n_bars, 24, dtype=torch.int32, device=dev)     # static input
        self.eps = [torch.empty(n_past * queries, Z, dtype=torch.float32, device=dev),
                    torch.empty(n_future * queries, Z, dtype=torch.float32, device=dev)]  # static noise
        self._arena = arena
        self.refresh_noise()
        # warm-up and capture run on ONE side stream: the arena's caches (bf16 weight pack, derived tables) remember the
        # stream that filled them and make other streams wait on an event -- which a capturing stream must not do
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):          # eager warm-up: derived tables, weight pack, TMA descriptors, kernel attributes
            for _ in range(max(1, warmup)):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: other threads of the process (NCCL watchdog under torchrun, data loaders) may call CUDA meanwhile
        with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
            self.out = self._body()
        arena.range_flag.zero_()

    def _body(self):
        n_p, n_t, n_f = self.split
        s = self.score.long()
        past, target, future = s[:, :n_p], s[:, n_p:n_p + n_t], s[:, n_p + n_t:]
        with torch.no_grad(), engine.inject_noise(eps=list(self.eps)):
            return self.model(past, future, target, n_t, train=False)

    def refresh_noise(self):
        """Fresh N(0,1) draws for the context latents (one Philox launch per buffer, outside the graph)."""
        for e in self.eps:
            engine.NOISE.fill_normal(self._arena, e)

    def __call__(self, score, fresh_noise=True):
        """score: int tensor with queries * n_bars * 24 tokens, host (pinned for an asynchronous upload) or device."""
        self.score.copy_(score.reshape(self.score.shape), non_blocking=True)
        if fresh_noise:
            self.refresh_noise()
        self.graph.replay()
        return self.out

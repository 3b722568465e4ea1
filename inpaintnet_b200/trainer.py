"""Trainer layer -- same API as the reference's utils/trainer.py (Trainer ABC, train_model,
loss_and_acc_on_epoch, overridable zero_grad/step) and MeasureVAE/vae_trainer.py, with
  * the fused CE+KL forward/backward kernel instead of ~10 small torch kernels,
  * the fused flat Adam instead of the per-tensor optimiser loop,
  * an NCCL gradient all-reduce over the contiguous gradient arena when torch.distributed is
    initialised (one process per GPU; the reference has no data parallelism at all),
  * one device->host sync per step (the loss read the reference also does, utils/trainer.py:154);
    the NaN / token-range guards are device flags polled with that same sync.
"""
import os
import random
import time
import datetime
from abc import ABC, abstractmethod

import numpy as np
import torch
import torch.distributed as dist

from . import functional as Fn
from .arena import arena_of
from .helpers import to_numpy, to_cuda_variable_long
from .optim import FusedAdam

try:  # progress bars only
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(x):
        return x


def dp_rank_world():
    """(rank, world) of the data-parallel job; (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_batch(batch):
    """This rank's share of a loader batch under data parallelism.  Every rank iterates the SAME loader (rank-shared
    torch seed -> identical shuffle), so a loader batch is the global batch: rank r keeps rows r, r + world, ... of its
    first (B // world) * world rows (equal local batches, so the mean of the local losses is the global mean).
    Tensors that are None (the VAE trainer ignores the metadata) pass through."""
    rank, world = dp_rank_world()
    if world == 1:
        return batch
    def take(t):
        if t is None or not torch.is_tensor(t) or t.dim() == 0:
            return t
        n = t.shape[0] // world * world
        if n == 0:
            raise ValueError(f"data-parallel batch of {t.shape[0]} rows cannot be split over {world} ranks")
        return t[rank:n:world]
    return tuple(take(t) for t in batch) if isinstance(batch, (tuple, list)) else take(batch)


class LaggedReadback:
    """Per-step results (loss, accuracy, NaN flag, token-range flag) travel device -> pinned host as an async copy
    queued behind the step's kernels; the host reads step i-1's slot while step i is already running, so the
    one device->host read per step the reference does (utils/trainer.py:154) no longer drains the GPU."""

    def __init__(self, depth=4):
        self.depth = depth
        self.slots = None
        self.events = []
        self.pending = []      # slot indices in flight, oldest first
        self.n = 0

    def push(self, loss, accuracy, arena):
        if self.slots is None:
            pin = loss.is_cuda
            self.slots = torch.zeros(self.depth, 4, dtype=torch.float32, pin_memory=pin)
            self.events = [torch.cuda.Event() if pin else None for _ in range(self.depth)]
        if len(self.pending) == self.depth:
            raise RuntimeError("LaggedReadback: pop before pushing more than `depth` steps")
        slot = self.n % self.depth
        self.n += 1
        acc = accuracy if accuracy is not None else loss.new_zeros(())
        row = torch.stack((loss.detach().float().mean(), acc.detach().float().mean(), arena.nan_flag[0].float(),
                           arena.range_flag[0].float()))
        self.slots[slot].copy_(row, non_blocking=True)
        if self.events[slot] is not None:
            self.events[slot].record()
        self.pending.append(slot)

    def pop(self, keep=1):
        """Yields (loss, accuracy) of every queued step but the newest `keep`; raises on a raised device flag."""
        out = []
        while len(self.pending) > keep:
            slot = self.pending.pop(0)
            if self.events[slot] is not None:
                self.events[slot].synchronize()
            loss, acc, nan_f, range_f = self.slots[slot].tolist()
            if nan_f != 0:
                print('Model parameters have become nan')
                raise ValueError
            if range_f != 0:
                print("Invalid Values of Indices")
                raise ValueError
            out.append((loss, acc))
        return out


class Trainer(ABC):
    def __init__(self, dataset, model, lr=1e-4, early_stopping=False):
        self.dataset = dataset
        self.model = model
        self.optimizer = FusedAdam(self.model, lr=lr)
        self.early_stopping = False
        if early_stopping:
            self.early_stopping = True
            self.early_stopper = EarlyStopping()
        self.check_flags_every = 1
        self.microbatches = 1
        self._mb_streams = []
        self._mb_active = 0

    # ---- true resume (absent in the reference): everything a bit-for-bit
    # continuation needs beyond the weights -- Adam moments + step, epoch, early-stopping state and the RNG streams
    # (torch seed + the arena's Philox offset drive dropout / eps on the device; python `random` drives the
    # teacher-forcing coin; the torch CPU generator drives the past/gap/future split and the loader shuffle).
    def training_state_path(self):
        return self.model.filepath + '.train_state'

    def training_state(self, epoch_index):
        a = arena_of(self.model)
        st = dict(epoch=int(epoch_index), model=self.model.state_dict(), optimizer=self.optimizer.state_dict(),
                  rng_offset=int(a.rng_offset), torch_seed=int(torch.initial_seed()), torch_rng=torch.get_rng_state(),
                  python_rng=random.getstate())
        if self.early_stopping:
            es = self.early_stopper
            st["early_stopping"] = dict(counter=es.counter, best_score=es.best_score, early_stop=es.early_stop,
                                        val_loss_min=es.val_loss_min)
        return st

    def save_training_state(self, epoch_index, path=None):
        path = path or self.training_state_path()
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        tmp = path + ".tmp"
        torch.save(self.training_state(epoch_index), tmp)
        os.replace(tmp, path)          # a killed job never leaves a half-written state behind
        return path

    def load_training_state(self, path=None):
        """Restores a state written by save_training_state; returns the index of the next epoch to run."""
        st = torch.load(path or self.training_state_path(), map_location="cpu", weights_only=False)
        self.model.load_state_dict(st["model"])
        self.optimizer.load_state_dict(st["optimizer"])
        torch.manual_seed(st["torch_seed"])
        torch.set_rng_state(st["torch_rng"])
        random.setstate(st["python_rng"])
        arena_of(self.model).rng_offset = st["rng_offset"]
        if self.early_stopping and "early_stopping" in st:
            for k, v in st["early_stopping"].items():
                setattr(self.early_stopper, k, v)
        return st["epoch"] + 1

    def train_model(self, batch_size, num_epochs, plot=False, log=False, resume=False):
        if log:
            try:
                from tensorboard_logger import configure, log_value
                st = datetime.datetime.fromtimestamp(time.time()).strftime('%Y-%m-%d_%H:%M:%S')
                configure(os.path.join('runs/' + self.model.__repr__() + st))
            except ImportError:
                print("tensorboard_logger not installed: logging disabled")
                log = False
        if plot:
            print("live plotting is not part of the B200 hot path: disabled")
            plot = False
        (generator_train, generator_val, _) = self.dataset.data_loaders(batch_size=batch_size, split=(0.70, 0.20))
        print('Num Train Batches: ', len(generator_train))
        print('Num Valid Batches: ', len(generator_val))
        first_epoch = 0
        if resume and os.path.exists(self.training_state_path()):
            first_epoch = self.load_training_state()
            if self.early_stopping and self.early_stopper.early_stop:
                print("Early Stopping")      # the saved run had already stopped: nothing left to train
                return
            print(f'Resuming at epoch {first_epoch + 1}/{num_epochs}')
        for epoch_index in range(first_epoch, num_epochs):
            self.update_scheduler(epoch_index)
            self.model.train()
            mean_loss_train, mean_accuracy_train = self.loss_and_acc_on_epoch(
                data_loader=generator_train, epoch_num=epoch_index, train=True)
            self.model.eval()
            mean_loss_val, mean_accuracy_val = self.loss_and_acc_on_epoch(
                data_loader=generator_val, epoch_num=epoch_index, train=False)
            data_element = {'epoch_index': epoch_index, 'num_epochs': num_epochs, 'mean_loss_train': mean_loss_train,
                            'mean_accuracy_train': mean_accuracy_train, 'mean_loss_val': mean_loss_val,
                            'mean_accuracy_val': mean_accuracy_val}
            if log:
                log_value('train_loss', mean_loss_train, epoch_index)
                log_value('train_accu', mean_accuracy_train, epoch_index)
                log_value('valid_loss', mean_loss_val, epoch_index)
                log_value('valid_accu', mean_accuracy_val, epoch_index)
            self.print_epoch_stats(**data_element)
            if self.early_stopping:   # before the state is saved, so that a resumed run sees this epoch's verdict
                self.early_stopper(mean_loss_val, self.model)
            if not dist.is_initialized() or dist.get_rank() == 0:
                self.model.save()
                if epoch_index > 0 and epoch_index % 10 == 0:
                    self.model.save_checkpoint(epoch_index)
                self.save_training_state(epoch_index)
            if self.early_stopping and self.early_stopper.early_stop:
                print("Early Stopping")
                return

    def run_batch(self, batch, epoch_num=None, train=True, readback=None, shard=False):
        """One step on a host batch from the loader: upload (async from pinned memory), forward, and for train=True
        backward + gradient exchange + Adam.  With `readback` the results are queued for a lagged host read
        (no sync here); without it they are returned as device tensors.  shard=True: `batch` is the GLOBAL batch of a
        data-parallel job and this rank trains on its share (shard_batch); callers that already hold per-rank data
        (bench.py, the tests) leave it False."""
        if shard:
            batch = shard_batch(batch)
        batch_data = self.process_batch_data(batch)
        self.zero_grad()
        if train:
            loss, accuracy = self.loss_and_acc_for_batch(batch_data, epoch_num, train=train)
            loss.backward()
            self.step()
        else:
            with torch.no_grad():
                loss, accuracy = self.loss_and_acc_for_batch(batch_data, epoch_num, train=train)
        if readback is not None:
            readback.push(loss, accuracy, arena_of(self.model))
        return loss, accuracy

    def loss_and_acc_on_epoch(self, data_loader, epoch_num=None, train=True):
        """utils/trainer.py:126-163.  The per-step loss / accuracy / guard-flag read lags one step behind the
        launches (LaggedReadback), so the NaN and token-range guards fire one step late but the GPU never idles."""
        mean_loss = 0
        mean_accuracy = 0
        rb = LaggedReadback()
        for sample_id, batch in tqdm(enumerate(data_loader)):
            self.run_batch(batch, epoch_num, train, readback=rb, shard=True)
            for loss, acc in rb.pop(keep=1):
                mean_loss += loss
                mean_accuracy += acc
        for loss, acc in rb.pop(keep=0):
            mean_loss += loss
            mean_accuracy += acc
        mean_loss /= len(data_loader)
        mean_accuracy /= len(data_loader)
        rank, world = dp_rank_world()
        if world > 1:
            # every rank saw a different shard: the epoch statistics (and therefore the early-stopping decision, which
            # must be the same everywhere or the ranks' collectives diverge) are the mean over ranks
            stats = torch.tensor([mean_loss, mean_accuracy], dtype=torch.float64, device=arena_of(self.model).device)
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
            mean_loss, mean_accuracy = (stats / world).tolist()
        return (mean_loss, mean_accuracy)

    def check_device_flags(self):
        """NaN-in-parameters (encoder.py:111-116) and token-range (decoder.py:34-45) guards -> ValueError."""
        a = arena_of(self.model)
        flags = torch.stack((a.nan_flag[0], a.range_flag[0])).cpu()
        if int(flags[0]) != 0:
            print('Model parameters have become nan')
            raise ValueError
        if int(flags[1]) != 0:
            print("Invalid Values of Indices")
            raise ValueError

    def zero_grad(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.grad_exchange()   # hook armed before the backward pass of this step
        self.optimizer.zero_grad()

    def step(self):
        """Gradient all-reduce (sum over ranks, scaled by 1/world inside the Adam kernel) + fused Adam."""
        scale = 1.0
        from .engine import SIDE
        SIDE.join()   # weight-gradient GEMMs deferred to the side stream (normally already joined by the autograd callback)
        if self._mb_active:   # backward nodes of the micro-batches ran on their own streams
            cur = torch.cuda.current_stream()
            for st in self._mb_streams[: self._mb_active]:
                cur.wait_stream(st)
            self._mb_active = 0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.grad_exchange().finish(arena_of(self.model))   # buckets not already reduced under the backward pass
            scale = 1.0 / dist.get_world_size()
        self.optimizer.step(grad_scale=scale)

    def grad_exchange(self):
        """The overlapped gradient exchange of this trainer; installs the backward-pass hook on first use."""
        from . import engine
        if getattr(self, "_grad_exchange", None) is None:
            self._grad_exchange = GradExchange()
        engine.GRAD_READY = self._grad_exchange.ready
        return self._grad_exchange

    @abstractmethod
    def loss_and_acc_for_batch(self, batch, epoch_num=None, train=True):
        pass

    @abstractmethod
    def process_batch_data(self, batch):
        pass

    @abstractmethod
    def update_scheduler(self, epoch_num):
        pass

    @staticmethod
    def print_epoch_stats(epoch_index, num_epochs, mean_loss_train, mean_accuracy_train, mean_loss_val,
                          mean_accuracy_val):
        print(f'Train Epoch: {epoch_index + 1}/{num_epochs}')
        print(f'\tTrain Loss: {mean_loss_train}'
              f'\tTrain Accuracy: {mean_accuracy_train * 100} %')
        print(f'\tValid Loss: {mean_loss_val}'
              f'\tValid Accuracy: {mean_accuracy_val * 100} %')

    @staticmethod
    def mean_crossentropy_loss(weights, targets):
        """weights (batch, seq_len, num_notes), targets (batch, seq_len) -> scalar (utils/trainer.py:271-288)"""
        batch_size, seq_len, num_notes = weights.size()
        assert (batch_size == targets.size(0))
        assert (seq_len == targets.size(1))
        return Fn.fused_ce_kl(weights, targets)[0]

    @staticmethod
    def mean_accuracy(weights, targets):
        """utils/trainer.py:290-306 (first maximal index == target)"""
        with torch.no_grad():
            return Fn.fused_ce_kl(weights.detach(), targets)[1]

    @staticmethod
    def mean_crossentropy_loss_alt(weights, targets):
        """weights (batch, num_measures, seq_len, num_notes) (utils/trainer.py:344-358)"""
        return Fn.fused_ce_kl(weights, targets)[0]

    @staticmethod
    def mean_accuracy_alt(weights, targets):
        with torch.no_grad():
            return Fn.fused_ce_kl(weights.detach(), targets)[1]


class GradExchange:
    """Bucketed gradient all-reduce OVERLAPPED with the backward pass (one process per GPU, NCCL over NVLink).
    The backward passes announce parameter groups whose gradients are complete (engine.grad_ready: the decoder
    before the encoder backward starts, the generation GRU before the context GRUs); their buckets are reduced on
    a communication stream while the remaining layers run.  finish() reduces whatever was not announced and makes
    the current stream wait for every bucket.  All ranks run the same kernels in the same order (rank-shared
    seeds), so the collectives are issued in the same order everywhere."""

    def __init__(self, bucket_bytes=32 << 20):
        self.bucket = max(1, bucket_bytes // 4)
        self.comm = None
        self.works = []
        self.done = []          # [lo, hi) ranges already queued this step
        self.n_early = 0        # buckets queued from inside the backward pass (diagnostics / tests)
        self.overlap = os.environ.get("IPN_DP_OVERLAP", "1") != "0"   # 0: everything is reduced at finish()
        self.hold = False       # set for a step whose gradients arrive from several micro-batches

    def _queue(self, arena, lo, hi):
        for a in range(lo, hi, self.bucket):
            self.works.append(dist.all_reduce(arena.grad[a:min(hi, a + self.bucket)], op=dist.ReduceOp.SUM, async_op=True))

    def ready(self, arena, prefixes, side_stream=None):
        r = arena.trainable_range(prefixes) if self.overlap and not self.hold else None
        if r is None or any(lo < r[1] and r[0] < hi for lo, hi in self.done):
            return
        if arena.grad.is_cuda:
            if self.comm is None:
                self.comm = torch.cuda.Stream()
            self.comm.wait_stream(torch.cuda.current_stream())
            if side_stream is not None:
                self.comm.wait_stream(side_stream)
            with torch.cuda.stream(self.comm):
                n0 = len(self.works)
                self._queue(arena, *r)
        else:
            n0 = len(self.works)
            self._queue(arena, *r)
        self.n_early += len(self.works) - n0
        self.done.append(r)

    def finish(self, arena):
        pos = 0
        for lo, hi in sorted(self.done) + [(arena.n_trainable, arena.n_trainable)]:
            if pos < lo:
                self._queue(arena, pos, lo)
            pos = max(pos, hi)
        for w in self.works:
            w.wait()
        self.works, self.done, self.hold = [], [], False


def allreduce_grads(arena, bucket_bytes=32 << 20):
    """Sum-all-reduce of the trainable gradient range in buckets (NCCL over NVLink when the backend is
    nccl; gloo on CPU for the tests).  The range is contiguous, so buckets are plain slices."""
    n = arena.n_trainable
    step = max(1, bucket_bytes // 4)
    works = []
    for lo in range(0, n, step):
        works.append(dist.all_reduce(arena.grad[lo:min(n, lo + step)], op=dist.ReduceOp.SUM, async_op=True))
    for w in works:
        w.wait()


class VAETrainer(Trainer):
    """reference: MeasureVAE/vae_trainer.py:10-139"""

    def __init__(self, dataset, model, lr=1e-4):
        super(VAETrainer, self).__init__(dataset, model, lr)
        # micro-batches pipelined on streams (see _loss_and_acc_pipelined); 1 = the whole batch on one stream
        self.microbatches = int(os.environ.get("IPN_MICROBATCHES", "1"))

    def loss_and_acc_for_batch(self, batch, epoch_num=None, train=True):
        score = batch
        n = self.microbatches
        if n > 1 and train and score.is_cuda and torch.is_grad_enabled() and score.shape[0] % n == 0:
            return self._loss_and_acc_pipelined(score, n)
        return self._loss_and_acc(score, train)

    def _loss_and_acc(self, score, train):
        weights, samples, z_dist, prior_dist, z_tilde, z_prior = self.model(measure_score_tensor=score, train=train)
        # recons_loss + 0.001 * KL and accuracy in one fused pass (vae_trainer.py:33-39)
        log_std = getattr(z_dist, "log_std", None)
        if log_std is None:
            log_std = z_dist.scale.log()
        loss, accuracy = Fn.fused_ce_kl(weights, score, z_dist.loc, log_std, beta=0.001)
        return loss, accuracy

    def _loss_and_acc_pipelined(self, score, n):
        """The batch runs as n equal micro-batches on n streams (forward here; autograd runs every backward node
        on its forward stream).  The serial GRU chain kernels occupy one SM per 128-row tile and are latency bound
        (a 2048-measure step takes 9.1 ms, a 4096-measure one 12.1 ms), so one micro-batch's chain kernels run
        next to the other's GEMMs instead of leaving most of the chip idle.  Same mathematics: equal micro-batches
        make the mean of the per-micro-batch losses the batch mean, gradients accumulate in the shared arena, and
        ONE teacher-forcing coin is drawn for the whole batch as the reference does (decoder.py:412-453)."""
        dec = self.model.decoder
        coin = (random.random() < dec.teacher_forcing_prob) if dec.use_teacher_forcing else False
        saved_prob = dec.teacher_forcing_prob
        main = torch.cuda.current_stream()
        if len(self._mb_streams) < n:
            self._mb_streams += [torch.cuda.Stream() for _ in range(n - len(self._mb_streams))]
        if getattr(self, "_grad_exchange", None) is not None:
            self._grad_exchange.hold = True      # a parameter group is complete only after ALL micro-batches
        losses, accs = [], []
        py_rng = random.getstate()               # the per-call coins below are forced: keep the stream as one draw
        dec.teacher_forcing_prob = 2.0 if coin else -1.0
        try:
            for k, part in enumerate(score.chunk(n)):
                st = self._mb_streams[k]
                st.wait_stream(main)
                part.record_stream(st)
                with torch.cuda.stream(st):
                    loss_k, acc_k = self._loss_and_acc(part, True)
                losses.append(loss_k)
                accs.append(acc_k)
        finally:
            dec.teacher_forcing_prob = saved_prob
            random.setstate(py_rng)
        for k in range(n):
            main.wait_stream(self._mb_streams[k])
            losses[k].record_stream(main)
            accs[k].record_stream(main)
        self._mb_active = n
        return torch.stack(losses).mean(), torch.stack(accs).mean()

    def process_batch_data(self, batch):
        score_tensor, _ = batch
        if hasattr(self.dataset, "n_bars") and score_tensor.dim() == 3:
            batch_size = score_tensor.size(0)
            score_tensor = score_tensor.view(batch_size, self.dataset.n_bars, -1)
            score_tensor = score_tensor.view(batch_size * self.dataset.n_bars, -1)
        return to_cuda_variable_long(score_tensor)

    def update_scheduler(self, epoch_num):
        return

    @staticmethod
    def compute_kld_loss(z_dist, prior_dist, beta=0.001):
        """vae_trainer.py:128-139, generic torch form (the step loop uses the fused kernel instead)."""
        kld = torch.distributions.kl.kl_divergence(z_dist, prior_dist)
        return beta * kld.sum(1).mean()


class LatentRNNTrainer(Trainer):
    """reference: LatentRNN/latent_rnn_trainer.py:8-176"""

    def __init__(self, dataset, model, lr=1e-4, early_stopping=False):
        super(LatentRNNTrainer, self).__init__(dataset, model, lr, early_stopping)
        self.min_num_measures_target = 2
        self.max_num_measure_target = 6
        assert (self.max_num_measure_target >= self.min_num_measures_target)
        assert (self.dataset.n_bars > self.min_num_measures_target)
        assert (self.dataset.n_bars > self.max_num_measure_target)
        self.measure_seq_len = self.dataset.subdivision * self.dataset.num_beats_per_bar

    def process_batch_data(self, batch):
        score_tensor, _ = batch
        return self.split_score_stochastic(score_tensor)

    def loss_and_acc_for_batch(self, batch, epoch_num=None, train=True):
        tensor_past, tensor_future, tensor_target = batch
        num_measures_past = tensor_past.size(1)
        num_measures_future = tensor_future.size(1)
        weights, pred, _ = self.model(past_context=tensor_past, future_context=tensor_future, target=tensor_target,
                                      measures_to_generate=self.dataset.n_bars - num_measures_past - num_measures_future,
                                      train=train)
        loss, accuracy = Fn.fused_ce_kl(weights, tensor_target)   # latent_rnn_trainer.py:58-66 in one pass
        return loss, accuracy

    def update_scheduler(self, epoch_num):
        return

    def draw_split(self, num_measures, fix_num_target=None):
        """(n_past, n_target, n_future) of one batch.  The two host draws, their bounds and their ORDER are those of
        latent_rnn_trainer.py:99-118 (gap length first, then the past length), so a seeded run splits exactly like
        the reference; under data parallelism every rank must hold the same torch seed (same shapes, same kernels)."""
        lo, hi = self.min_num_measures_target, self.max_num_measure_target
        n_target = fix_num_target if fix_num_target is not None else int(torch.randint(lo, hi + 1, (1,)))
        n_past = int(torch.randint(1, num_measures - n_target - 1, (1,)))     # leaves at least two future measures
        return n_past, n_target, num_measures - n_past - n_target

    def split_score_stochastic(self, score_tensor, extra_outs=False, fix_num_target=None):
        """latent_rnn_trainer.py:77-132: a random past / gap / future partition of the (B, 1, n_bars * 24) score."""
        n_bars = score_tensor.shape[-1] // self.measure_seq_len
        assert n_bars == self.dataset.n_bars
        n_past, n_target, n_future = self.draw_split(n_bars, fix_num_target)
        parts = LatentRNNTrainer.split_score(score_tensor, n_past, n_future, n_target, self.measure_seq_len)
        return parts + (n_past, n_target) if extra_outs else parts

    @staticmethod
    def split_score(score_tensor, num_past, num_future, num_target, measure_seq_len):
        """latent_rnn_trainer.py:134-160 -> (past, future, target) int64 device tensors (B, n, 24).  ONE upload of the
        whole score (asynchronous when the loader pinned it), the three blocks are views of the device copy."""
        measures = LatentRNNTrainer.split_to_measures(score_tensor, measure_seq_len)
        if measures.shape[1] != num_past + num_target + num_future:
            raise AssertionError((measures.shape[1], num_past, num_target, num_future))
        dev = to_cuda_variable_long(measures)
        gap_end = num_past + num_target
        return dev[:, :num_past], dev[:, gap_end:], dev[:, num_past:gap_end]

    @staticmethod
    def split_to_measures(score_tensor, measure_seq_len):
        """(B, 1, L) -> (B, L / measure_seq_len, measure_seq_len); ValueError when L is not a whole number of measures
        (latent_rnn_trainer.py:162-176)."""
        if score_tensor.shape[-1] % measure_seq_len:
            raise ValueError
        return score_tensor.reshape(score_tensor.shape[0], -1, measure_seq_len)


class EarlyStopping:
    """Patience counter on the validation loss (behaviour of utils/trainer.py:379-413: an epoch counts as an
    improvement only when the loss drops by at least `min_delta` below the best one seen; `patience` epochs in a row
    without one raise `early_stop`).  Its four state fields are part of the saved training state."""
    min_delta = 1e-5

    def __init__(self, patience=5, verbose=False):
        self.patience, self.verbose = patience, verbose
        self.counter, self.best_score, self.early_stop, self.val_loss_min = 0, None, False, np.inf

    def __call__(self, val_loss, model):
        score = -val_loss
        if self.best_score is None:            # first epoch: the reference records the score, not the loss
            self.best_score = score
            return
        if score - self.best_score >= self.min_delta:
            self.best_score, self.val_loss_min, self.counter = score, val_loss, 0
            return
        self.counter += 1
        self.early_stop = self.early_stop or self.counter >= self.patience

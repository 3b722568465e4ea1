"""Held-out loss / accuracy loops of the reference testers (MeasureVAE/vae_tester.py:34-49,114-155;
LatentRNN/latent_rnn_tester.py:28-50,297-340) and the inpainting call that returns scores
(latent_rnn_tester.py:191-262; scores are `score.Score` objects with MIDI export instead of music21 streams).
Plotting and t-SNE are CPU post-processing outside the hot path and are not provided."""
import torch

from . import functional as Fn
from .helpers import to_cuda_variable_long, to_numpy
from .trainer import LatentRNNTrainer


class VAETester(object):
    def __init__(self, dataset, model):
        self.dataset = dataset
        self.model = model
        self.model.eval()
        self.decoder = self.model.decoder
        self.z_dim = self.decoder.z_dim
        self.batch_size = 1
        self.measure_seq_len = 24

    def test_model(self, batch_size=64):
        (_, gen_val, gen_test) = self.dataset.data_loaders(batch_size=batch_size, split=(0.01, 0.01))
        print('Num Test Batches: ', len(gen_test))
        mean_loss_test, mean_accuracy_test = self.loss_and_acc_test(gen_test)
        print('Test Epoch:')
        print('\tTest Loss: ', mean_loss_test, '\n\tTest Accuracy: ', mean_accuracy_test * 100)
        return mean_loss_test, mean_accuracy_test

    def loss_and_acc_test(self, data_loader):
        mean_loss = 0
        mean_accuracy = 0
        n = 0
        for sample_id, (score_tensor, metadata_tensor) in enumerate(data_loader):
            if hasattr(self.dataset, "n_bars") and score_tensor.dim() == 3:
                batch_size = score_tensor.size(0)
                score_tensor = score_tensor.view(batch_size * self.dataset.n_bars, -1)
            score_tensor = to_cuda_variable_long(score_tensor)
            with torch.no_grad():
                weights, samples, _, _, _, _ = self.model(measure_score_tensor=score_tensor, train=False)
                loss, accuracy = Fn.fused_ce_kl(weights, score_tensor)
            mean_loss += to_numpy(loss.mean())
            mean_accuracy += to_numpy(accuracy)
            n += 1
        n = max(n, 1)
        return mean_loss / n, mean_accuracy / n


class LatentRNNTester(object):
    def __init__(self, dataset, model):
        self.dataset = dataset
        self.model = model
        self.model.eval()
        self.measure_seq_len = 24
        self._split = LatentRNNTrainer.__new__(LatentRNNTrainer)
        self._split.dataset = dataset
        self._split.min_num_measures_target, self._split.max_num_measure_target = 2, 6
        self._split.measure_seq_len = 24

    def split_score_stochastic(self, score_tensor, extra_outs=False, fix_num_target=None):
        """latent_rnn_tester.py:359-414: as the trainer's split, but the past context may be as short as one
        measure (low=1) and at least one future measure is kept."""
        measures = LatentRNNTrainer.split_to_measures(score_tensor, self.measure_seq_len)
        n = measures.size(1)
        assert n == self.dataset.n_bars
        if fix_num_target is None:
            num_target = int(torch.randint(low=self._split.min_num_measures_target,
                                           high=self._split.max_num_measure_target + 1, size=(1,)).item())
        else:
            num_target = fix_num_target
        num_past = int(torch.randint(low=1, high=n - num_target - 1, size=(1,)).item())
        num_future = n - num_past - num_target
        past, future, target = LatentRNNTrainer.split_score(score_tensor=score_tensor, num_past=num_past,
                                                            num_future=num_future, num_target=num_target,
                                                            measure_seq_len=self.measure_seq_len)
        if extra_outs:
            return past, future, target, num_past, num_target
        return past, future, target

    def generate(self, tensor_past, tensor_future, tensor_target, num_target_measures=None, eval=False):
        """latent_rnn_tester.py:191-262: inpaint `num_target_measures` measures between the two contexts
        ((B, n, 24) token tensors) and return (generated score, generated token tensor (B, n_total, 24), original score
        or None).  Scores come from `dataset.tensor_to_score` (first sequence of the batch laid out in time, as the
        reference's flatten does for its batch of one)."""
        if tensor_target is not None:
            assert num_target_measures in (None, tensor_target.size(1))
            num_target_measures = tensor_target.size(1)
        elif num_target_measures is None:
            raise ValueError("the number of measures to generate is needed when there is no target")
        past, future = to_cuda_variable_long(tensor_past), to_cuda_variable_long(tensor_future)
        target = to_cuda_variable_long(tensor_target) if tensor_target is not None else \
            torch.zeros(past.size(0), num_target_measures, self.measure_seq_len, dtype=torch.long, device=past.device)
        with torch.no_grad():
            weights, gen_target, _ = self.model(past_context=past, future_context=future, target=target,
                                                measures_to_generate=num_target_measures, train=False)
            if tensor_target is not None and eval:
                loss, acc = Fn.fused_ce_kl(weights, target)
                print('Accuracy for Test Case:')
                print(f'\tLoss: {to_numpy(loss.mean())}\tAccuracy: {to_numpy(acc) * 100} %')
        gen_target = gen_target.view(past.size(0), num_target_measures, self.measure_seq_len)
        gen_score_tensor = torch.cat((past, gen_target, future), 1)
        gen_score = self.dataset.tensor_to_score(gen_score_tensor[:1].cpu())
        original_score = None
        if tensor_target is not None:
            original_score = self.dataset.tensor_to_score(torch.cat((past, target, future), 1)[:1].cpu())
        return gen_score, gen_score_tensor, original_score

    def test_model(self, batch_size=64):
        (_, _, gen_test) = self.dataset.data_loaders(batch_size=batch_size, split=(0.01, 0.01))
        mean_loss, mean_acc, n = 0, 0, 0
        for score_tensor, _ in gen_test:
            past, future, target = self.split_score_stochastic(score_tensor)
            with torch.no_grad():
                weights, _, _ = self.model(past_context=past, future_context=future, target=target,
                                           measures_to_generate=target.size(1), train=False)
                loss, acc = Fn.fused_ce_kl(weights, target)
            mean_loss += to_numpy(loss.mean())
            mean_acc += to_numpy(acc)
            n += 1
        n = max(n, 1)
        print('Test Epoch:')
        print('\tTest Loss: ', mean_loss / n, '\n\tTest Accuracy: ', mean_acc / n * 100)
        return mean_loss / n, mean_acc / n


class AnticipationRNNTester(object):
    """Held-out inpainting loss / accuracy of the AnticipationRNN baseline
    (AnticipationRNN/anticipation_rnn_tester.py:20-86,245-356): every test batch is inpainted over a fixed gap
    (2 measures starting at measure 8) with ConstraintModelGaussianReg.forward_inpaint."""

    def __init__(self, dataset, model):
        self.dataset = dataset
        self.model = model
        self.model.eval()
        self.batch_size = 1
        self.measure_seq_len = 24

    def test_model(self, batch_size=512):
        (_, _, gen_test) = self.dataset.data_loaders(batch_size=batch_size, split=(0.01, 0.01))
        print('Num Test Batches: ', len(gen_test))
        mean_loss_test, mean_accuracy_test = self.loss_and_acc_test(gen_test)
        print('Test Epoch: 1/1')
        print(f'\tTest Loss: {mean_loss_test}\tTest Accuracy: {mean_accuracy_test * 100} %')
        return mean_loss_test, mean_accuracy_test

    def loss_and_acc_test(self, data_loader):
        mean_loss, mean_accuracy, n = 0, 0, 0
        for batch in data_loader:
            score, metadata, constraints_loc, start_tick, end_tick = self.process_batch_data(batch)
            weights, _ = self.model.forward_inpaint(score_tensor=score, metadata_tensor=metadata,
                                                    constraints_loc=constraints_loc, start_tick=start_tick,
                                                    end_tick=end_tick)
            targets = score[:, :, start_tick:end_tick].transpose(0, 1)      # (num_voices, batch, gap ticks)
            mean_loss += to_numpy(self.mean_crossentropy_loss(weights, targets))
            mean_accuracy += to_numpy(self.mean_accuracy(weights, targets))
            n += 1
        n = max(n, 1)
        return mean_loss / n, mean_accuracy / n

    def process_batch_data(self, batch):
        tensor_score, tensor_metadata = batch
        tensor_score = to_cuda_variable_long(tensor_score)
        tensor_metadata = to_cuda_variable_long(tensor_metadata)
        constraints_location, start_tick, end_tick = self.get_constraints_location(tensor_score, is_stochastic=False)
        return tensor_score, tensor_metadata, constraints_location, start_tick, end_tick

    def get_constraints_location(self, tensor_score, is_stochastic, start_measure=None, num_measures=None):
        """-> constraints (1 = given, 0 = to inpaint) shaped like tensor_score, start_tick, end_tick.
        Stochastic: gap of 2..n-11 measures with at least 5 measures of context on both sides
        (anticipation_rnn_tester.py:278-303); otherwise measures [8, 10) unless told otherwise (:304-308)."""
        ticks = self.dataset.subdivision * self.dataset.num_beats_per_bar
        if is_stochastic:
            n = int(tensor_score.size(2) / ticks)
            assert n == self.dataset.n_bars
            num_target = int(torch.randint(low=2, high=n - 10, size=(1,)).item())
            num_past = int(torch.randint(low=5, high=n - num_target - 5, size=(1,)).item())
            assert n - num_past - num_target >= 5
            start_measure, num_measures = num_past + 1, num_target
        else:
            start_measure = 8 if start_measure is None else start_measure
            num_measures = 2 if num_measures is None else num_measures
        start_tick = start_measure * ticks
        end_tick = start_tick + num_measures * ticks
        constraints_location = torch.zeros_like(tensor_score)
        if start_tick > 0:
            constraints_location[:, :, :start_tick] = 1
        if end_tick < constraints_location.size(2) - 1:
            constraints_location[:, :, end_tick:] = 1
        return constraints_location, start_tick, end_tick

    @staticmethod
    def mean_crossentropy_loss(weights, targets):
        """weights: list (per voice) of (batch, seq_len, num_notes); targets (num_voices, batch, seq_len)"""
        total = 0
        for i, w in enumerate(weights):
            total = total + Fn.fused_ce_kl(w, targets[i])[0]
        return total / len(weights)

    @staticmethod
    def mean_accuracy(weights, targets):
        total = 0
        with torch.no_grad():
            for i, w in enumerate(weights):
                total = total + Fn.fused_ce_kl(w.detach(), targets[i])[1]
        return total / len(weights)

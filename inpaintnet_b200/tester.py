"""Held-out loss / accuracy loops of the reference testers (MeasureVAE/vae_tester.py:34-49,114-155;
LatentRNN/latent_rnn_tester.py:28-50,297-340).  Plotting, t-SNE and music21 score export are CPU
post-processing outside the hot path and are not provided."""
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

    def test_model(self, batch_size=64):
        (_, _, gen_test) = self.dataset.data_loaders(batch_size=batch_size, split=(0.01, 0.01))
        mean_loss, mean_acc, n = 0, 0, 0
        for score_tensor, _ in gen_test:
            past, future, target = self._split.split_score_stochastic(score_tensor)
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

"""reference: utils/helpers.py:5-49 (device convention: everything goes to the current CUDA device)."""
import torch
from torch.autograd import Variable  # noqa: F401  (re-exported: the reference star-imports it)


def to_cuda_variable(tensor):
    if torch.cuda.is_available():
        return tensor.cuda()
    return tensor


def to_cuda_variable_long(tensor):
    if torch.cuda.is_available():
        if tensor.device.type == "cpu" and tensor.dtype != torch.int64 and tensor.is_pinned():
            # pinned host batch (the loaders pin): upload the narrow type asynchronously, widen on the device
            return tensor.cuda(non_blocking=True).long()
        return tensor.long().cuda()
    return tensor.long()


def to_numpy(variable):
    if torch.cuda.is_available():
        return variable.data.cpu().numpy()
    return variable.data.numpy()


def init_hidden_lstm(num_layers, batch_size, lstm_hidden_size):
    hidden = (to_cuda_variable(torch.zeros(num_layers, batch_size, lstm_hidden_size)),
              to_cuda_variable(torch.zeros(num_layers, batch_size, lstm_hidden_size)))
    return hidden

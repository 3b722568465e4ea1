"""Checkpoint plumbing shared by the models.  API of the reference's utils/model.py:5-53 (`save`, `save_checkpoint`,
`load(cpu=False)`, the `filepath` attribute): a checkpoint is `torch.save(state_dict)` at `filepath` (per-epoch copies
at `filepath_<epoch>`), and the state_dict layout (names, shapes, fp32) is the reference's, so files written by
either implementation load in the other.  Optimiser / RNG state for a true resume lives in
`Trainer.save_training_state`, next to this file."""
import os

import torch


class Model(torch.nn.Module):
    def __init__(self):
        super(Model, self).__init__()
        self.filepath = None

    def forward(self):
        pass

    def _write_state(self, path):
        folder = os.path.dirname(path)
        if folder:
            os.makedirs(folder, exist_ok=True)
        tmp = path + ".tmp"
        torch.save(self.state_dict(), tmp)
        os.replace(tmp, path)      # never leave a truncated checkpoint behind

    def save(self):
        self._write_state(self.filepath)
        print(f'Model {self.__repr__()} saved')

    def save_checkpoint(self, epoch_num):
        self._write_state(f'{self.filepath}_{epoch_num}')
        print(f'Model checkpoint {self.__repr__()} saved for epoch')

    def load(self, cpu=False):
        # parameters are copied INTO the arena views, so the target device is the model's own either way;
        # `cpu=True` (reference: map_location to host storage) only decides where the file is staged
        state = torch.load(self.filepath, map_location="cpu" if cpu else None)
        self.load_state_dict(state)
        print(f'Model {self.__repr__()} loaded')

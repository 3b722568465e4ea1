"""reference: utils/model.py:5-53 -- save / save_checkpoint / load = torch.save(state_dict) at `filepath`.
The state_dict layout (names, shapes, fp32) is the reference's, so checkpoints are interchangeable."""
import os

import torch


class Model(torch.nn.Module):
    def __init__(self):
        super(Model, self).__init__()
        self.filepath = None

    def forward(self):
        pass

    def save(self):
        save_dir = os.path.dirname(self.filepath)
        if not os.path.exists(save_dir):
            os.makedirs(save_dir, exist_ok=True)
        torch.save(self.state_dict(), self.filepath)
        print(f'Model {self.__repr__()} saved')

    def save_checkpoint(self, epoch_num):
        torch.save(self.state_dict(), self.filepath + '_' + str(epoch_num))
        print(f'Model checkpoint {self.__repr__()} saved for epoch')

    def load(self, cpu=False):
        if cpu:
            self.load_state_dict(torch.load(self.filepath, map_location=lambda storage, loc: storage))
        else:
            self.load_state_dict(torch.load(self.filepath))
        print(f'Model {self.__repr__()} loaded')

"""Synthetic stand-in for the reference's DatasetManager (music21 + network are not available and are
out of scope): exposes exactly the dataset surface the hot path touches
 and yields uniform-random token tensors of the reference's shapes."""
import torch
from torch.utils.data import TensorDataset, DataLoader

from . import score as _score


class _Metadata:
    def __init__(self, num_values, name):
        self.num_values = num_values
        self.is_global = False
        self.name = name


class TickMetadata(_Metadata):
    def __init__(self, subdivision=6):
        super().__init__(subdivision, "tick")


class BeatMarkerMetadata(_Metadata):
    def __init__(self, subdivision=6):
        super().__init__(6, "beatmarker")


class SyntheticFolkDataset:
    """Attributes: note2index_dicts[0] (len = V), n_bars, subdivision, num_beats_per_bar, num_voices,
    metadatas, NOTES, tick_values / tick_durations, tensor_to_score, data_loaders(batch_size, split) -> 3 DataLoaders of (score int32 (B,1,L), metadata
    int32 (B,1,L,3)) -- DatasetManager/music_dataset.py:177-221, folk_dataset.py:751-861."""

    def __init__(self, num_notes=64, n_bars=16, num_sequences=2048, seed=0, name="folk_4by4nbars_train"):
        self.name = name
        names = _score.default_vocabulary(num_notes)           # rest, slur, START, END, then pitches from G3 upwards
        self.note2index_dicts = [{n: i for i, n in enumerate(names)}]
        self.index2note_dicts = [{i: n for i, n in enumerate(names)}]
        self.tick_values = list(_score.TICK_VALUES)
        self.tick_durations = _score.tick_durations(self.tick_values)
        self.n_bars = n_bars
        self.subdivision = 6
        self.num_beats_per_bar = 4
        self.num_voices = 1
        self.NOTES = 0
        self.metadatas = [BeatMarkerMetadata(6), TickMetadata(6)]
        self.num_sequences = num_sequences
        self.seed = seed
        self.seq_len = n_bars * 24
        self.dataset_filenames = []
        self._tensor_dataset = None

    def __repr__(self):
        return f'SyntheticFolkDataset({len(self.note2index_dicts[0])},{self.n_bars})'

    def empty_score_tensor(self, score_length):
        return torch.zeros(self.num_voices, score_length).long()

    def tensor_to_score(self, tensor_score):
        """Token tensor (any shape, flattened in time order) -> score.Score (folk_dataset.py:472-502 without music21):
        `dataset.tensor_to_score(samples.cpu()).write("midi", fp=path)` is the reference testers' export call."""
        slur = self.note2index_dicts[self.NOTES][_score.SLUR_SYMBOL]
        toks = torch.as_tensor(tensor_score).detach().cpu().reshape(-1).tolist()
        return _score.tokens_to_score(toks, self.index2note_dicts[self.NOTES], slur, self.tick_durations)

    def tensor_dataset(self):
        if self._tensor_dataset is None:
            g = torch.Generator().manual_seed(self.seed)
            V = len(self.note2index_dicts[0])
            score = torch.randint(0, V, (self.num_sequences, 1, self.seq_len), generator=g, dtype=torch.int32)
            t = torch.arange(self.seq_len)
            md = torch.stack([((t // 6) % 4 == 0).long(), t % 6, torch.zeros_like(t)], 1).to(torch.int32)
            md = md.view(1, 1, self.seq_len, 3).expand(self.num_sequences, 1, self.seq_len, 3).contiguous()
            self._tensor_dataset = TensorDataset(score, md)
        return self._tensor_dataset

    def data_loaders(self, batch_size, split=(0.85, 0.10)):
        ds = self.tensor_dataset()
        n = len(ds)
        a, b = int(n * split[0]), int(n * (split[0] + split[1]))
        tr = torch.utils.data.Subset(ds, range(0, a))
        va = torch.utils.data.Subset(ds, range(a, b))
        te = torch.utils.data.Subset(ds, range(b, n))
        pin = torch.cuda.is_available()
        mk = lambda d, shuffle: DataLoader(d, batch_size=batch_size, shuffle=shuffle and len(d) > 0, pin_memory=pin,
                                       drop_last=True)
        return mk(tr, True), mk(va, False), mk(te, False)


class DatasetManager:
    """reference: DatasetManager/dataset_manager.py:122-190 (get_dataset by name)."""

    def get_dataset(self, name, **kwargs):
        n_bars = kwargs.get("num_bars", 16)
        return SyntheticFolkDataset(n_bars=n_bars, name=name)

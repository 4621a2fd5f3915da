"""Dataset item layouts either side of the train step (reference ptb_v2.py:400-489).

The MIDI parsing that produces the arrays (ptb_v2.get_classic_piano / get_vgmidi: pretty_midi, magenta, ...)
is out of scope; what the accelerated step depends on is the TUPLE LAYOUT the reference's DataLoaders
deliver, so these mirrors take the same constructor arguments (already tokenised arrays), apply the same
train/val/test splits and density definitions, and return the same tuples:

    YamahaDataset[i]  -> (x, r, n, c, r_density, n_density)                 (ptb_v2.py:436)
    VGMIDIDataset[i]  -> (x, r, n, c, a, v, r_density, n_density)           (ptb_v2.py:489)

`synthetic_yamaha` / `synthetic_vgmidi` generate arrays of that shape from a seed (the benchmark's and the
tests' data: there is no network for the real corpora).
"""
from __future__ import annotations

from collections import Counter

import numpy as np
import torch
from torch.utils.data import Dataset


def _split(seq, mode, lo, hi):
    tlen, vlen = int(lo * len(seq)), int(hi * len(seq))
    if mode == "train":
        return seq[:tlen]
    if mode == "val":
        return seq[tlen:vlen]
    if mode == "test":
        return seq[vlen:]
    raise ValueError(f"mode must be train / val / test, got {mode!r}")


class YamahaDataset(Dataset):
    """Yamaha e-competition segments, no labels; 80 / 10 / 10 split (ptb_v2.py:400-436)."""

    def __init__(self, data, rhythm, note, chroma, mode="train"):
        super().__init__()
        self.data, self.rhythm, self.note, self.chroma = (_split(a, mode, 0.8, 0.9) for a in (data, rhythm, note, chroma))
        # rhythm density = share of onset tokens (1); note density = mean simultaneous-note count (:421-422)
        self.r_density = [Counter(list(np.asarray(k).tolist()))[1] / len(k) for k in self.rhythm]
        self.n_density = np.array([float(np.sum(k)) / len(k) for k in self.note])

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return (self.data[idx], self.rhythm[idx], self.note[idx], self.chroma[idx], self.r_density[idx],
                self.n_density[idx])


class VGMIDIDataset(Dataset):
    """VGMIDI segments with arousal / valence labels; 90 / 5 / 5 split, EOS (1) inserted before the last token,
    zero padding to the longest sequence, arousal binarised at 0 (ptb_v2.py:439-489)."""

    def __init__(self, data, rhythm, note, chroma, arousal, valence, mode="train"):
        super().__init__()
        data, rhythm, note, chroma, arousal, valence = (_split(a, mode, 0.9, 0.95)
                                                        for a in (data, rhythm, note, chroma, arousal, valence))
        self.r_density = [Counter(list(np.asarray(k).tolist()))[1] / len(k) for k in rhythm]
        self.n_density = np.array([float(np.sum(k)) / len(k) for k in note])
        pad = torch.nn.utils.rnn.pad_sequence
        self.data = pad([torch.Tensor(np.insert(np.asarray(k), -1, 1)) for k in data], batch_first=True)
        self.rhythm = pad([torch.Tensor(np.asarray(k)) for k in rhythm], batch_first=True)
        self.note = pad([torch.Tensor(np.asarray(k)) for k in note], batch_first=True)
        self.chroma = chroma
        self.arousal = np.array(arousal, dtype=np.float64).copy()
        self.arousal[self.arousal >= 0] = 1
        self.arousal[self.arousal < 0] = 0
        self.valence = valence

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return (self.data[idx], self.rhythm[idx], self.note[idx], self.chroma[idx], self.arousal[idx], self.valence[idx],
                self.r_density[idx], self.n_density[idx])


def synthetic_yamaha(n: int, T: int, seed: int = 0):
    """(data, rhythm, note, chroma) arrays shaped like get_classic_piano()'s output: event tokens in [2, 342) with
    an EOS (1) + zero padding tail, rhythm classes {0,1,2}, note counts [0,16), 24-d chroma."""
    g = np.random.default_rng(seed)
    data = g.integers(2, 342, size=(n, T)).astype(np.int64)
    tail = max(1, T // 8)
    data[:, T - tail] = 1
    data[:, T - tail + 1:] = 0
    rhythm = g.integers(0, 3, size=(n, T)).astype(np.int64)
    note = g.integers(0, 16, size=(n, T)).astype(np.int64)
    chroma = g.random((n, 24)).astype(np.float32)
    return data, rhythm, note, chroma


def synthetic_vgmidi(n: int, T: int, seed: int = 0):
    """(data, rhythm, note, arousal, valence, chroma) like get_vgmidi(): ragged sequences (VGMIDIDataset pads)."""
    g = np.random.default_rng(seed)
    lens = g.integers(max(2, T // 2), T, size=n)
    data = [g.integers(2, 342, size=int(l)).astype(np.int64) for l in lens]
    rhythm = [g.integers(0, 3, size=int(l) + 1).astype(np.int64) for l in lens]       # +1: the inserted EOS position
    note = [g.integers(0, 16, size=int(l) + 1).astype(np.int64) for l in lens]
    arousal = g.uniform(-1, 1, size=n)
    valence = g.uniform(-1, 1, size=n)
    chroma = g.random((n, 24)).astype(np.float32)
    return data, rhythm, note, arousal, valence, chroma

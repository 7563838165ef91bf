"""Dataset classes with the reference's interface (VarAutoEncoder/data.py).

``Loader`` walks ``path/<class_dir>/*.mid`` and tokenises every first surviving track on the GPU (one K1
launch for the whole directory tree); ``MelodyDataset`` chunks the ids into rows with the rules of
``_get_token_arrays`` (data.py:133-173, quirks included) and iterates shuffled batches the way
``mx.io.NDArrayIter(shuffle=True)`` with ``last_batch_handle='pad'`` does.  Batches carry float32 arrays as
the reference's do (``data = [tokens, seq_lens, classes]``, ``label = [labels]``)."""
import glob
import os
from typing import Dict, List

import numpy as np
import torch

from ..MIDIUtil.defaults import *  # noqa: F401,F403
from ..MIDIUtil.defaults import EOS_ID, NUM_EVENTS, PAD_ID, SOS_ID
from ..MIDIUtil.Melody import Melody
from ..MIDIUtil.midi_io import EventBasedMIDIReader


class DataBatch:
    def __init__(self, data, label, pad=0):
        self.data = data
        self.label = label
        self.pad = pad


class Loader:
    def __init__(self, path: str, max_sequence_length: int, slices_per_quarter_note: int):
        self.path = path
        self.max_sequence_length = max_sequence_length
        self.slices_per_quarter_note = slices_per_quarter_note
        self.midi_reader = EventBasedMIDIReader()
        self.melodies = self.read_melodies()

    def read_melodies(self):
        print("Reading from {}".format(self.path))
        melodies = {}
        directories = next(os.walk(self.path))[1]
        files = {d: glob.glob(self.path + '/' + d + "/*.mid") for d in sorted(directories)}
        parsed = self.midi_reader.read_files([f for d in files for f in files[d]])
        for directory in sorted(directories):
            melodies[directory] = [parsed[f][0] for f in files[directory]]       # data.py:35 keeps track [0]
            print("Read {} files from {}".format(len(files[directory]), directory))
        return melodies


class Dataset:
    def __init__(self, batch_size: int):
        self.batch_size = batch_size

    def num_classes(self):
        raise NotImplementedError

    def num_tokens(self):
        raise NotImplementedError

    def __iter__(self):
        raise NotImplementedError


class _ArrayIter:
    """mx.io.NDArrayIter semantics used by the reference: optional shuffle once per reset, last batch padded by
    wrapping around to the first samples."""

    def __init__(self, data, label, batch_size, shuffle, seed=0):
        self.data, self.label, self.batch_size, self.shuffle = data, label, batch_size, shuffle
        self.n = data[0].shape[0]
        self.rng = np.random.RandomState(seed)
        self.order = np.arange(self.n)

    def reset(self):
        if self.shuffle:
            self.rng.shuffle(self.order)

    def __iter__(self):
        for start in range(0, self.n, self.batch_size):
            idx = self.order[start:start + self.batch_size]
            pad = self.batch_size - len(idx)
            if pad > 0:
                idx = np.concatenate([idx, self.order[:pad]])
            yield DataBatch([torch.from_numpy(a[idx]) for a in self.data], [torch.from_numpy(a[idx]) for a in self.label], pad)


class ToyData(Dataset):
    """data.py:57-81."""

    def __init__(self, batch_size: int = 3):
        super().__init__(batch_size)
        tokens = np.array([[1, 5, 6, 7, 0], [1, 6, 7, 8, 0], [1, 7, 8, 9, 0]], dtype=np.float32)
        seq_lens = np.array([4, 4, 4], dtype=np.float32)
        classes = np.array([0, 1, 2], dtype=np.float32)
        labels = np.array([[5, 6, 7, 2, 0], [6, 7, 8, 2, 0], [7, 8, 9, 2, 0]], dtype=np.float32)
        self.iter = _ArrayIter([tokens, seq_lens, classes], [labels], self.batch_size, shuffle=False)

    def num_classes(self):
        return 3

    def num_tokens(self):
        return 10

    def __iter__(self):
        self.iter.reset()
        for batch in self.iter:
            yield batch


class MelodyDataset(Dataset):
    def __init__(self, batch_size: int, maximum_sequence_length: int, melodies: Dict[str, List[Melody]], seed: int = 0):
        super().__init__(batch_size)
        self.max_seq_len = maximum_sequence_length
        self.mask_offset = 1
        self._seed = seed
        self._initialize(melodies)
        self._log_dataset()
        del self.melodies

    def _initialize(self, melodies):
        self.melodies = dict(sorted(melodies.items(), key=lambda x: x[0]))
        self.n_classes = len(self.melodies)
        self.n_melodies = sum(len(m) for m in self.melodies.values())
        self.seen_max_sequence_length = max(max(len(x.notes) for x in ms) for ms in self.melodies.values())
        self._get_token_arrays()
        self.iter = _ArrayIter([self.tokens, self.classes], [self.labels], self.batch_size, shuffle=True, seed=self._seed)

    def _log_dataset(self):
        print("")
        print("Dataset information: ")
        print("Number of classes: {}".format(self.num_classes()))
        print("Number of tokens: {}".format(self.num_tokens()))
        print("Tokens dataset shape {}".format(self.tokens.shape))
        print("Classes dataset shape {}".format(self.classes.shape))
        for c, m in self.melodies.items():
            print("Class {} has {} melodies of maximum length {}".format(c, len(m), self.seen_max_sequence_length))
        print("")

    def num_classes(self):
        return self.n_classes

    def num_tokens(self):
        return NUM_EVENTS

    def _get_token_arrays(self):
        """data.py:133-173: rows of L ids; the remainder row of every melody is flushed even when empty (:149-150);
        after each class the last row is appended once more when it is non-empty (:152-155); labels are the data
        shifted by one with EOS written by ``labels[:, seq_lens] = EOS`` (:166-168, NumPy advanced indexing:
        every row gets EOS in every column that is some row's length)."""
        L = self.max_seq_len
        rows, classes = [], []
        last = None
        for class_idx, melodies_for_class in enumerate(self.melodies.values()):
            for melody in melodies_for_class:
                ids = np.fromiter((e.id for e in melody), dtype=np.int64, count=len(melody))
                n_full = len(ids) // L
                for r in range(n_full):
                    rows.append(ids[r * L:(r + 1) * L])
                    classes.append(class_idx)
                last = np.full((L,), PAD_ID, dtype=np.int64)
                rem = ids[n_full * L:]
                last[:len(rem)] = rem
                rows.append(last)
                classes.append(class_idx)
            if last[0] != PAD_ID:
                rows.append(last)
                classes.append(class_idx)
        num_samples = len(rows)
        assert num_samples > 0, "Empty sequences were found"
        data = np.stack(rows).astype(np.float32)
        self.tokens = np.concatenate([np.full((num_samples, 1), SOS_ID, np.float32), data], axis=1)
        seq_lens = self._count_sequence_length(data).astype(np.int64)
        self.labels = np.concatenate([data, np.full((num_samples, 1), PAD_ID, np.float32)], axis=1)
        self.labels[:, seq_lens] = EOS_ID
        self.classes = np.asarray(classes, dtype=np.float32)
        print("Tokens.shape {}".format(self.tokens.shape))
        print("Labels.shape {}".format(self.labels.shape))
        print("classes.shape {}".format(self.classes.shape))

    def _count_sequence_length(self, tokens):
        t = tokens.numpy() if torch.is_tensor(tokens) else tokens
        return (t != PAD_ID).sum(axis=1).astype(np.float32)

    def __iter__(self):
        self.iter.reset()
        for batch in self.iter:
            self._preprocess_batch(batch)
            yield batch

    def _preprocess_batch(self, batch):
        """data.py:187-198: seq_lens = #non-PAD incl. SOS inserted at data[1]; trim to the batch maximum.
        The lengths are counted on host copies, so no device synchronisation is involved."""
        tokens = batch.data[0]
        seq_lens = torch.from_numpy(self._count_sequence_length(tokens))
        batch.data.insert(1, seq_lens)
        max_seq_len = int(seq_lens.max())
        batch.data[0] = batch.data[0][:, :max_seq_len]
        batch.label[0] = batch.label[0][:, :max_seq_len]


class DeviceMelodyDataset(Dataset):
    """MelodyDataset with the rows built and kept on the GPU (A2 on the device): the tracks' note events are tokenised by
    K1, ``msx_rows_plan`` / ``msx_rows_build`` chunk them with the rules of ``_get_token_arrays`` (data.py:133-173, quirks
    included) and every batch is one ``msx_rows_gather_batch`` launch — no row ever exists on the host.  Batches carry int32
    CUDA tensors (``data = [tokens, seq_lens, classes]``, ``label = [labels]``), shuffled and wrap-padded like
    ``mx.io.NDArrayIter(shuffle=True)`` and trimmed to the batch's longest row (data.py:187-198; the lengths are known on
    the host from one copy made when the dataset is built, so iterating never synchronises)."""

    def __init__(self, batch_size: int, maximum_sequence_length: int, melodies: Dict[str, List[Melody]], seed: int = 0,
                 device="cuda"):
        super().__init__(batch_size)
        from .. import featurise
        self.max_seq_len = maximum_sequence_length
        melodies = dict(sorted(melodies.items(), key=lambda x: x[0]))
        self.n_classes = len(melodies)
        flat = [(c, m) for c, ms in enumerate(melodies.values()) for m in ms]
        assert flat, "Empty sequences were found"
        class_start = np.zeros(self.n_classes + 1, np.int32)
        for c, _ in flat:
            class_start[c + 1:] += 1
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        if all(getattr(m, "soa", None) is not None for _, m in flat):
            tokens, n_tokens = featurise.tokenize_tracks_device([m.soa for _, m in flat], device)      # A1 on the device
        else:                                                   # melodies given as ids (no note events): upload the ids
            width = max(1, max(len(m) for _, m in flat))
            host = np.zeros((len(flat), width + 1), np.int32)
            host[:, 0] = SOS_ID
            for i, (_, m) in enumerate(flat):
                host[i, 1:1 + len(m)] = [e.id for e in m]
            tokens, n_tokens = t(host), t(np.asarray([len(m) for _, m in flat], np.int32))
        self.rows = featurise.build_rows(tokens, n_tokens, t(np.asarray([c for c, _ in flat], np.int32)), t(class_start),
                                         self.max_seq_len)
        self.tokens, self.labels, self.classes, self.seq_lens = self.rows
        self._host_lens = self.seq_lens.cpu().numpy()
        self.n = int(self.tokens.shape[0])
        self.rng = np.random.RandomState(seed)
        self.order = np.arange(self.n)
        self.device = device
        print("Tokens.shape {}".format(tuple(self.tokens.shape)))
        print("Labels.shape {}".format(tuple(self.labels.shape)))
        print("classes.shape {}".format(tuple(self.classes.shape)))

    def num_classes(self):
        return self.n_classes

    def num_tokens(self):
        return NUM_EVENTS

    def __iter__(self):
        from .. import featurise
        self.rng.shuffle(self.order)
        B = self.batch_size
        for start in range(0, self.n, B):
            idx = self.order[start:start + B]
            pad = B - len(idx)
            if pad > 0:
                idx = np.concatenate([idx, self.order[:pad]])
            t_out = int(self._host_lens[idx].max())
            index = torch.from_numpy(idx.astype(np.int32)).to(self.device, non_blocking=True)
            tok, lab, cls, lens = featurise.gather_batch(self.rows, index, t_out)
            yield DataBatch([tok, lens, cls], [lab], pad)


class RollDataset(Dataset):
    """Piano-roll windows for ``--featurisation roll``: every track is rasterised by K1 into consecutive windows of
    ``n_slices`` slices (``featurise.rasterize_windows``, up to 1024 slices per track), each valid window is one row
    ``uint8 [n_slices, 128]`` with its track's class.  Rows live on the GPU; batches are shuffled / wrap-padded like
    MelodyDataset's and carry ``data = [roll, classes]``, ``label = [roll]``."""

    def __init__(self, batch_size: int, n_slices: int, melodies: Dict[str, List[Melody]], slices_per_quarter_note=4,
                 seed: int = 0, device="cuda"):
        super().__init__(batch_size)
        from .. import featurise
        melodies = dict(sorted(melodies.items(), key=lambda x: x[0]))
        self.n_classes = len(melodies)
        self.n_slices = int(n_slices)
        max_windows = max(1, 1024 // self.n_slices)
        rolls, classes = [], []
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        by_res = {}
        for c, ms in enumerate(melodies.values()):
            for m in ms:
                assert getattr(m, "soa", None) is not None, "RollDataset needs melodies read from MIDI files (note events)"
                by_res.setdefault(int(m.resolution), []).append((c, m.soa))
        for res, items in sorted(by_res.items()):
            lens = [len(s[0]) for _, s in items]
            offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
            cat = lambda i, dt: np.concatenate([np.asarray(s[i], dtype=dt) for _, s in items])
            roll, n_win = featurise.rasterize_windows(t(cat(0, np.int32)), t(cat(1, np.uint8)), t(cat(2, np.uint8)), t(offs),
                                                      resolution=res, slices_per_quarter=int(slices_per_quarter_note),
                                                      n_slices=self.n_slices, max_windows=max_windows)
            n_win_h = n_win.cpu().numpy()
            for i, (c, _) in enumerate(items):
                rolls.append(roll[i, :int(n_win_h[i])])
                classes += [c] * int(n_win_h[i])
        self.rolls = torch.cat(rolls, dim=0).contiguous()                       # [R, S, 128] uint8 on the device
        self.classes = t(np.asarray(classes, np.int32))
        self.n = int(self.rolls.shape[0])
        assert self.n > 0, "Empty sequences were found"
        self.rng = np.random.RandomState(seed)
        self.order = np.arange(self.n)
        self.device = device
        print("Rolls.shape {}".format(tuple(self.rolls.shape)))

    def num_classes(self):
        return self.n_classes

    def num_tokens(self):
        return NUM_EVENTS

    def __iter__(self):
        self.rng.shuffle(self.order)
        B = self.batch_size
        for start in range(0, self.n, B):
            idx = self.order[start:start + B]
            pad = B - len(idx)
            if pad > 0:
                idx = np.concatenate([idx, self.order[:pad]])
            index = torch.from_numpy(idx.astype(np.int64)).to(self.device, non_blocking=True)
            roll = self.rolls.index_select(0, index)
            yield DataBatch([roll, self.classes.index_select(0, index)], [roll], pad)


def load_dataset(loader_train: Loader, batch_size: int, split_percentage: float = None, loader_val: Loader = None,
                 device_rows: bool = False):
    """data.py:201-223.  device_rows: build and iterate the rows on the GPU (DeviceMelodyDataset)."""
    make = DeviceMelodyDataset if device_rows else MelodyDataset
    if loader_val is not None:
        return (make(batch_size, loader_train.max_sequence_length, loader_train.melodies),
                make(batch_size, loader_val.max_sequence_length, loader_val.melodies))
    if split_percentage <= 0.:
        return make(batch_size, loader_train.max_sequence_length, loader_train.melodies), None
    assert 0.0 < split_percentage < 1.0
    train_split, valid_split = {}, {}
    for c, m in loader_train.melodies.items():
        n_validation_melodies = int(split_percentage * len(m))
        valid_split[c] = m[:n_validation_melodies]
        train_split[c] = m[n_validation_melodies:]
    return (make(batch_size, loader_train.max_sequence_length, train_split),
            make(batch_size, loader_train.max_sequence_length, valid_split))

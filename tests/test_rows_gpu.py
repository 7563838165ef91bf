"""A2 on the device: msx_rows_plan / msx_rows_build / msx_rows_gather_batch against the rows the reference's own
MelodyDataset produced for the 37 fixture tracks (tests/golden/rows_fixtures.npz) and against the host restatement
(VarAutoEncoder/data.py MelodyDataset, itself pinned to the same goldens) on synthetic edge cases.  Bit-exact."""
import os

import numpy as np
import pytest
import torch

from musicstyletransfer_b200.MIDIUtil import Melody as M

pytestmark = pytest.mark.gpu


def _melodies_from_golden(golden_dir, with_soa):
    g = np.load(os.path.join(golden_dir, "tokens_fixtures.npz"))
    melodies = {}
    for name in g["names"]:
        mel = M.Melody()
        mel.notes = [M.create_event_from_id(int(i)) for i in g["ids:" + name]]
        if with_soa:
            mel.soa = (g["dtick:" + name], g["pitch:" + name], g["vel:" + name])
        melodies.setdefault(name.split("/")[0], []).append(mel)
    return melodies


@pytest.mark.parametrize("with_soa", [False, True])
@pytest.mark.parametrize("L", [64, 16])
def test_device_rows_match_reference_dataset(golden_dir, L, with_soa):
    """with_soa: note events -> K1 tokens -> rows, everything on the device; else the golden ids are uploaded."""
    from musicstyletransfer_b200.VarAutoEncoder.data import DeviceMelodyDataset
    r = np.load(os.path.join(golden_dir, "rows_fixtures.npz"))
    ds = DeviceMelodyDataset(32, L, _melodies_from_golden(golden_dir, with_soa))
    assert np.array_equal(ds.tokens.cpu().numpy(), r["tokens_L%d" % L].astype(np.int32))
    assert np.array_equal(ds.labels.cpu().numpy(), r["labels_L%d" % L].astype(np.int32))
    assert np.array_equal(ds.classes.cpu().numpy(), r["classes_L%d" % L].astype(np.int32))
    assert np.array_equal(ds.seq_lens.cpu().numpy(), (r["tokens_L%d" % L] != 0).sum(1))
    # iteration: wrap-padded shuffled batches, trimmed to the longest row of the batch (data.py:187-198)
    n, seen = 0, []
    tok_all, lab_all = ds.tokens.cpu().numpy(), ds.labels.cpu().numpy()
    order = None
    for batch in ds:
        order = ds.order
        tokens, seq_lens, classes = [x.cpu().numpy() for x in batch.data]
        labels = batch.label[0].cpu().numpy()
        idx = np.concatenate([order, order])[n * 32:(n + 1) * 32] if (n + 1) * 32 <= len(order) else \
            np.concatenate([order[n * 32:], order[:(n + 1) * 32 - len(order)]])
        T = int(seq_lens.max())
        assert tokens.shape == (32, T) and labels.shape == (32, T)
        assert np.array_equal(tokens, tok_all[idx][:, :T]) and np.array_equal(labels, lab_all[idx][:, :T])
        assert np.array_equal(seq_lens, (tok_all[idx] != 0).sum(1))
        assert np.array_equal(classes, ds.classes.cpu().numpy()[idx])
        seen.append(idx)
        n += 1
    assert n == int(np.ceil(tok_all.shape[0] / 32))
    if L == 64:
        assert n == 28                       # 880 rows -> 28 batches of 32


@pytest.mark.parametrize("L", [8, 5])
def test_device_rows_edge_cases_match_host_dataset(L):
    """Lengths that are exact multiples of L (empty remainder row, no duplicate), empty melodies, one-melody classes,
    a class whose last melody ends on a row boundary after one that does not."""
    from musicstyletransfer_b200.VarAutoEncoder.data import DeviceMelodyDataset, MelodyDataset
    rng = np.random.RandomState(L)
    lens = {"a": [L, 2 * L, 3], "b": [0, 1, L - 1, L + 1, 4 * L], "c": [7], "d": [2 * L], "e": [L + 2, 0]}
    melodies = {}
    for c, ls in lens.items():
        melodies[c] = []
        for n in ls:
            mel = M.Melody()
            mel.notes = [M.create_event_from_id(int(i)) for i in rng.randint(3, 293, size=n)]
            melodies[c].append(mel)
    host = MelodyDataset(4, L, melodies)
    dev = DeviceMelodyDataset(4, L, melodies)
    assert np.array_equal(dev.tokens.cpu().numpy(), host.tokens.astype(np.int32))
    assert np.array_equal(dev.labels.cpu().numpy(), host.labels.astype(np.int32))
    assert np.array_equal(dev.classes.cpu().numpy(), host.classes.astype(np.int32))


def test_device_rows_many_tracks():
    """More tracks than one scan chunk of the plan kernel (1024)."""
    from musicstyletransfer_b200.VarAutoEncoder.data import DeviceMelodyDataset, MelodyDataset
    rng = np.random.RandomState(1)
    melodies = {"x": [], "y": []}
    for c in melodies:
        for _ in range(1500):
            mel = M.Melody()
            mel.notes = [M.create_event_from_id(int(i)) for i in rng.randint(3, 293, size=int(rng.randint(0, 40)))]
            melodies[c].append(mel)
    host = MelodyDataset(32, 16, melodies)
    dev = DeviceMelodyDataset(32, 16, melodies)
    assert np.array_equal(dev.tokens.cpu().numpy(), host.tokens.astype(np.int32))
    assert np.array_equal(dev.labels.cpu().numpy(), host.labels.astype(np.int32))
    assert np.array_equal(dev.classes.cpu().numpy(), host.classes.astype(np.int32))

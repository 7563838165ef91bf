"""K1 fed from .mid bytes: C++ SMF parser -> note-event SoA -> msx_rasterize -> the token ids the reference's
EventBasedMIDIReader produced for the same files (tests/golden/tokens_fixtures.npz), all 37 fixtures; and the whole chain
bytes -> rows of the reference's MelodyDataset (rows_fixtures.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fixture_paths(golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "midi_fixtures.npz"))
    root = tmp_path / "guitar_bass"
    for name in g["names"]:
        if not name.startswith("guitar_bass/"):
            continue
        _, cls, fname = name.split("/")
        (root / cls).mkdir(parents=True, exist_ok=True)
        (root / cls / fname).write_bytes(g["bytes:" + name].tobytes())
    return str(root)


def test_reader_from_mid_bytes_matches_reference_ids(golden_dir, tmp_path):
    from musicstyletransfer_b200.MIDIUtil.midi_io import EventBasedMIDIReader
    root = _fixture_paths(golden_dir, tmp_path)
    t = np.load(os.path.join(golden_dir, "tokens_fixtures.npz"))
    reader = EventBasedMIDIReader()
    files = [os.path.join(root, n) for n in t["names"]]
    parsed = reader.read_files(files)
    for name, f in zip(t["names"], files):
        mel = parsed[f][0]                                  # data.py:35 keeps the first surviving track
        assert np.array_equal(np.asarray([e.id for e in mel.notes]), t["ids:" + name]), name
        assert mel.resolution == int(t["res:" + name])


def test_loader_to_device_rows_matches_reference_dataset(golden_dir, tmp_path):
    """scripts/train-vae.sh's data path end to end on the device: Loader (.mid bytes -> C++ parser -> K1) ->
    DeviceMelodyDataset (A2 kernels) == the reference's MelodyDataset arrays."""
    from musicstyletransfer_b200.VarAutoEncoder.data import Loader, load_dataset
    root = _fixture_paths(golden_dir, tmp_path)
    r = np.load(os.path.join(golden_dir, "rows_fixtures.npz"))
    # glob order inside a class directory is file-system order; the golden rows were made from sorted names
    import glob as _glob
    orig = _glob.glob
    _glob.glob = lambda p: sorted(orig(p))
    try:
        loader = Loader(root, 64, 4)
    finally:
        _glob.glob = orig
    ds, val = load_dataset(loader, 32, 0.0, device_rows=True)
    assert val is None
    assert np.array_equal(ds.tokens.cpu().numpy(), r["tokens_L64"].astype(np.int32))
    assert np.array_equal(ds.labels.cpu().numpy(), r["labels_L64"].astype(np.int32))
    assert np.array_equal(ds.classes.cpu().numpy(), r["classes_L64"].astype(np.int32))

"""K1 parity: CUDA rasteriser (through the C ABI) vs the oracle, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import featurise as of

pytestmark = pytest.mark.gpu


def _run(dtick, pitch, vel, offs, **kw):
    from musicstyletransfer_b200 import featurise
    dev = "cuda:0"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tok, roll, cnt = featurise.rasterize(t(dtick), t(pitch), t(vel), t(offs), **kw)
    torch.cuda.synchronize()
    return tok.cpu().numpy(), roll.cpu().numpy(), cnt.cpu().numpy()


@pytest.mark.parametrize("velocity_roll", [False, True])
def test_config2_subset_bit_exact(velocity_roll):
    dtick, pitch, vel, offs = of.synth_note_events(n_seq=2048, ev_per_seq=32, seed=0)
    tok, roll, cnt = _run(dtick, pitch, vel, offs, velocity_roll=velocity_roll)
    otok, oroll, ocnt = of.rasterize_batch(dtick, pitch, vel, offs, velocity_roll=velocity_roll)
    assert np.array_equal(cnt, ocnt)
    assert np.array_equal(tok, otok)
    assert np.array_equal(roll, oroll)


def test_ragged_empty_and_long_sequences():
    rng = np.random.RandomState(5)
    lens = [0, 1, 2, 31, 32, 33, 64, 100, 257, 0, 7]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    E = int(offs[-1])
    dtick = (rng.randint(0, 4, size=E) * 15).astype(np.int32)       # off-grid ticks, many same-slice events
    dtick[rng.rand(E) < 0.03] = 2500
    pitch = rng.randint(58, 63, size=E).astype(np.uint8)            # few pitches -> re-triggers, unmatched offs
    vel = np.where(rng.rand(E) < 0.5, rng.randint(1, 128, size=E), 0).astype(np.uint8)
    for vr in (False, True):
        for S, L in ((64, 64), (16, 8), (128, 200)):
            tok, roll, cnt = _run(dtick, pitch, vel, offs, n_slices=S, max_seq_len=L, velocity_roll=vr)
            otok, oroll, ocnt = of.rasterize_batch(dtick, pitch, vel, offs, n_slices=S, max_seq_len=L, velocity_roll=vr)
            assert np.array_equal(cnt, ocnt)
            assert np.array_equal(tok, otok)
            assert np.array_equal(roll, oroll), (vr, S, L)


def test_other_resolutions():
    dtick, pitch, vel, offs = of.synth_note_events(n_seq=64, ev_per_seq=48, seed=3)
    for res, spq in ((96, 4), (220, 4), (480, 8), (120, 3)):
        tok, roll, cnt = _run(dtick, pitch, vel, offs, resolution=res, slices_per_quarter=spq)
        otok, oroll, ocnt = of.rasterize_batch(dtick, pitch, vel, offs, resolution=res, slices_per_quarter=spq)
        assert np.array_equal(tok, otok) and np.array_equal(roll, oroll) and np.array_equal(cnt, ocnt)


def test_fixture_tracks_first_window(golden_dir):
    """The 37 reference MIDI fixtures: first token row == reference reader's first 64 tokens; roll == oracle."""
    import os
    g = np.load(os.path.join(golden_dir, "tokens_fixtures.npz"))
    names = list(g["names"])
    dt = [g["dtick:" + n] for n in names]
    offs = np.concatenate([[0], np.cumsum([len(x) for x in dt])]).astype(np.int32)
    dtick = np.concatenate(dt).astype(np.int32)
    pitch = np.concatenate([g["pitch:" + n] for n in names]).astype(np.uint8)
    vel = np.concatenate([g["vel:" + n] for n in names]).astype(np.uint8)
    tok, roll, cnt = _run(dtick, pitch, vel, offs)
    otok, oroll, ocnt = of.rasterize_batch(dtick, pitch, vel, offs)
    for i, n in enumerate(names):
        ref_ids = g["ids:" + n]
        assert cnt[i] == len(ref_ids)
        k = min(64, len(ref_ids))
        assert np.array_equal(tok[i, 1:1 + k], ref_ids[:k]) and tok[i, 0] == 1
    assert np.array_equal(roll, oroll) and np.array_equal(tok, otok)


def test_full_size_properties():
    """BASELINE config 2 at full size (1 M events): size-independent properties + sampled oracle rows."""
    dtick, pitch, vel, offs = of.synth_note_events()
    tok, roll, cnt = _run(dtick, pitch, vel, offs)
    assert tok.shape == (32768, 65) and roll.shape == (32768, 64, 128)
    assert (tok[:, 0] == 1).all()
    assert set(np.unique(roll)) <= {0, 1}
    # token count = events + shift tokens
    d = dtick.reshape(32768, 32).astype(np.int64)
    assert np.array_equal(cnt, (32 + ((d + 999) // 1000).sum(axis=1)).astype(np.int32))
    # PAD exactly beyond the count
    pos = np.arange(64)[None, :]
    assert np.array_equal(tok[:, 1:] != 0, pos < np.minimum(cnt, 64)[:, None])
    idx = np.random.RandomState(1).choice(32768, 256, replace=False)
    for i in idx:
        a, b = offs[i], offs[i + 1]
        ids, r = of.rasterize_sequence(dtick[a:b], pitch[a:b], vel[a:b], 120, 4, 64, 1)
        assert np.array_equal(roll[i], r[0])
    # the whole 1 M-event output, bit for bit, against the C restatement of the oracle
    from oracle import raster_c
    if raster_c.available():
        ctok, croll, ccnt = raster_c.rasterize_batch(dtick, pitch, vel, offs, threads=8)
        assert np.array_equal(tok, ctok) and np.array_equal(roll, croll) and np.array_equal(cnt, ccnt)
        vtok, vroll, vcnt = _run(dtick, pitch, vel, offs, velocity_roll=True)
        ctok, croll, ccnt = raster_c.rasterize_batch(dtick, pitch, vel, offs, velocity_roll=True, threads=8)
        assert np.array_equal(vroll, croll) and np.array_equal(vtok, ctok)


@pytest.mark.parametrize("velocity_roll", [False, True])
def test_whole_tracks_as_windows(velocity_roll):
    """featurise.rasterize_windows: ragged tracks of up to a few hundred events as consecutive 64-slice windows, every
    valid window bit-exact against oracle rasterize_sequence(max_windows); empty and one-event tracks included."""
    from musicstyletransfer_b200 import featurise
    rng = np.random.RandomState(11)
    lens = [0, 1, 5, 40, 130, 260, 33, 700, 64, 2]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    E = int(offs[-1])
    dtick = (30 * rng.randint(0, 9, size=E)).astype(np.int32)
    dtick[rng.rand(E) < 0.02] = rng.randint(1000, 5000, size=int((rng.rand(E) < 0.02).sum()) or 1)[0]
    pitch = rng.randint(40, 72, size=E).astype(np.uint8)
    vel = np.where(rng.rand(E) < 0.5, rng.randint(1, 128, size=E), 0).astype(np.uint8)
    S, W = 64, 8
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    roll, nwin = featurise.rasterize_windows(t(dtick), t(pitch), t(vel), t(offs), n_slices=S, max_windows=W,
                                             velocity_roll=velocity_roll)
    torch.cuda.synchronize()
    roll, nwin = roll.cpu().numpy(), nwin.cpu().numpy()
    for i in range(len(lens)):
        a, b = offs[i], offs[i + 1]
        _, want = of.rasterize_sequence(dtick[a:b], pitch[a:b], vel[a:b], 120, 4, S, W, velocity_roll=velocity_roll)
        want = want.reshape(-1, S, 128)
        assert nwin[i] == want.shape[0], (i, nwin[i], want.shape)
        assert np.array_equal(roll[i, :nwin[i]], want), i


def test_malformed_data_bytes_are_masked():
    """Data bytes >= 0x80 (malformed MIDI: pitch 200, velocity 255) must not index past the 128-entry pitch table or
    produce ids >= the vocabulary; kernel and oracle both treat them as 7-bit."""
    rng = np.random.RandomState(11)
    n_seq, ev = 300, 40
    offs = (np.arange(n_seq + 1) * ev).astype(np.int32)
    E = n_seq * ev
    dtick = (rng.randint(0, 6, size=E) * 30).astype(np.int32)
    pitch = rng.randint(0, 256, size=E).astype(np.uint8)
    pitch[::7] = 200
    vel = np.where(rng.rand(E) < 0.5, rng.randint(1, 256, size=E), 0).astype(np.uint8)
    for vr in (False, True):
        tok, roll, cnt = _run(dtick, pitch, vel, offs, velocity_roll=vr)
        otok, oroll, ocnt = of.rasterize_batch(dtick, pitch, vel, offs, velocity_roll=vr)
        assert tok.max() < 293 and tok.min() >= 0
        assert np.array_equal(tok, otok) and np.array_equal(roll, oroll) and np.array_equal(cnt, ocnt)
    # a clean launch afterwards proves no sticky fault / corrupted neighbour tile
    d2, p2, v2, o2 = of.synth_note_events(n_seq=64, ev_per_seq=32, seed=1)
    tok, roll, cnt = _run(d2, p2, v2, o2)
    otok, oroll, ocnt = of.rasterize_batch(d2, p2, v2, o2)
    assert np.array_equal(tok, otok) and np.array_equal(roll, oroll)

"""Pins the CPU oracle to golden vectors produced by executing the reference's own source
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import torch

from oracle import featurise as of
from oracle import model as om


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_tokens_match_reference_reader(golden_dir):
    """A1: oracle tokenisation == EventBasedMIDIReader._parse_track (midi_io.py:70-93), 37 files."""
    g = _load(golden_dir, "tokens_fixtures.npz")
    total = 0
    for name in g["names"]:
        ids = of.tokenize_note_events(g["dtick:" + name], g["pitch:" + name], g["vel:" + name])
        assert np.array_equal(np.asarray(ids, dtype=np.int32), g["ids:" + name]), name
        total += len(ids)
    assert total == 55036
    assert len(g["names"]) == 37


def test_rows_match_reference_dataset(golden_dir):
    """A2: oracle chunking == MelodyDataset._get_token_arrays (data.py:133-173), L=64 and L=16."""
    g = _load(golden_dir, "tokens_fixtures.npz")
    r = _load(golden_dir, "rows_fixtures.npz")
    by_class = {}
    for name in g["names"]:
        by_class.setdefault(name.split("/")[0], []).append(list(g["ids:" + name]))
    melodies = [by_class[c] for c in r["class_names"]]
    for L in (64, 16):
        tok, lab, cls = of.chunk_rows(melodies, L)
        assert np.array_equal(tok, r["tokens_L%d" % L])
        assert np.array_equal(lab, r["labels_L%d" % L])
        assert np.array_equal(cls, r["classes_L%d" % L])
    assert r["tokens_L64"].shape == (880, 65)


def test_losses_match_reference(golden_dir):
    """A8/A9/A10: oracle formulas == reference loss.py classes."""
    g = _load(golden_dir, "loss_golden.npz")
    t = lambda k: torch.from_numpy(g[k])
    np.testing.assert_allclose(om.kl_loss(t("kl_means"), t("kl_stds")).numpy(), g["kl"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(om.softmax_ce(t("ce_probs"), t("ce_labels")).numpy(), g["ce"], rtol=1e-5, atol=1e-6)
    pred, label = t("bce_pred"), t("bce_label")
    np.testing.assert_allclose(om.bce_loss(pred, label).numpy(), g["bce_default"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(om.bce_loss(pred, label, label_smoothing=0.1).numpy(), g["bce_smooth"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(om.bce_loss(pred, label, negative_label_downweighting=False).numpy(),
                               g["bce_noweight"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(om.bce_loss(torch.sigmoid(pred), label, from_sigmoid=True).numpy(),
                               g["bce_fromsig"], rtol=2e-5, atol=1e-6)


def _params(g, prefix, rename=None):
    p = {}
    for k in g.files:
        if k.startswith("param:" + prefix):
            name = k[len("param:"):]
            if rename:
                name = rename(name)
            p[name] = torch.from_numpy(g[k])
    return p


def test_toy_model_forward_matches_reference(golden_dir):
    """A3-A5, A7-A9: oracle forward == reference Model on ToyData (data.py:62-70, main.py:14-38)."""
    g = _load(golden_dir, "model_toy.npz")
    cfg = om.toy_cfg()
    p = _params(g, "")
    assert set(p) == set(om.param_shapes(cfg))
    t = lambda k: torch.from_numpy(g[k])
    loss, ce, kl, probs, means, stds = om.step_losses(cfg, p, t("tokens"), t("seq_lens"), t("classes"),
                                                      t("labels"), t("eps"))
    np.testing.assert_allclose(means.numpy(), g["means"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(stds.numpy(), g["stds"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(probs.numpy(), g["probs"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ce.numpy(), g["ce"], rtol=1e-4)
    np.testing.assert_allclose(kl.numpy(), g["kl"], rtol=1e-4)


def test_small_model_forward_matches_reference(golden_dir):
    """Ragged batch with PAD rows: Encoder, HEAD Transformer Decoder and LSTMDecoder forward."""
    g = _load(golden_dir, "model_small.npz")
    t = lambda k: torch.from_numpy(g[k])
    cfg = om.Cfg(vocab=293, num_classes=2, enc_size=64, enc_layers=2, enc_heads=4, latent=32,
                 dec_type="transformer", dec_size=32, dec_layers=1, dec_heads=4)
    p = _params(g, "encoder.")
    p.update(_params(g, "tdec.", lambda n: n[len("tdec."):]))
    assert set(p) == set(om.param_shapes(cfg))
    means, stds = om.encoder_forward(cfg, p, t("tokens"), t("classes"))
    np.testing.assert_allclose(means.numpy(), g["means"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(stds.numpy(), g["stds"], rtol=1e-4, atol=1e-5)
    logits = om.decoder_logits(cfg, p, t("tokens"), t("seq_lens"), t("z"), t("classes"))
    np.testing.assert_allclose(torch.softmax(logits, -1).numpy(), g["tdec_probs"], rtol=2e-4, atol=1e-6)

    cfg_l = om.Cfg(vocab=293, num_classes=2, enc_size=64, enc_layers=2, enc_heads=4, latent=32,
                   dec_type="lstm", dec_size=32, dec_layers=1)
    pl = _params(g, "encoder.")
    pl.update(_params(g, "ldec.", lambda n: n[len("ldec."):]))
    assert set(pl) == set(om.param_shapes(cfg_l))
    logits = om.decoder_logits(cfg_l, pl, t("tokens"), t("seq_lens"), t("z"), t("classes"))
    np.testing.assert_allclose(torch.softmax(logits, -1).numpy(), g["ldec_probs"], rtol=2e-4, atol=1e-6)


def test_param_count_matches_survey():
    assert sum(int(np.prod(s)) for s in om.param_shapes(om.Cfg()).values()) == 2060325
    assert sum(int(np.prod(s)) for s in om.param_shapes(om.Cfg(dec_type="transformer")).values()) == 2093349
    assert len(om.param_shapes(om.Cfg())) == 46


def test_roll_small_cases():
    """Hand-checked piano-roll cases of the derived spec (oracle/featurise.py:rasterize_sequence)."""
    # on@0, off after 60 ticks (2 slices at res 120 / 4 spq = 30 ticks per slice)
    ids, roll = of.rasterize_sequence([0, 60], [60, 60], [100, 0], 120, 4, 8, 1)
    assert ids == [of.NOTE_ON_FIRST + 60, of.TIMESHIFT_FIRST + 2, of.NOTE_OFF_FIRST + 60]
    assert roll.shape == (1, 8, 128)
    assert roll[0, :, 60].tolist() == [1, 1, 0, 0, 0, 0, 0, 0]
    # zero-length note still occupies its onset slice; unmatched on runs to the end; velocity roll
    ids, roll = of.rasterize_sequence([30, 0, 30], [10, 10, 11], [90, 0, 70], 120, 4, 4, 1, velocity_roll=True)
    assert roll[0, :, 10].tolist() == [0, 90, 0, 0]
    assert roll[0, :, 11].tolist() == [0, 0, 70, 70]
    # long gap quirk: delta 2500 -> three shift tokens of bin 16 -> clock 1440 ticks = slice 48
    ids, roll = of.rasterize_sequence([2500], [5], [1], 120, 4, 64, 1)
    assert ids == [of.TIMESHIFT_FIRST + 16] * 3 + [of.NOTE_ON_FIRST + 5]
    assert roll[0, :, 5].tolist() == [0] * 48 + [1] * 16
    # delta 1000 -> one shift token of bin 0 -> clock does not move
    assert of.played_delta(1000) == 0 and of.played_delta(29) == 0 and of.played_delta(59) == 30


def test_c_oracle_matches_numpy_oracle():
    """oracle/raster.c (the full-size checker / CPU baseline) == oracle/featurise.py on ragged and quirky inputs."""
    import __graft_entry__ as ge
    from oracle import raster_c
    if not raster_c.available():
        ge.build()
    rng = np.random.RandomState(11)
    lens = [0, 1, 5, 32, 33, 200, 0, 64]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    E = int(offs[-1])
    dtick = (rng.randint(0, 5, size=E) * 17).astype(np.int32)
    dtick[rng.rand(E) < 0.05] = rng.randint(1000, 6000, size=int((rng.rand(E) < 0.05).sum()) or 1)[0]
    pitch = rng.randint(40, 46, size=E).astype(np.uint8)
    vel = np.where(rng.rand(E) < 0.5, rng.randint(1, 128, size=E), 0).astype(np.uint8)
    for vr in (False, True):
        for S, L, res, spq in ((64, 64, 120, 4), (16, 8, 96, 4), (128, 100, 220, 3)):
            a = of.rasterize_batch(dtick, pitch, vel, offs, res, spq, S, L, vr)
            b = raster_c.rasterize_batch(dtick, pitch, vel, offs, res, spq, S, L, vr, threads=2)
            assert all(np.array_equal(x, y) for x, y in zip(a, b)), (vr, S, L)

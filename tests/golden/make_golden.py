#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE in the build container.

    python tests/golden/make_golden.py          (needs /root/reference; not run on the GPU box)

The reference cannot run as shipped (mxnet 1.3 and python-midi are not installable here), so the
two missing third-party modules are replaced by the NumPy stand-ins in mx_shim.py; every line of
featurisation / chunking / loss / model-forward logic that produces these vectors is the
reference's (/root/reference/music_style_transfer/...):

  tokens_fixtures.npz   EventBasedMIDIReader.read_file (MIDIUtil/midi_io.py:35-93) on the 37 files
                        under work/data/guitar_bass + the note-event SoA our SMF reader extracted
  rows_fixtures.npz     MelodyDataset._get_token_arrays (VarAutoEncoder/data.py:133-173) on them
  loss_golden.npz       loss.py classes on seeded inputs
  model_toy.npz         Model (Transformer enc + Transformer dec, main.py:14-38 toy config) forward
                        on data.ToyData's arrays (data.py:62-70) with seeded weights and eps
  model_small.npz       Encoder + Decoder (HEAD classes) and LSTMDecoder (model.py:131-203) forward
                        on a ragged random batch, seeded weights
"""
import contextlib
import glob
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
# reference first (its package is also called music_style_transfer), then its inner dir because
# data.py imports ``MIDIUtil`` / ``VarAutoEncoder`` as top-level packages, then the repo for oracle.*
sys.path[:0] = [REF, os.path.join(REF, "music_style_transfer"), HERE]
sys.path.append(REPO)

import mx_shim  # noqa: E402

mx = mx_shim.install()
NDArray = mx_shim.NDArray

with contextlib.redirect_stdout(io.StringIO()):
    from music_style_transfer.MIDIUtil import midi_io as ref_midi_io          # noqa: E402
    from music_style_transfer.VarAutoEncoder import loss as ref_loss          # noqa: E402
    from music_style_transfer.VarAutoEncoder import model as ref_model        # noqa: E402
    from music_style_transfer.VarAutoEncoder import transformer as ref_tf     # noqa: E402
    import VarAutoEncoder.data as ref_data                                     # noqa: E402

from oracle import smf, featurise  # noqa: E402


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


# ----------------------------------------------------------------------------- A1 / A2
def gen_fixture_tokens():
    root = os.path.join(REF, "work/data/guitar_bass")
    reader = quiet(ref_midi_io.EventBasedMIDIReader)
    classes = sorted(next(os.walk(root))[1])
    out = {}
    melodies = {}
    names = []
    for c in classes:
        melodies[c] = []
        for f in sorted(glob.glob(os.path.join(root, c, "*.mid"))):
            mel = quiet(reader.read_file, f)[0]                       # data.py:35 keeps track [0]
            melodies[c].append(mel)
            key = "%s/%s" % (c, os.path.basename(f))
            names.append(key)
            out["ids:" + key] = np.asarray([e.id for e in mel.notes], dtype=np.int32)
            # inputs: note-event SoA of the same (first surviving) track, from our SMF reader
            pat = smf.read_midifile(f)
            for track in pat:
                soa = featurise.note_events_of_track(track)
                if len(featurise.tokenize_note_events(*soa)) >= 10:
                    break
            out["dtick:" + key], out["pitch:" + key], out["vel:" + key] = soa
            out["res:" + key] = np.int32(pat.resolution)
    out["names"] = np.asarray(names)
    np.savez_compressed(os.path.join(HERE, "tokens_fixtures.npz"), **out)

    rows = {}
    for L in (64, 16):
        ds = quiet(ref_data.MelodyDataset, 32, L, melodies)
        rows["tokens_L%d" % L] = ds.tokens.asnumpy()
        rows["labels_L%d" % L] = ds.labels.asnumpy()
        rows["classes_L%d" % L] = ds.classes.asnumpy()
    rows["class_names"] = np.asarray(classes)
    np.savez_compressed(os.path.join(HERE, "rows_fixtures.npz"), **rows)
    print("fixtures: %d files, %d tokens, rows(L=64)=%s" % (
        len(names), sum(len(out["ids:" + n]) for n in names), rows["tokens_L64"].shape))


def gen_midi_bytes():
    """The reference's own .mid inputs (work/data/guitar_bass: the 37 parity fixtures; work/data/splits: 73 more single-track
    files) as raw bytes, so that the SMF parser tests run where /root/reference does not exist (the GPU box).  Per file also
    what the reference reader makes of it: resolution, bpm and, for every track, the note-event walk of
    EventBasedMIDIReader._parse_track (midi_io.py:70-93) through the `midi` stand-in."""
    out, names = {}, []
    for sub in ("guitar_bass/bass", "guitar_bass/guitar", "splits"):
        for f in sorted(glob.glob(os.path.join(REF, "work/data", sub, "*.mid"))):
            key = "%s/%s" % (sub, os.path.basename(f))
            names.append(key)
            with open(f, "rb") as fh:
                out["bytes:" + key] = np.frombuffer(fh.read(), dtype=np.uint8)
            pat = smf.read_midifile(f)
            out["res:" + key] = np.int32(pat.resolution)
            out["bpm:" + key] = np.float64(quiet(ref_midi_io.EventBasedMIDIReader)._extract_bpm(pat))
            out["ntracks:" + key] = np.int32(len(pat))
            for ti, track in enumerate(pat):
                dt, pi, ve = featurise.note_events_of_track(track)
                out["dtick:%d:%s" % (ti, key)], out["pitch:%d:%s" % (ti, key)], out["vel:%d:%s" % (ti, key)] = dt, pi, ve
    out["names"] = np.asarray(names)
    np.savez_compressed(os.path.join(HERE, "midi_fixtures.npz"), **out)
    print("midi bytes: %d files, %d bytes" % (len(names), sum(out["bytes:" + n].size for n in names)))


# ----------------------------------------------------------------------------- losses
def gen_losses():
    rng = np.random.RandomState(0)
    out = {}
    B, Z = 5, 16
    means = rng.randn(B, Z).astype(np.float32)
    stds = (rng.randn(B, Z) * 0.7).astype(np.float32)
    out["kl_means"], out["kl_stds"] = means, stds
    out["kl"] = ref_loss.VariationalKLLoss()(NDArray(means), NDArray(stds)).asnumpy()

    B, T, V = 4, 7, 11
    logits = rng.randn(B, T, V).astype(np.float32)
    probs = mx.nd.softmax(NDArray(logits)).asnumpy()
    labels = rng.randint(1, V, size=(B, T)).astype(np.float32)
    labels[0, 5:] = 0
    labels[2, 3:] = 0
    out["ce_probs"], out["ce_labels"] = probs, labels
    out["ce"] = ref_loss.SoftmaxCrossEntropy(axis=-1, batch_axis=0)(NDArray(probs), NDArray(labels)).asnumpy()

    B, S, P = 3, 8, 128
    pred = (rng.randn(B, S, P) * 2).astype(np.float32)
    label = (rng.rand(B, S, P) < 0.06).astype(np.float32)
    label[2] = 0                                       # a sample without positives
    out["bce_pred"], out["bce_label"] = pred, label
    for tag, kw in (("default", {}),
                    ("smooth", dict(label_smoothing=0.1)),
                    ("noweight", dict(negative_label_downweighting=False)),
                    ("fromsig", dict(from_sigmoid=True))):
        x = NDArray(1 / (1 + np.exp(-pred))) if tag == "fromsig" else NDArray(pred)
        out["bce_" + tag] = ref_loss.BinaryCrossEntropy(**kw)(x, NDArray(label)).asnumpy()
    np.savez_compressed(os.path.join(HERE, "loss_golden.npz"), **out)
    print("losses ok")


# ----------------------------------------------------------------------------- model forward
def set_params(block, seed):
    """Seeded Xavier-like uniform weights, small random biases / gammas (so every parameter matters)."""
    rng = np.random.RandomState(seed)
    vals = {}
    for name, p in block.collect_params().items():
        shape = p.shape
        if name.endswith("gamma"):
            v = 1.0 + 0.1 * rng.randn(*shape)
        elif name.endswith("bias") or name.endswith("beta"):
            v = 0.1 * rng.randn(*shape)
        else:
            scale = np.sqrt(3.0 / ((shape[0] + shape[1]) / 2.0))
            v = rng.uniform(-scale, scale, size=shape)
        p.value = NDArray(v.astype(np.float32))
        vals[name] = p.value.asnumpy()
    return vals


def tcfg(size, layers, heads, vocab):
    return ref_tf.TransformerConfig(model_size=size, dropout=0.0, num_layers=layers, num_heads=heads, vocab_size=vocab)


def gen_model_toy():
    V, C = 10, 3
    cfg = ref_model.ModelConfig(
        encoder_config=ref_model.EncoderConfig(transformer_config=tcfg(32, 1, 2, V), latent_dim=16,
                                               num_classes=C, input_dim=V),
        decoder_config=ref_model.DecoderConfig(transformer_config=tcfg(32, 1, 2, V), latent_dim=16,
                                               num_classes=C, output_dim=V))
    m = quiet(ref_model.Model, cfg)
    params = set_params(m, 1)
    tokens = np.array([[1, 5, 6, 7, 0], [1, 6, 7, 8, 0], [1, 7, 8, 9, 0]], dtype=np.float32)   # data.py:62-64
    seq_lens = np.array([4, 4, 4], dtype=np.float32)                                           # data.py:65
    classes = np.array([0, 1, 2], dtype=np.float32)                                            # data.py:66
    labels = np.array([[5, 6, 7, 2, 0], [6, 7, 8, 2, 0], [7, 8, 9, 2, 0]], dtype=np.float32)   # data.py:67-69
    eps = np.random.RandomState(2).randn(3, 16).astype(np.float32)
    mx_shim.set_normal(eps)
    probs, means, stds = quiet(m, NDArray(tokens), NDArray(seq_lens), NDArray(classes))
    ce = ref_loss.SoftmaxCrossEntropy(axis=-1, batch_axis=0)(probs, NDArray(labels))
    kl = ref_loss.VariationalKLLoss()(means, stds)
    out = {"param:" + k: v for k, v in params.items()}
    out.update(tokens=tokens, seq_lens=seq_lens, classes=classes, labels=labels, eps=eps,
               probs=probs.asnumpy(), means=means.asnumpy(), stds=stds.asnumpy(),
               ce=ce.asnumpy(), kl=kl.asnumpy())
    np.savez_compressed(os.path.join(HERE, "model_toy.npz"), **out)
    print("toy model ok: ce", out["ce"], "kl", out["kl"])


def gen_model_small():
    """HEAD Encoder / Decoder and the LSTMDecoder on a ragged batch (D=64, 4 heads, 2 enc layers)."""
    V, C, D, Z, Hd = 293, 2, 64, 32, 32
    rng = np.random.RandomState(3)
    B, T = 6, 12
    lens = np.array([12, 9, 5, 12, 2, 7])
    tokens = np.zeros((B, T), dtype=np.float32)
    for b in range(B):
        tokens[b, 0] = 1
        tokens[b, 1:lens[b]] = rng.randint(3, V, size=lens[b] - 1)
    seq_lens = lens.astype(np.float32)
    classes = rng.randint(0, C, size=B).astype(np.float32)
    out = dict(tokens=tokens, seq_lens=seq_lens, classes=classes)

    enc = ref_model.Encoder(ref_model.EncoderConfig(transformer_config=tcfg(D, 2, 4, V), latent_dim=Z,
                                                     num_classes=C, input_dim=V))
    for k, v in set_params(enc, 4).items():
        out["param:encoder." + k] = v
    means, stds = quiet(enc, NDArray(tokens), NDArray(seq_lens), NDArray(classes))
    out["means"], out["stds"] = means.asnumpy(), stds.asnumpy()
    z = (means.asnumpy() + rng.randn(B, Z).astype(np.float32) * stds.asnumpy()).astype(np.float32)
    out["z"] = z

    dec = ref_model.Decoder(ref_model.DecoderConfig(transformer_config=tcfg(Hd, 1, 4, V), latent_dim=Z,
                                                     num_classes=C, output_dim=V))
    for k, v in set_params(dec, 5).items():
        out["param:tdec.decoder." + k] = v
    probs = quiet(dec.forward_train, mx.nd, NDArray(tokens), NDArray(seq_lens), NDArray(z), NDArray(classes))
    out["tdec_probs"] = probs.asnumpy()

    # LSTMDecoder reads config.lstm_config (model.py:139-153), which DecoderConfig at HEAD cannot
    # carry (main.py:109-117 TypeError) -> hand it a namespace with the fields it touches.
    lcfg = types.SimpleNamespace(latent_dim=Z, num_classes=C, output_dim=V,
                                 lstm_config=types.SimpleNamespace(hidden_dim=Hd, n_layers=1, dropout=0.0))
    ldec = ref_model.LSTMDecoder(lcfg)
    for k, v in set_params(ldec, 6).items():
        out["param:ldec.decoder." + k] = v
    probs = quiet(ldec.forward_train, mx.nd, NDArray(tokens), NDArray(seq_lens), NDArray(z), NDArray(classes))
    out["ldec_probs"] = probs.asnumpy()
    np.savez_compressed(os.path.join(HERE, "model_small.npz"), **out)
    print("small model ok", out["means"].shape, out["tdec_probs"].shape, out["ldec_probs"].shape)


if __name__ == "__main__":
    gen_fixture_tokens()
    gen_midi_bytes()
    gen_losses()
    gen_model_toy()
    gen_model_small()

"""CPU tests of the host-side mirrors: vocabulary, SMF round trip, Melody/event factories, dataset chunking
against rows produced by the reference's own MelodyDataset (tests/golden/rows_fixtures.npz), config flags,
YAML config round trip, synthetic workload generator, data-parallel sharding rule (gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from musicstyletransfer_b200.MIDIUtil import defaults, smf
from musicstyletransfer_b200.MIDIUtil import Melody as M

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_vocabulary_matches_reference_constants():
    assert (defaults.PAD_ID, defaults.SOS_ID, defaults.EOS_ID) == (0, 1, 2)
    assert defaults.NOTE_ON_EVENTS == (3, 130) and defaults.NOTE_OFF_EVENTS == (131, 258)
    assert defaults.TIMESHIFT_EVENTS == (259, 292) and defaults.NUM_EVENTS == 293 and defaults.NUM_BINS == 34


def test_event_factories_and_ranges():
    assert M.create_note_on_event(60).id == 63 and M.create_note_off_event(60).id == 191
    assert M.create_timeshift_event(0).id == 259 and M.create_timeshift_event(999).id == 292
    with pytest.raises(AssertionError):
        M.create_timeshift_event(1000)
    with pytest.raises(ValueError):
        M.create_event_from_id(293)
    with pytest.raises(ValueError):
        M.create_event_from_id(2)
    mel = M.get_melody_from_ids([1, 63, 260, 191, 0, 0])
    assert [type(e).__name__ for e in mel] == ["NoteOnEvent", "TimeshiftEvent", "NoteOffEvent"]
    assert mel[1].get_tick_delay() == 30


def test_smf_write_read_roundtrip(tmp_path):
    from musicstyletransfer_b200.MIDIUtil.midi_io import MelodyWriter, note_event_soa
    mel = M.get_melody_from_ids([63, 260, 191, 70, 275, 198])
    mel.resolution = 120
    path = str(tmp_path / "x.mid")
    MelodyWriter().write_to_file(path, mel)
    pat = smf.read_midifile(path)
    assert pat.resolution == 120 and len(pat) == 1
    dt, pi, ve = note_event_soa(pat[0])
    assert dt.tolist() == [0, 30, 0, 480] and pi.tolist() == [60, 60, 67, 67] and ve.tolist() == [127, 0, 127, 0]
    assert abs(pat[0][0].get_bpm() - 120.0) < 1e-6


def test_smf_reader_agrees_with_oracle_reader_on_fixture_soa(golden_dir):
    """The product's SMF reader and the oracle's independent one fold a synthetic multi-event file identically."""
    from oracle import smf as osmf, featurise as of
    from musicstyletransfer_b200.MIDIUtil.midi_io import note_event_soa
    import tempfile
    pat = smf.Pattern(resolution=96)
    tr = smf.Track()
    tr.append(smf.SetTempoEvent(tick=0, data=[7, 161, 32]))
    rng = np.random.RandomState(0)
    for i in range(200):
        cls = smf.NoteOnEvent if rng.rand() < 0.6 else smf.NoteOffEvent
        tr.append(cls(tick=int(rng.randint(0, 300)), pitch=int(rng.randint(0, 128)), velocity=int(rng.randint(0, 128))))
        if i % 17 == 0:
            e = smf.OtherChannelEvent(tick=int(rng.randint(0, 50)), data=[7, 100])
            tr.append(e)
    tr.append(smf.EndOfTrackEvent(tick=1))
    pat.append(tr)
    with tempfile.NamedTemporaryFile(suffix=".mid", delete=False) as f:
        name = f.name
    smf.write_midifile(name, pat)
    a = note_event_soa(smf.read_midifile(name)[0])
    b = of.note_events_of_track(osmf.read_midifile(name)[0])
    os.unlink(name)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_melody_dataset_rows_match_reference(golden_dir):
    """A2: MelodyDataset._get_token_arrays == the reference's (data.py:133-173) on the 37 fixtures, L = 64 and 16."""
    from musicstyletransfer_b200.VarAutoEncoder.data import MelodyDataset
    g = np.load(os.path.join(golden_dir, "tokens_fixtures.npz"))
    r = np.load(os.path.join(golden_dir, "rows_fixtures.npz"))
    melodies = {}
    for name in g["names"]:
        mel = M.Melody()
        mel.notes = [M.create_event_from_id(int(i)) for i in g["ids:" + name]]
        melodies.setdefault(name.split("/")[0], []).append(mel)
    for L in (64, 16):
        ds = MelodyDataset(32, L, melodies)
        assert np.array_equal(ds.tokens, r["tokens_L%d" % L])
        assert np.array_equal(ds.labels, r["labels_L%d" % L])
        assert np.array_equal(ds.classes, r["classes_L%d" % L])
    ds = MelodyDataset(32, 64, melodies)
    n = 0
    for batch in ds:
        tokens, seq_lens, classes = batch.data
        assert tokens.shape[0] == 32 and batch.label[0].shape == tokens.shape
        assert int(seq_lens.max()) == tokens.shape[1]
        assert np.array_equal(seq_lens.numpy(), (tokens.numpy() != 0).sum(1))
        n += 1
    assert n == 28          # 880 rows -> 28 batches of 32


def test_cli_flags_and_config_roundtrip(tmp_path):
    from musicstyletransfer_b200.VarAutoEncoder.config import get_config, Config
    from musicstyletransfer_b200.VarAutoEncoder import model
    from musicstyletransfer_b200.VarAutoEncoder.transformer import TransformerConfig
    a = get_config("--batch-size 32 --kl-loss 1.0 --validation-split 0.0 --max-seq-len 64 --optimizer adam "
                   "--optimizer-params clip_gradient:1.0 --e-n-layers 2 --e-dropout 0.2 --e-rnn-hidden-dim 256 "
                   "--latent-dim 256 --d-rnn-hidden-dim 128 --g-n-layers 1 --noise-dim 64".split())
    assert a.batch_size == 32 and a.e_num_heads == 8 and a.learning_rate == 3e-4 and a.sampling_type == "sampling"
    from musicstyletransfer_b200.VarAutoEncoder.trainer import OptimizerConfig
    assert OptimizerConfig("adam", a.optimizer_params, 3e-4).params_to_dict() == {"clip_gradient": 1.0}
    c = model.ModelConfig(model.EncoderConfig(TransformerConfig(32, 0.0, 1, 2, 10), 16, 3, 10),
                          model.DecoderConfig(16, 3, 10, transformer_config=TransformerConfig(32, 0.0, 1, 2, 10)))
    c.save(str(tmp_path / "config"))
    c2 = Config.load(str(tmp_path / "config"))
    assert c2 == c and c2.encoder_config.transformer_config.num_heads == 2
    c.freeze()
    with pytest.raises(AttributeError):
        c.encoder_config.latent_dim = 3
    assert c.copy(decoder_config=None).decoder_config is None
    ec = model.to_engine_config(c2)
    assert ec.dec_type == "transformer" and ec.enc_size == 32 and ec.latent == 16


def test_checkpoint_index_uses_whole_number(tmp_path):
    from musicstyletransfer_b200.VarAutoEncoder import utils
    for n in (1, 9, 12):
        (tmp_path / ("params.%d" % n)).write_bytes(b"")
    assert utils.get_latest_checkpoint_index(str(tmp_path)) == 12
    empty = tmp_path / "empty"
    empty.mkdir()
    with pytest.raises(ValueError):
        utils.get_latest_checkpoint_index(str(empty))


def test_synthetic_rows():
    from musicstyletransfer_b200 import synth
    tok, lens, cls, lab = synth.token_rows_4_4(16, 64, seed=0)
    assert tok.shape == (16, 65) and (tok[:, 0] == 1).all() and (lab[:, 64] == 2).all() and (lens == 65).all()
    assert np.array_equal(tok[:, 1:], lab[:, :64])
    on = tok[:, 1::3] - 3
    assert ((on[cls == 0] >= 28) & (on[cls == 0] <= 62)).all() and ((on[cls == 1] >= 40) & (on[cls == 1] <= 88)).all()


DP_WORKER = r'''
import os, sys
sys.path.insert(0, %(repo)r)
import torch, torch.distributed as dist
from oracle import model as om
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
cfg = om.Cfg(enc_size=32, enc_layers=1, enc_heads=2, latent=8, dec_type="lstm", dec_size=32)
p = om.init_params(cfg, seed=0)
p["encoder.latent_proj.bias"][8:] = 2.0
g = torch.Generator().manual_seed(0)
B, T = 8, 9
tokens = torch.randint(3, 293, (B, T), generator=g).float(); tokens[:, 0] = 1
lens = torch.full((B,), float(T)); classes = torch.randint(0, 2, (B,), generator=g).float()
labels = torch.cat([tokens[:, 1:], torch.full((B, 1), 2.0)], 1); eps = torch.randn(B, 8, generator=g)
def grads(sl):
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    loss = om.step_losses(cfg, q, tokens[sl], lens[sl], classes[sl], labels[sl], eps[sl])[0]
    loss.sum().backward()
    return torch.cat([(v.grad if v.grad is not None else torch.zeros_like(v)).reshape(-1) for v in q.values()])
# the rule of trainer.Trainer._shard / _step: rows rank::world, SUM all-reduce of the flat arena, rescale 1/B_global
flat = grads(slice(rank, None, world))
dist.all_reduce(flat)
full = grads(slice(None))
err = float((flat - full).abs().max() / full.abs().max())
assert err < 1e-5, err
if rank == 0:
    print("DP_OK", err)
dist.destroy_process_group()
'''


def test_data_parallel_rule_gloo_world2(tmp_path):
    """Rows rank::world + SUM all-reduce of the flat gradient arena == full-batch gradient (trainer.py:176-177:
    loss.backward() sums over the batch, step(batch_size) rescales by 1/B_global)."""
    script = tmp_path / "dp_worker.py"
    script.write_text(DP_WORKER % {"repo": REPO})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "DP_OK" in out.stdout


def test_param_arena_layout_for_tma_and_bf16_shadow():
    """Host-side layout rules of the flat parameter arena (engine.ParamArena): every tensor starts on a 32-byte boundary
    (so that its bf16 shadow at the same element offset is a 16-byte aligned TMA base), the K | Q | V projections are
    adjacent without gaps (one [3D, D] GEMM), and the parameter count is the reference's."""
    import torch
    from musicstyletransfer_b200.engine import ParamArena, VAEConfig, param_entries
    for dec, n_ref in (("lstm", 2060325), ("transformer", 2093349)):
        cfg = VAEConfig(dec_type=dec)
        a = ParamArena(cfg, torch.device("cpu"))
        assert a.n_params == n_ref
        for name, (off, n, shape) in a.offsets.items():
            assert off % 8 == 0, name
        D = cfg.enc_size
        for l in range(cfg.enc_layers):
            pre = "encoder.encoder.layer%d.self_attention." % l
            ok, oq, ov = (a.offsets[pre + k + ".weight"][0] for k in ("W_k", "W_q", "W_v"))
            assert oq == ok + D * D and ov == oq + D * D
            assert a.span(pre + "W_k.weight", pre + "W_v.weight").numel() == 3 * D * D
        assert [n for n, _ in param_entries(cfg)] == a.names()


def test_cli_accepts_bf16_precision():
    from musicstyletransfer_b200.VarAutoEncoder import config as cfgmod
    got = cfgmod.get_config(["--precision", "bf16"])
    args = got[0] if isinstance(got, tuple) else got
    assert args.precision == "bf16"


def test_model_folder_files_are_data_not_code(tmp_path):
    """a model folder may come from somewhere else.  `config` is loaded with a SafeLoader that only knows the
    registered Config classes, `train_state.pkl` with an unpickler that only constructs TrainingState."""
    import os
    import pickle
    from musicstyletransfer_b200.VarAutoEncoder import trainer, utils
    from musicstyletransfer_b200.VarAutoEncoder.config import Config
    import yaml
    (tmp_path / "config").write_text("!!python/object/apply:os.system ['echo pwned']\n")
    with pytest.raises(yaml.YAMLError):
        Config.load(str(tmp_path / "config"))
    st = trainer.TrainingState()
    st.n_batches, st.best_resconstruction_loss = 7, np.float64(1.5)
    utils.save_object(st, str(tmp_path / "train_state.pkl"))
    back = utils.load_object(str(tmp_path / "train_state.pkl"))
    assert back.n_batches == 7 and float(back.best_resconstruction_loss) == 1.5

    class Evil:
        def __reduce__(self):
            return (os.system, ("echo pwned",))
    with open(tmp_path / "evil.pkl", "wb") as f:
        pickle.dump(Evil(), f)
    with pytest.raises(pickle.UnpicklingError):
        utils.load_object(str(tmp_path / "evil.pkl"))


def test_seed_flag_changes_initial_weights_signature():
    """Model(seed=...) must reach initialize(); checked on the host through the stored attribute (the arena
    itself needs a GPU: tests/test_modules_gpu.py)."""
    import inspect
    from musicstyletransfer_b200.VarAutoEncoder import model
    src = inspect.getsource(model.Model)
    assert "self.engine_seed = seed" in src and "engine_seed = 0" not in src
